"""Rollout post-processing and the actor network of the reference, restated in NumPy.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Sources: `OpenRoboRL/agents/ppo_imitation.py:68-93`
(add_vtarg_and_adv, including its non-terminal index `(step*N+i) + (1+i)`), `:329-338` (per-robot
advantage standardisation), `OpenRoboRL/agents/imitation_policies.py:44-110` (fixed-std diagonal Gaussian
MLP policy, ReLU, sigma = 0.125) and `OpenRoboRL/run.py:101-105` (net_arch pi/vf [512, 256]).
"""
import numpy as np


def add_vtarg_and_adv(rewards, vpred, nextvpreds, episode_starts, num_robot, gamma, lam):
    """Flat time-major interleaved arrays of length T*num_robot -> (adv, tdlamret), float32 like the reference."""
    episode_starts = np.append(episode_starts, [False] * num_robot)
    rew_len = len(rewards)
    adv = np.empty(rew_len, "float32")
    lastgaelam = [0] * num_robot
    for step in reversed(range(int(rew_len / num_robot))):
        for i in range(num_robot):
            nonterminal = 1 - float(episode_starts[(step * num_robot + i) + (1 + i)])
            delta = rewards[step * num_robot + i] + gamma * nextvpreds[step * num_robot + i] - vpred[step * num_robot + i]
            adv[step * num_robot + i] = lastgaelam[i] = delta + gamma * lam * nonterminal * lastgaelam[i]
    return adv, adv + vpred


def standardize_per_robot(adv, num_robot):
    a = adv.reshape(-1, num_robot).astype(np.float64)
    return ((a - a.mean(axis=0)) / a.std(axis=0)).reshape(-1)


def mlp_forward(params, obs, head):
    """head = 'pi' or 'vf': relu(obs W0 + b0) -> relu(. W1 + b1) -> . W + b (float32 like the TF graph)."""
    h = np.maximum(obs.astype(np.float32) @ params[head + "_fc0/w"] + params[head + "_fc0/b"], 0)
    h = np.maximum(h @ params[head + "_fc1/w"] + params[head + "_fc1/b"], 0)
    return h @ params[head + "/w"] + params[head + "/b"]
