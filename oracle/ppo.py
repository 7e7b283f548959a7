"""Rollout post-processing and the actor network of the reference, restated in NumPy.

Pinned: add_vtarg_and_adv and mpi_adam_update against outputs of the reference's own functions
(tests/golden/ref_ppo_kat.npz, scripts/make_ppo_golden.py); the PPO1 loss is restated from reading only.

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


# ---- PPO1 loss, its gradient and MpiAdam (row f3) -------------------------------------------------------------------
PARAM_ORDER = ("pi_fc0/w", "pi_fc0/b", "vf_fc0/w", "vf_fc0/b", "pi_fc1/w", "pi_fc1/b", "vf_fc1/w", "vf_fc1/b",
               "vf/w", "vf/b", "pi/w", "pi/b")          # order of the reference's trainable variables (policy zips: parameter_list)
LOG_STD = np.log(0.125)


def _mlp_fwd_cache(params, obs, head):
    z0 = obs @ params[head + "_fc0/w"] + params[head + "_fc0/b"]
    h0 = np.maximum(z0, 0)
    z1 = h0 @ params[head + "_fc1/w"] + params[head + "_fc1/b"]
    h1 = np.maximum(z1, 0)
    out = h1 @ params[head + "/w"] + params[head + "/b"]
    return out, (obs, z0, h0, z1, h1)


def _mlp_bwd(params, head, cache, dout, grads):
    obs, z0, h0, z1, h1 = cache
    grads[head + "/w"] = h1.T @ dout; grads[head + "/b"] = dout.sum(0)
    dz1 = (dout @ params[head + "/w"].T) * (z1 > 0)
    grads[head + "_fc1/w"] = h0.T @ dz1; grads[head + "_fc1/b"] = dz1.sum(0)
    dz0 = (dz1 @ params[head + "_fc1/w"].T) * (z0 > 0)
    grads[head + "_fc0/w"] = obs.T @ dz0; grads[head + "_fc0/b"] = dz0.sum(0)


def ppo_losses_and_grad(params, old_params, obs, actions, atarg, ret, clip_param, entcoeff, lrmult=1.0):
    """ppo_imitation.py:176-218 for the fixed-sigma Gaussian policy, float64, with the analytic gradient of
    total_loss = pol_surr + pol_entpen + vf_loss.  Returns ([pol_surr, pol_entpen, vf_loss, kl, ent], {name: grad})."""
    P = {k: np.asarray(v, dtype=np.float64) for k, v in params.items()}
    O = {k: np.asarray(v, dtype=np.float64) for k, v in old_params.items()}
    obs = np.asarray(obs, np.float64); actions = np.asarray(actions, np.float64)
    atarg = np.asarray(atarg, np.float64); ret = np.asarray(ret, np.float64)
    n, sig2 = obs.shape[0], np.exp(2 * LOG_STD)
    clip = clip_param * lrmult
    mean, cpi = _mlp_fwd_cache(P, obs, "pi")
    mean_old, _ = _mlp_fwd_cache(O, obs, "pi")
    val, cvf = _mlp_fwd_cache(P, obs, "vf")
    logp = -0.5 * np.sum((actions - mean) ** 2, axis=1) / sig2
    logp_old = -0.5 * np.sum((actions - mean_old) ** 2, axis=1) / sig2       # the constants cancel in the ratio
    ratio = np.exp(logp - logp_old)
    surr1 = ratio * atarg
    surr2 = np.clip(ratio, 1 - clip, 1 + clip) * atarg
    pol_surr = -np.mean(np.minimum(surr1, surr2))
    ent = mean.shape[1] * (LOG_STD + 0.5 * np.log(2 * np.pi * np.e))
    kl = np.mean(np.sum((mean_old - mean) ** 2, axis=1) / (2 * sig2))
    vf_loss = np.mean((val[:, 0] - ret) ** 2)
    # gradient: d pol_surr / d ratio = -atarg / n where the unclipped branch is the minimum (ties -> surr1, like tf.minimum's
    # subgradient convention is irrelevant on a set of measure zero)
    use1 = surr1 <= surr2
    inside = (ratio >= 1 - clip) & (ratio <= 1 + clip)
    dratio = np.where(use1 | inside, -atarg / n, 0.0)
    dlogp = dratio * ratio
    dmean = dlogp[:, None] * (actions - mean) / sig2
    dval = (2.0 / n) * (val[:, 0] - ret)
    grads = {}
    _mlp_bwd(P, "pi", cpi, dmean, grads)
    _mlp_bwd(P, "vf", cvf, dval[:, None], grads)
    return [pol_surr, -entcoeff * ent, vf_loss, kl, ent], grads


def mpi_adam_update(theta, exp_avg, exp_avg_sq, step, global_grad, learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8):
    """stable_baselines/common/mpi_adam.py:40-63 after the all-reduce (global_grad = mean over ranks), float32 state."""
    g = np.asarray(global_grad, dtype=np.float32)
    step = step + 1
    step_size = learning_rate * np.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
    exp_avg = beta1 * exp_avg + (1 - beta1) * g
    exp_avg_sq = beta2 * exp_avg_sq + (1 - beta2) * (g * g)
    delta = (-step_size) * exp_avg / (np.sqrt(exp_avg_sq) + epsilon)
    return (theta + delta).astype(np.float32), exp_avg.astype(np.float32), exp_avg_sq.astype(np.float32), step
