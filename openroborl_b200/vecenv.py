"""VecEnv-style facade over the batched CUDA environment.

The reference's `WrapperEnv` is NOT a VecEnv (`imitation_runners.py:204` relies on that) and never resets robots
individually.  Newer stable-baselines style algorithms expect the opposite contract: `reset() -> obs[N, ...]`,
`step(actions) -> (obs, rewards, dones, infos)` where a finished environment is reset inside `step`, the returned
observation is the first one of the new episode and the last one of the old episode travels in
`infos[i]["terminal_observation"]`.  This adapter provides that contract on top of `BatchedImitationEnv` without
in-kernel auto-reset: one step launch, one masked reset launch when some robot finished.
"""
import numpy as np
import torch

from .compat import _box
from .env import BatchedImitationEnv


class VecImitationEnv(object):
    def __init__(self, robot="laikago", motions=None, num_envs=1, device=0, seed=0, max_episode_steps=600, **kwargs):
        self.env = BatchedImitationEnv(robot, motions, num_robot=num_envs, device=device, seed=seed, auto_reset=False,
                                       max_episode_steps=max_episode_steps, **kwargs)
        self.num_envs = int(num_envs)
        low, high = self.env.observation_bounds()
        self.observation_space = _box(low, high)
        self.action_space = _box(-2 * np.pi * np.ones(12), 2 * np.pi * np.ones(12))          # minitaur.py:145-148
        self._actions = None
        self._ep_ret = np.zeros(self.num_envs, np.float64)
        self._ep_len = np.zeros(self.num_envs, np.int64)

    def reset(self):
        self._ep_ret[:] = 0; self._ep_len[:] = 0
        return self.env.reset().cpu().numpy()

    def step_async(self, actions):
        self._actions = torch.as_tensor(np.ascontiguousarray(actions, dtype=np.float32), device=self.env.device)

    def step_wait(self):
        obs, rew, done = self.env.step(self._actions)
        obs_h, rew_h, done_h = obs.cpu().numpy().copy(), rew.cpu().numpy().copy(), done.cpu().numpy().astype(bool)
        self._ep_ret += rew_h; self._ep_len += 1
        infos = [{} for _ in range(self.num_envs)]
        if done_h.any():
            new_obs = self.env.reset(mask=done).cpu().numpy()
            for i in np.nonzero(done_h)[0]:
                infos[i]["terminal_observation"] = obs_h[i].copy()
                infos[i]["episode"] = {"r": float(self._ep_ret[i]), "l": int(self._ep_len[i])}
                obs_h[i] = new_obs[i]
            self._ep_ret[done_h] = 0; self._ep_len[done_h] = 0
        return obs_h, rew_h, done_h, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.env.close()

    def seed(self, seed=None):
        return [seed] * self.num_envs

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False] * self.num_envs

    def get_attr(self, name, indices=None):
        return [getattr(self.env, name)] * self.num_envs
