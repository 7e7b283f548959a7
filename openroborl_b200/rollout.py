"""Batched policy inference + rollout collection + GAE on the device.

Replaces, for large batches, the reference's per-robot, batch-1 `policy.step` loop and flat NumPy buffers
(`OpenRoboRL/agents/imitation_runners.py:22-208`) and `add_vtarg_and_adv` + per-robot advantage
standardisation (`OpenRoboRL/agents/ppo_imitation.py:68-93, 329-338`).  The actor/critic are the reference's
fixed-sigma MLPs (`agents/imitation_policies.py:44-110`, `run.py:101-105`): 160 -> 512 -> 256 -> {12, 1}, ReLU,
sigma = 0.125.  The two dense layers are plain library GEMMs (torch/cuBLAS): this is not the env hot path
and nothing here is hand-fused yet.  Buffers keep the runner's time-major interleaved layout
(index t * num_robot + j) so a learner written for the reference can consume them unchanged.
"""
import ctypes as C
import io
import json
import math
import zipfile

import numpy as np
import torch

from . import _abi
from . import dist as odist

LOG_STD = math.log(0.125)          # imitation_policies.py:48,106 (pi_init_std=0.125, not trainable)
_KEYS = ("pi_fc0", "pi_fc1", "pi", "vf_fc0", "vf_fc1", "vf")


def load_policy_zip(path):
    """Read a stable-baselines 2.10 policy zip (`data` json, `parameters` npz, `parameter_list` json;
    reference stable_baselines/common/base_class.py:629-690).  Returns {'pi_fc0/w': ndarray, ...}."""
    with zipfile.ZipFile(path) as z:
        names = json.loads(z.read("parameter_list"))
        npz = np.load(io.BytesIO(z.read("parameters")))
        out = {}
        for n in names:
            key = n[len("model/"):-len(":0")] if n.startswith("model/") and n.endswith(":0") else n
            out[key] = npz[n]
    return out


def load_policy_npz(path):
    z = np.load(path)
    return {k.rsplit("_", 1)[0] + "/" + k.rsplit("_", 1)[1]: z[k] for k in z.files if k.rsplit("_", 1)[-1] in ("w", "b")}


class MlpGaussianPolicy(object):
    """Actor + critic of the reference on one GPU; parameters are torch tensors (float32)."""

    def __init__(self, params, device, tf32=False):
        self.device = torch.device(device)
        self.p = {}
        for k in _KEYS:
            self.p[k + "/w"] = torch.as_tensor(np.asarray(params[k + "/w"], dtype=np.float32), device=self.device).contiguous()
            self.p[k + "/b"] = torch.as_tensor(np.asarray(params[k + "/b"], dtype=np.float32), device=self.device).contiguous()
        self.tf32 = tf32
        self.std = math.exp(LOG_STD)

    def parameters(self):
        return [self.p[k + s] for k in _KEYS for s in ("/w", "/b")]

    def broadcast(self, src=0):
        """MpiAdam.sync (stable_baselines/common/mpi_adam.py:65-70): one flat broadcast from rank `src`."""
        odist.broadcast_policy(self.parameters(), src=src)

    def _mlp(self, obs, head):
        p = self.p
        h = torch.relu(torch.addmm(p[head + "_fc0/b"], obs, p[head + "_fc0/w"]))
        h = torch.relu(torch.addmm(p[head + "_fc1/b"], h, p[head + "_fc1/w"]))
        return torch.addmm(p[head + "/b"], h, p[head + "/w"])

    @torch.no_grad()
    def step(self, obs, deterministic=False, generator=None, noise=None):
        """obs [N,160] -> (action [N,12], value [N], neglogp [N]); mirrors ActorCriticPolicy.step.
        `noise`: pre-drawn standard normal [N,12] (CUDA-graph friendly) instead of drawing from `generator`."""
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = self.tf32
        try:
            mean = self._mlp(obs, "pi")
            value = self._mlp(obs, "vf")[:, 0]
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        if deterministic:
            action = mean
        elif noise is not None:
            action = mean + self.std * noise
        else:
            action = mean + self.std * torch.randn(mean.shape, device=mean.device, dtype=mean.dtype, generator=generator)
        # DiagGaussianProbabilityDistribution.neglogp
        neglogp = 0.5 * (((action - mean) / self.std) ** 2).sum(dim=1) + 0.5 * math.log(2.0 * math.pi) * mean.shape[1] \
            + LOG_STD * mean.shape[1]
        return action, value, neglogp


def gae(rewards, vpred, nextvpred, episode_starts, gamma, lam, reference_index_quirk=True, standardize=True):
    """Device GAE over [T,N] buffers (C ABI `orl_gae`).  Returns (adv [T,N], tdlamret [T,N])."""
    lib = _abi.load_library()
    T, N = rewards.shape
    for t in (rewards, vpred, nextvpred):
        if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous() or tuple(t.shape) != (T, N):
            raise ValueError("rewards/vpred/nextvpred must be contiguous float32 CUDA tensors of the same [T,N] shape")
    es = episode_starts.to(torch.uint8).contiguous()
    adv = torch.empty_like(rewards)
    ret = torch.empty_like(rewards)
    P = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.orl_gae(P(rewards), P(vpred), P(nextvpred), P(es), T, N, float(gamma), float(lam), int(reference_index_quirk),
                     int(standardize), P(adv), P(ret), C.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream))
    if rc != 0:
        raise RuntimeError("orl_gae failed: %d" % rc)
    return adv, ret


class RolloutCollector(object):
    """Collects `horizon` control steps of every robot with batched policy inference in the loop.

    reset_all_on_any_done=True reproduces the runner's episode handling (imitation_runners.py:185-205): when
    ANY robot is done the whole env is reset, `nextvpreds` of that row is zeroed for every robot and
    `episode_start[j] = done[j]`.  With an auto-reset env (`BatchedImitationEnv(auto_reset=True)`) pass False:
    robots restart individually inside the step kernel and only their own bootstrap value is cut.
    """

    def __init__(self, env, policy, horizon, gamma=0.95, lam=0.95, reset_all_on_any_done=True, seed=0, use_cuda_graph=False):
        self.env, self.policy, self.T = env, policy, int(horizon)
        self.gamma, self.lam = gamma, lam
        self.reset_all = reset_all_on_any_done
        # The T-step loop is ~25 small launches per step (policy GEMMs, elementwise bookkeeping, the env step kernel):
        # with an auto-reset env nothing in it depends on the host, so it can be captured once and replayed.
        if use_cuda_graph and reset_all_on_any_done:
            raise ValueError("use_cuda_graph needs reset_all_on_any_done=False (the reference's reset-all is a host decision)")
        if not reset_all_on_any_done and not getattr(env, "auto_reset", True):
            raise ValueError("reset_all_on_any_done=False needs an auto-reset env (finished robots would be stepped forever)")
        self.use_graph = bool(use_cuda_graph)
        self._graph, self._seg, self._graph_version = None, None, None
        N, dev = env.num_robot, env.device
        self.obs_buf = torch.empty((self.T, N, 160), dtype=torch.float32, device=dev)
        self.act_buf = torch.empty((self.T, N, 12), dtype=torch.float32, device=dev)
        self.rew_buf = torch.empty((self.T, N), dtype=torch.float32, device=dev)
        self.vpred_buf = torch.empty((self.T, N), dtype=torch.float32, device=dev)
        self.nextv_buf = torch.zeros((self.T, N), dtype=torch.float32, device=dev)
        self.done_buf = torch.empty((self.T, N), dtype=torch.uint8, device=dev)
        self.start_buf = torch.empty((self.T, N), dtype=torch.uint8, device=dev)
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(seed)
        self.low, self.high = -2.0 * math.pi, 2.0 * math.pi            # env.action_space (minitaur.py:145-148)
        self.obs = env.reset().clone()
        self.episode_start = torch.ones(N, dtype=torch.uint8, device=dev)
        self.noise_buf = torch.empty((self.T, N, 12), dtype=torch.float32, device=dev) if self.use_graph else None

    @torch.no_grad()
    def collect(self):
        if not self.use_graph:
            return self._collect()
        if self._graph is not None and self._graph_version != (getattr(self.env, "config_version", 0), self.policy.tf32):
            self._graph = None                       # a setter changed by-value kernel arguments (or the matmul mode) since the capture
        if self._graph is None:                      # warm up on a side stream, then capture the whole segment
            self.noise_buf.normal_(generator=self.gen)
            side = torch.cuda.Stream(device=self.env.device)
            side.wait_stream(torch.cuda.current_stream(self.env.device))
            with torch.cuda.stream(side):
                self._collect()
            torch.cuda.current_stream(self.env.device).wait_stream(side)
            torch.cuda.synchronize(self.env.device)
            self.env.stats(reset=True)               # the warm-up pass stepped the env; its episodes are not part of a segment
            self._graph_version = (getattr(self.env, "config_version", 0), self.policy.tf32)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._seg = self._collect()
        self.noise_buf.normal_(generator=self.gen)
        self._graph.replay()
        return self._seg

    @torch.no_grad()
    def _collect(self):
        env, pol, T = self.env, self.policy, self.T
        self.nextv_buf.zero_()
        for t in range(T):
            if self.use_graph:
                action, value, _ = pol.step(self.obs, noise=self.noise_buf[t])
            else:
                action, value, _ = pol.step(self.obs, generator=self.gen)
            self.obs_buf[t] = self.obs
            self.vpred_buf[t] = value
            self.act_buf[t] = action                                   # the runner stores the UNclipped action (:132)
            self.start_buf[t] = self.episode_start
            if t > 0:                                                  # imitation_runners.py:134-136
                keep = self.episode_start == 0
                self.nextv_buf[t - 1] = torch.where(keep, value, self.nextv_buf[t - 1])
            clipped = torch.clamp(action, self.low, self.high).contiguous()          # :138-143
            obs, rew, done = env.step(clipped)
            self.rew_buf[t] = rew
            self.done_buf[t] = done
            self.episode_start.copy_(done)
            if self.reset_all:
                if bool(done.any()):                                   # :185-205
                    self.nextv_buf[t].zero_()
                    self.obs.copy_(env.reset())
                else:
                    self.obs.copy_(obs)
            else:
                self.obs.copy_(obs)
        # segment end: the reference bootstraps the last row with 0.0 (imitation_runners.py:93-97)
        self.nextv_buf[T - 1].zero_()
        # The runner's index arithmetic (ppo_imitation.py:86-90 reads new[t*N + 2j + 1]) only cuts an episode correctly when
        # every robot restarts together (reset-all: the whole row of flags is equal).  With per-robot restarts inside the
        # step kernel each robot's recursion has to be cut at ITS OWN restart: flag of robot j in row t + 1.
        adv, ret = gae(self.rew_buf, self.vpred_buf, self.nextv_buf, self.start_buf, self.gamma, self.lam,
                       reference_index_quirk=self.reset_all, standardize=True)
        N = env.num_robot
        flat = lambda x: x.reshape(T * N, *x.shape[2:])              # time-major interleaved: index t*N + j
        return dict(observations=flat(self.obs_buf), actions=flat(self.act_buf), rewards=flat(self.rew_buf),
                    vpred=flat(self.vpred_buf), nextvpreds=flat(self.nextv_buf), dones=flat(self.done_buf),
                    episode_starts=flat(self.start_buf), adv=flat(adv), tdlamret=flat(ret), total_timestep=T * N)

    def reduced_stats(self):
        """[sum reward, sum episode length, episodes, robot steps] summed over ranks (one small all-reduce)."""
        return odist.allreduce_stats(self.env.stats(reset=True), device=self.env.device)
