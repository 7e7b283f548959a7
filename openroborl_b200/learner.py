"""PPO learner with MpiAdam semantics over torch.distributed.

Replaces `PPOImitation.setup_model/learn` (`OpenRoboRL/agents/ppo_imitation.py:156-440`) and `MpiAdam`
(`OpenRoboRL/stable_baselines/common/mpi_adam.py:8-83`) for the reference's fixed-sigma MLP actor-critic:
    total_loss = pol_surr + pol_entpen + vf_loss            (:176-208, PPO1)
    ratio = exp(logp_new(a) - logp_old(a)),  pol_surr = -mean(min(ratio A, clip(ratio, 1 -+ clip_param lrmult) A))
with the per-robot standardised GAE advantages of the rollout collector (rollout.py).  Every minibatch: flat gradient
(trainable-variable order of the reference's graph) -> all-reduce SUM / world size -> Adam with the bias-corrected step
size and epsilon OUTSIDE the square root (mpi_adam.py:50-63).  The dense layers and their backward pass are library
code (torch autograd / cuBLAS): the learner is a caller of the environment, not the hot path.  `save` writes a
stable-baselines 2.10 zip (`data`, `parameters`, `parameter_list`) that `run.py --mode test` can load.
"""
import io
import json
import math
import zipfile

import numpy as np
import torch
import torch.distributed as dist

from . import dist as odist
from .rollout import LOG_STD, MlpGaussianPolicy, RolloutCollector

# trainable variables in graph order (parameter_list of the reference's policy zips); q/* is created by the policy
# class but never reaches the loss, so its gradient is zero and Adam leaves it unchanged
PARAM_ORDER = ("pi_fc0/w", "pi_fc0/b", "vf_fc0/w", "vf_fc0/b", "pi_fc1/w", "pi_fc1/b", "vf_fc1/w", "vf_fc1/b",
               "vf/w", "vf/b", "pi/w", "pi/b")


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class MpiAdam(object):
    """mpi_adam.py:8-83 on a list of torch parameters (flat float32 state, one all-reduce per update)."""

    def __init__(self, params, beta1=0.9, beta2=0.999, epsilon=1e-8, scale_grad_by_procs=True):
        self.params = list(params)
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon
        self.scale_grad_by_procs = scale_grad_by_procs
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step = 0

    def getflat(self):
        return torch.cat([p.detach().reshape(-1) for p in self.params])

    def setfromflat(self, flat):
        off = 0
        with torch.no_grad():
            for p in self.params:
                n = p.numel()
                p.copy_(flat[off:off + n].view_as(p))
                off += n

    @torch.no_grad()
    def update(self, local_grad, learning_rate):
        if self.step % 100 == 0:
            self.check_synced()
        g = local_grad.to(torch.float32).clone()
        if _world() > 1:
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
            if self.scale_grad_by_procs:
                g /= _world()
        self.step += 1
        step_size = learning_rate * math.sqrt(1 - self.beta2 ** self.step) / (1 - self.beta1 ** self.step)
        self.exp_avg.mul_(self.beta1).add_(g, alpha=1 - self.beta1)
        self.exp_avg_sq.mul_(self.beta2).addcmul_(g, g, value=1 - self.beta2)
        step = (-step_size) * self.exp_avg / (self.exp_avg_sq.sqrt() + self.epsilon)
        self.setfromflat(self.getflat() + step)

    @torch.no_grad()
    def sync(self):
        theta = self.getflat()
        if _world() > 1:
            dist.broadcast(theta, src=0)
        self.setfromflat(theta)

    @torch.no_grad()
    def check_synced(self):
        if _world() == 1:
            return
        # the verdict is all-reduced (MIN), so that EVERY rank raises: a root that carried on alone would hang in its next collective
        if not odist.check_synced(self.params):
            raise AssertionError("MpiAdam: parameters diverged between ranks")


class PPOLearner(object):
    """PPO1 update of an MlpGaussianPolicy on segments produced by RolloutCollector."""

    def __init__(self, policy, clip_param=0.2, entcoeff=0.01, optim_epochs=1, optim_stepsize=1e-5, optim_batchsize=64,
                 adam_epsilon=1e-5, schedule="constant", seed=0):
        if schedule not in ("constant", "linear"):
            raise NotImplementedError(schedule)                       # ppo_imitation.py:277-282
        self.policy = policy
        self.clip_param, self.entcoeff = clip_param, entcoeff
        self.optim_epochs, self.optim_stepsize, self.optim_batchsize = optim_epochs, optim_stepsize, optim_batchsize
        self.schedule = schedule
        self.train_params = [policy.p[k] for k in PARAM_ORDER]
        for p in self.train_params:
            p.requires_grad_(True)
        self.adam = MpiAdam(self.train_params, epsilon=adam_epsilon)
        self.gen = torch.Generator(device=policy.device)
        self.gen.manual_seed(seed)
        self.loss_names = ["pol_surr", "pol_entpen", "vf_loss", "kl", "ent"]
        self.num_timesteps = 0

    def _forward(self, p, obs, head):
        h = torch.relu(torch.addmm(p[head + "_fc0/b"], obs, p[head + "_fc0/w"]))
        h = torch.relu(torch.addmm(p[head + "_fc1/b"], h, p[head + "_fc1/w"]))
        return torch.addmm(p[head + "/b"], h, p[head + "/w"])

    def losses(self, old, obs, actions, atarg, ret, lrmult):
        """[pol_surr, pol_entpen, vf_loss, meankl, meanent] (ppo_imitation.py:176-208) as 0-d tensors; first three sum to the loss."""
        p = self.policy.p
        clip = self.clip_param * lrmult
        inv2s2 = 0.5 / math.exp(2 * LOG_STD)
        mean = self._forward(p, obs, "pi")
        with torch.no_grad():
            mean_old = self._forward(old, obs, "pi")
        val = self._forward(p, obs, "vf")[:, 0]
        logp = -inv2s2 * ((actions - mean) ** 2).sum(dim=1)
        logp_old = -inv2s2 * ((actions - mean_old) ** 2).sum(dim=1)
        ratio = torch.exp(logp - logp_old)
        surr1 = ratio * atarg
        surr2 = torch.clamp(ratio, 1.0 - clip, 1.0 + clip) * atarg
        pol_surr = -torch.minimum(surr1, surr2).mean()
        ent = torch.tensor(mean.shape[1] * (LOG_STD + 0.5 * math.log(2 * math.pi * math.e)), device=obs.device)
        kl = (inv2s2 * ((mean_old - mean) ** 2).sum(dim=1)).mean()
        vf_loss = ((val - ret) ** 2).mean()
        return [pol_surr, -self.entcoeff * ent, vf_loss, kl, ent]

    def flatgrad(self, total):
        grads = torch.autograd.grad(total, self.train_params, allow_unused=True)
        return torch.cat([(g if g is not None else torch.zeros_like(p)).reshape(-1) for g, p in zip(grads, self.train_params)])

    def update(self, seg, lrmult=1.0):
        """One PPO iteration on a segment (ppo_imitation.py:340-395): optim_epochs passes over shuffled minibatches (only
        full batches, Dataset.iterate_once), then the losses on the whole segment.  Returns the mean losses."""
        obs, act = seg["observations"], seg["actions"]
        atarg, ret = seg["adv"], seg["tdlamret"]
        n = obs.shape[0]
        bs = self.optim_batchsize or n
        nb = (n - bs) // bs + 1 if n >= bs else 0        # full minibatches (Dataset.iterate_once)
        if _world() > 1:
            # every adam.update is a collective: ranks with uneven shards (robots % world != 0) must agree on the count
            t = torch.tensor([nb], dtype=torch.int64, device=obs.device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            nb = int(t.item())
        old = {k: v.detach().clone() for k, v in self.policy.p.items()}              # assign_old_eq_new
        for _ in range(self.optim_epochs):
            perm = torch.randperm(n, device=obs.device, generator=self.gen)
            for s in range(0, nb * bs, bs):
                idx = perm[s:s + bs]
                ls = self.losses(old, obs[idx], act[idx], atarg[idx], ret[idx], lrmult)
                grad = self.flatgrad(ls[0] + ls[1] + ls[2])
                self.adam.update(grad, self.optim_stepsize * lrmult)
        with torch.no_grad():
            tot = torch.zeros(5, dtype=torch.float64, device=obs.device)
            cnt = 0
            for s in range(0, nb * bs, bs):
                ls = self.losses(old, obs[s:s + bs], act[s:s + bs], atarg[s:s + bs], ret[s:s + bs], lrmult)
                tot += torch.stack([x.double() for x in ls]); cnt += 1
            mean = tot / max(cnt, 1)
            if _world() > 1:                                                           # mpi_moments over ranks
                dist.all_reduce(mean, op=dist.ReduceOp.SUM)
                mean /= _world()
        return dict(zip(self.loss_names, [float(x) for x in mean]))

    def learn(self, env, total_timesteps, horizon, gamma=0.95, lam=0.95, use_cuda_graph=True, seed=0, log=None):
        """PPOImitation.learn (:240-440) with an auto-reset env: adam.sync, then rollout -> update until the budget is spent."""
        self.adam.sync()
        col = RolloutCollector(env, self.policy, horizon, gamma=gamma, lam=lam, reset_all_on_any_done=False, seed=seed,
                               use_cuda_graph=use_cuda_graph)
        iters, history = 0, []
        while self.num_timesteps < total_timesteps:
            lrmult = 1.0 if self.schedule == "constant" else max(1.0 - float(self.num_timesteps) / total_timesteps, 0.0)
            seg = col.collect()
            losses = self.update(seg, lrmult)
            stats = col.reduced_stats()                      # [sum reward, sum episode length, episodes, robot steps] over ranks
            steps = torch.tensor([int(seg["total_timestep"])], dtype=torch.int64, device=self.policy.device)
            if _world() > 1:
                dist.all_reduce(steps, op=dist.ReduceOp.SUM)                     # MPI allreduce of seg["total_timestep"] (:424-426)
            self.num_timesteps += int(steps.item())
            rec = dict(iteration=iters, timesteps=self.num_timesteps, mean_reward=float(stats[0] / max(stats[3], 1)),
                       ep_len_mean=float(stats[1] / max(stats[2], 1)), episodes=int(stats[2]), **losses)
            history.append(rec)
            if log:
                log(rec)
            iters += 1
        return history

    def save(self, path, data=None):
        """stable-baselines 2.10 zip (base_class.py:552-590): `data` json, `parameters` npz, `parameter_list` json."""
        names = ["model/%s:0" % k for k in PARAM_ORDER] + ["model/q/w:0", "model/q/b:0"]
        arrays = {("model/%s:0" % k): self.policy.p[k].detach().cpu().numpy() for k in PARAM_ORDER}
        arrays["model/q/w:0"] = np.zeros((256, 12), np.float32); arrays["model/q/b:0"] = np.zeros(12, np.float32)
        buf = io.BytesIO()
        np.savez(buf, **arrays)
        meta = dict(gamma=0.95, lam=0.95, clip_param=self.clip_param, entcoeff=self.entcoeff, optim_epochs=self.optim_epochs,
                    optim_stepsize=self.optim_stepsize, optim_batchsize=self.optim_batchsize, adam_epsilon=self.adam.epsilon,
                    schedule=self.schedule, n_envs=1)
        meta.update(data or {})
        with zipfile.ZipFile(path, "w") as z:
            z.writestr("data", json.dumps(meta))
            z.writestr("parameters", buf.getvalue())
            z.writestr("parameter_list", json.dumps(names))
