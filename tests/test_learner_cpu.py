"""Row f3: PPO1 loss / gradient and MpiAdam of openroborl_b200/learner.py against the NumPy restatement (oracle/ppo.py)
and over a world_size-2 gloo group.  Runs on CPU tensors: the learner is library code around the environment."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ppo as oppo
from tests import common


def _params(seed=0, scale=0.05):
    rng = np.random.default_rng(seed)
    shp = {"pi_fc0": (160, 512), "pi_fc1": (512, 256), "pi": (256, 12), "vf_fc0": (160, 512), "vf_fc1": (512, 256), "vf": (256, 1)}
    p = {}
    for k, (a, b) in shp.items():
        p[k + "/w"] = (rng.normal(size=(a, b)) * scale).astype(np.float32)
        p[k + "/b"] = (rng.normal(size=b) * scale).astype(np.float32)
    return p


def _batch(n, seed=1):
    rng = np.random.default_rng(seed)
    obs = rng.normal(size=(n, 160)).astype(np.float32)
    act = rng.normal(scale=0.2, size=(n, 12)).astype(np.float32)
    atarg = rng.normal(size=n).astype(np.float32)
    ret = rng.normal(size=n).astype(np.float32)
    return obs, act, atarg, ret


def test_ppo_loss_and_gradient_match_the_restated_reference():
    from openroborl_b200.learner import PPOLearner, PARAM_ORDER
    from openroborl_b200.rollout import MlpGaussianPolicy
    assert PARAM_ORDER == oppo.PARAM_ORDER
    params, old = _params(0), _params(0)
    for k in old:                                  # an "old" policy a little away from the new one: ratios around 1 +- 0.3
        old[k] = old[k] + np.random.default_rng(5).normal(size=old[k].shape).astype(np.float32) * 2e-3
    pol = MlpGaussianPolicy(params, "cpu")
    lrn = PPOLearner(pol, clip_param=0.2, entcoeff=0.01)
    obs, act, atarg, ret = _batch(256)
    T = lambda a: torch.as_tensor(a)
    oldt = {k: T(v) for k, v in old.items()}
    ls = lrn.losses(oldt, T(obs), T(act), T(atarg), T(ret), 1.0)
    grad = lrn.flatgrad(ls[0] + ls[1] + ls[2]).numpy()
    want, g = oppo.ppo_losses_and_grad(params, old, obs, act, atarg, ret, 0.2, 0.01)
    got = [float(x.detach()) for x in ls]
    np.testing.assert_allclose(got, want, rtol=2e-4, atol=1e-6)
    flat = np.concatenate([g[k].reshape(-1) for k in oppo.PARAM_ORDER])
    assert np.abs(grad - flat).max() <= 1e-3 * np.abs(flat).max()
    # some ratios are clipped (the clipped branch has zero gradient), and the entropy term is a constant
    assert 0.0 < want[3] and abs(want[4] - 12 * (np.log(0.125) + 0.5 * np.log(2 * np.pi * np.e))) < 1e-12


def test_mpi_adam_update_matches_mpi_adam_py():
    from openroborl_b200.learner import MpiAdam
    rng = np.random.default_rng(0)
    ps = [torch.tensor(rng.normal(size=(7, 3)).astype(np.float32)), torch.tensor(rng.normal(size=5).astype(np.float32))]
    adam = MpiAdam(ps, epsilon=1e-5)
    theta = adam.getflat().numpy().copy()
    m, v, step = np.zeros_like(theta), np.zeros_like(theta), 0
    for it in range(5):
        g = rng.normal(size=theta.shape).astype(np.float32)
        adam.update(torch.tensor(g), 1e-2)
        theta, m, v, step = oppo.mpi_adam_update(theta, m, v, step, g, 1e-2, epsilon=1e-5)
        np.testing.assert_allclose(adam.getflat().numpy(), theta, rtol=2e-6, atol=1e-7)
    assert adam.step == 5


def test_update_improves_the_surrogate_and_keeps_unused_heads():
    from openroborl_b200.learner import PPOLearner
    from openroborl_b200.rollout import MlpGaussianPolicy
    pol = MlpGaussianPolicy(_params(3), "cpu")
    lrn = PPOLearner(pol, optim_epochs=4, optim_stepsize=3e-4, optim_batchsize=128, seed=1)
    obs, act, atarg, ret = _batch(1000, seed=7)            # 1000 samples, batch 128: 7 full batches, the remainder is dropped
    seg = dict(observations=torch.as_tensor(obs), actions=torch.as_tensor(act), adv=torch.as_tensor(atarg), tdlamret=torch.as_tensor(ret))
    before = {k: v.detach().clone() for k, v in pol.p.items()}
    with torch.no_grad():
        l0 = [float(x) for x in lrn.losses(before, seg["observations"], seg["actions"], seg["adv"], seg["tdlamret"], 1.0)]
    out = lrn.update(seg, 1.0)
    assert lrn.adam.step == 4 * 7
    with torch.no_grad():
        l1 = [float(x) for x in lrn.losses(before, seg["observations"], seg["actions"], seg["adv"], seg["tdlamret"], 1.0)]
    assert l1[0] + l1[2] < l0[0] + l0[2]                    # surrogate + value loss went down on the training data
    assert l1[3] > 0 and set(out) == {"pol_surr", "pol_entpen", "vf_loss", "kl", "ent"}
    assert any(not torch.equal(before[k], pol.p[k]) for k in before)


def test_saved_zip_round_trips_and_lists_the_reference_variables(tmp_path):
    from openroborl_b200.learner import PPOLearner
    from openroborl_b200.rollout import MlpGaussianPolicy, load_policy_zip
    import json, zipfile
    params = _params(2)
    lrn = PPOLearner(MlpGaussianPolicy(params, "cpu"))
    path = str(tmp_path / "model.zip")
    lrn.save(path)
    back = load_policy_zip(path)
    for k in params:
        assert np.array_equal(back[k], params[k])
    names = json.loads(zipfile.ZipFile(path).read("parameter_list"))
    ref = ['model/pi_fc0/w:0', 'model/pi_fc0/b:0', 'model/vf_fc0/w:0', 'model/vf_fc0/b:0', 'model/pi_fc1/w:0', 'model/pi_fc1/b:0',
           'model/vf_fc1/w:0', 'model/vf_fc1/b:0', 'model/vf/w:0', 'model/vf/b:0', 'model/pi/w:0', 'model/pi/b:0', 'model/q/w:0', 'model/q/b:0']
    assert names == ref                                      # parameter_list of the reference's own policy zips


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from openroborl_b200.learner import MpiAdam
        rng = np.random.default_rng(100 + rank)
        p = [torch.tensor(np.random.default_rng(rank).normal(size=(4, 3)).astype(np.float32))]     # ranks start apart
        adam = MpiAdam(p, epsilon=1e-5)
        adam.sync()                                                                                  # rank 0's weights everywhere
        start = adam.getflat().numpy().copy()
        grads = [rng.normal(size=12).astype(np.float32) for _ in range(3)]
        for g in grads:
            adam.update(torch.tensor(g), 1e-2)
        q.put((rank, start, [g for g in grads], adam.getflat().numpy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_mpi_adam_averages_gradients():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, s0, g0, e0), (_, s1, g1, e1) = res
    assert np.array_equal(s0, s1) and np.array_equal(e0, e1)          # synced start, identical end (check_synced would assert otherwise)
    theta, m, v, step = s0.copy(), np.zeros(12, np.float32), np.zeros(12, np.float32), 0
    for a, b in zip(g0, g1):
        theta, m, v, step = oppo.mpi_adam_update(theta, m, v, step, (a + b) / 2, 1e-2, epsilon=1e-5)
    np.testing.assert_allclose(e0, theta, rtol=2e-6, atol=1e-7)
