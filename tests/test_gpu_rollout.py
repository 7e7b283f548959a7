"""device policy inference, rollout buffers and GAE against the reference's formulas."""
import io
import json
import zipfile

import numpy as np
import pytest
import torch

from tests import common
from oracle import ppo as oppo

pytestmark = pytest.mark.gpu


def _params():
    z = np.load(common.GOLD + "/policy_laikago_pace.npz")
    return {k.rsplit("_", 1)[0] + "/" + k.rsplit("_", 1)[1]: z[k] for k in z.files if k.rsplit("_", 1)[-1] in ("w", "b")}


def test_gae_matches_add_vtarg_and_adv():
    from openroborl_b200.rollout import gae
    rng = np.random.default_rng(0)
    for T, N in ((16, 4), (32, 13), (5, 64)):
        rew = rng.uniform(0, 1, size=(T, N)).astype(np.float32)
        vp = rng.normal(size=(T, N)).astype(np.float32)
        nv = rng.normal(size=(T, N)).astype(np.float32)
        es = rng.random((T, N)) < 0.15
        dev = lambda a: torch.as_tensor(a, device="cuda")
        adv, ret = gae(dev(rew), dev(vp), dev(nv), dev(es.astype(np.uint8)), 0.95, 0.95, reference_index_quirk=True, standardize=False)
        a_ref, r_ref = oppo.add_vtarg_and_adv(rew.reshape(-1), vp.reshape(-1), nv.reshape(-1), es.reshape(-1), N, 0.95, 0.95)
        np.testing.assert_allclose(adv.cpu().numpy().reshape(-1), a_ref, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(ret.cpu().numpy().reshape(-1), r_ref, rtol=1e-5, atol=1e-5)
        adv_s, ret_s = gae(dev(rew), dev(vp), dev(nv), dev(es.astype(np.uint8)), 0.95, 0.95, True, True)
        np.testing.assert_allclose(adv_s.cpu().numpy().reshape(-1), oppo.standardize_per_robot(a_ref, N), rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(ret_s.cpu().numpy().reshape(-1), r_ref, rtol=1e-5, atol=1e-5)   # returns use the raw advantage
        # textbook indexing (episode_starts of the NEXT step of the SAME robot)
        adv_c, _ = gae(dev(rew), dev(vp), dev(nv), dev(es.astype(np.uint8)), 0.95, 0.95, False, False)
        last, want = np.zeros(N), np.zeros((T, N))
        for t in reversed(range(T)):
            nonterm = 1.0 - (es[t + 1] if t + 1 < T else np.zeros(N))
            last = rew[t] + 0.95 * nv[t] - vp[t] + 0.95 * 0.95 * nonterm * last
            want[t] = last
        np.testing.assert_allclose(adv_c.cpu().numpy(), want, rtol=1e-5, atol=1e-5)


def test_gae_kernel_against_the_reference_function():
    """orl_gae (reference index arithmetic) against outputs of the reference's own add_vtarg_and_adv
    (tests/golden/ref_ppo_kat.npz, scripts/make_ppo_golden.py)."""
    from openroborl_b200.rollout import gae
    kat = np.load(common.GOLD + "/ref_ppo_kat.npz")
    for c, (T, N) in enumerate(kat["gae/cases"]):
        T, N = int(T), int(N)
        g = lambda k: kat["gae/%d/%s" % (c, k)]
        dev = lambda a, dt=np.float32: torch.as_tensor(np.ascontiguousarray(a.reshape(T, N).astype(dt)), device="cuda")
        adv, ret = gae(dev(g("rewards")), dev(g("vpred")), dev(g("nextvpreds")), dev(g("episode_starts"), np.uint8), 0.95, 0.95,
                       reference_index_quirk=True, standardize=False)
        np.testing.assert_allclose(adv.cpu().numpy().reshape(-1), g("adv"), rtol=2e-6, atol=2e-6)
        np.testing.assert_allclose(ret.cpu().numpy().reshape(-1), g("tdlamret"), rtol=2e-6, atol=2e-6)


def test_policy_forward_and_zip_reader(tmp_path):
    from openroborl_b200.rollout import MlpGaussianPolicy, load_policy_zip
    params = _params()
    # write a stable-baselines style zip (data / parameters / parameter_list) and read it back
    names = ["model/%s:0" % k for k in params]
    buf = io.BytesIO()
    np.savez(buf, **{n: params[n[6:-2]] for n in names})
    zp = tmp_path / "model.zip"
    with zipfile.ZipFile(zp, "w") as z:
        z.writestr("data", json.dumps({"gamma": 0.95}))
        z.writestr("parameters", buf.getvalue())
        z.writestr("parameter_list", json.dumps(names))
    back = load_policy_zip(str(zp))
    assert set(back) == set(params) and all(np.array_equal(back[k], params[k]) for k in params)
    pol = MlpGaussianPolicy(back, "cuda")
    assert sum(p.numel() for p in pol.parameters()) == 433945 - 256 * 12 - 12            # SURVEY 2.1 count minus the unused q head
    rng = np.random.default_rng(1)
    obs = rng.normal(size=(257, 160)).astype(np.float32)
    a, v, nlp = pol.step(torch.as_tensor(obs, device="cuda"), deterministic=True)
    np.testing.assert_allclose(a.cpu().numpy(), oppo.mlp_forward(params, obs, "pi"), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(v.cpu().numpy(), oppo.mlp_forward(params, obs, "vf")[:, 0], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(nlp.cpu().numpy(), 0.5 * np.log(2 * np.pi) * 12 + np.log(0.125) * 12, rtol=1e-5)
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    a2, _, _ = pol.step(torch.as_tensor(obs, device="cuda"), generator=g)
    assert abs(float((a2 - a).std()) - 0.125) < 0.01                                      # fixed sigma (imitation_policies.py:48)


def test_rollout_collector_layout_and_semantics():
    from openroborl_b200.rollout import MlpGaussianPolicy, RolloutCollector
    N, T = 64, 12
    pol = MlpGaussianPolicy(_params(), "cuda")
    env = common.make_product("laikago", ["laikago_pace"], N, seed=2, max_episode_steps=600)
    col = RolloutCollector(env, pol, T, reset_all_on_any_done=True, seed=5)
    seg = col.collect()
    assert seg["observations"].shape == (T * N, 160) and seg["actions"].shape == (T * N, 12) and seg["total_timestep"] == T * N
    obs = seg["observations"].reshape(T, N, 160)
    vp, nv = seg["vpred"].reshape(T, N), seg["nextvpreds"].reshape(T, N)
    st, dn = seg["episode_starts"].reshape(T, N).bool(), seg["dones"].reshape(T, N).bool()
    assert st[0].all() and torch.equal(st[1:], dn[:-1])                                     # episode_start[j] = done[j] of the previous step
    assert torch.equal(nv[T - 1], torch.zeros_like(nv[T - 1]))                              # segment end bootstraps with 0
    for t in range(T - 1):
        keep = ~st[t + 1]
        assert torch.equal(nv[t][keep], vp[t + 1][keep])                                    # next value of a running episode
        if dn[t].any():
            assert torch.equal(nv[t][~keep], torch.zeros_like(nv[t][~keep]))               # cut at episode ends
    # whole-env reset on any done: the row after a done shows the stale default-pose histories for every robot
    if dn[:-1].any():
        t = int(torch.nonzero(dn[:-1].any(dim=1))[0])
        assert torch.equal(obs[t + 1][:, :12], torch.zeros_like(obs[t + 1][:, :12]))
    assert torch.isfinite(seg["adv"]).all() and abs(float(seg["adv"].reshape(T, N).mean(dim=0).abs().max())) < 1e-3
    # auto-reset mode: robots restart individually
    env2 = common.make_product("laikago", ["laikago_pace"], N, seed=2, auto_reset=True, max_episode_steps=5)
    seg2 = RolloutCollector(env2, pol, T, reset_all_on_any_done=False, seed=5).collect()
    assert seg2["dones"].reshape(T, N)[4].all() and torch.isfinite(seg2["tdlamret"]).all()
    stats = env2.stats()
    assert stats[3] == T * N and stats[2] >= 2 * N
    env.close(); env2.close()


def test_rollout_collector_cuda_graph_mode():
    """The captured T-step segment (policy forward + env step + bookkeeping + GAE) replays with the same semantics."""
    from openroborl_b200.rollout import MlpGaussianPolicy, RolloutCollector
    N, T = 96, 8
    pol = MlpGaussianPolicy(_params(), "cuda")
    env = common.make_product("laikago", ["laikago_pace"], N, seed=2, auto_reset=True, max_episode_steps=6)
    with pytest.raises(ValueError):
        RolloutCollector(env, pol, T, reset_all_on_any_done=True, use_cuda_graph=True)
    col = RolloutCollector(env, pol, T, reset_all_on_any_done=False, seed=5, use_cuda_graph=True)
    env.stats(reset=True)
    segs = []
    for it in range(3):                                  # capture + two replays
        seg = col.collect()
        segs.append({k: (v.clone() if torch.is_tensor(v) else v) for k, v in seg.items()})
    for seg in segs:
        vp, nv = seg["vpred"].reshape(T, N), seg["nextvpreds"].reshape(T, N)
        st, dn = seg["episode_starts"].reshape(T, N).bool(), seg["dones"].reshape(T, N).bool()
        assert torch.equal(st[1:], dn[:-1]) and torch.equal(nv[T - 1], torch.zeros_like(nv[T - 1]))
        for t in range(T - 1):
            keep = ~st[t + 1]
            assert torch.equal(nv[t][keep], vp[t + 1][keep])
        assert torch.isfinite(seg["adv"]).all() and torch.isfinite(seg["tdlamret"]).all()
        # GAE of the captured kernel on the segment's own buffers: with per-robot restarts every robot's recursion is cut
        # at ITS OWN restart (textbook GAE per robot), not at the runner's index (step*N+i)+(1+i), which only coincides
        # with it when all robots restart together
        rew, vpn, nvn = (seg[k].cpu().numpy().reshape(T, N).astype(np.float64) for k in ("rewards", "vpred", "nextvpreds"))
        es = seg["episode_starts"].cpu().numpy().reshape(T, N).astype(bool)
        last, want = np.zeros(N), np.zeros((T, N))
        for t in reversed(range(T)):
            nonterm = 1.0 - (es[t + 1] if t + 1 < T else np.zeros(N))
            last = rew[t] + 0.95 * nvn[t] - vpn[t] + 0.95 * 0.95 * nonterm * last
            want[t] = last
        np.testing.assert_allclose(seg["tdlamret"].cpu().numpy().reshape(T, N), want + vpn, rtol=1e-4, atol=1e-4)
        assert es[1:].any()
    # segments continue each other: every robot hits the 6-step limit in each 8-step segment, noise differs between replays
    total = sum(seg["dones"].reshape(T, N).sum(dim=0) for seg in segs)
    assert int(total.min()) >= 3
    assert not torch.equal(segs[1]["actions"], segs[2]["actions"])
    assert torch.equal(segs[1]["episode_starts"].reshape(T, N)[0], segs[0]["dones"].reshape(T, N)[T - 1])
    assert torch.equal(segs[2]["episode_starts"].reshape(T, N)[0], segs[1]["dones"].reshape(T, N)[T - 1])
    st = env.stats()
    assert st[3] == 3 * T * N                            # the three replays; the warm-up pass is dropped from the statistics, the capture does not run
    env.close()


def test_ppo_learner_end_to_end_on_the_cuda_env():
    """Row f3: rollout (CUDA-graph replay) -> PPO1 update -> MpiAdam, two iterations on a small batch."""
    from openroborl_b200.learner import PPOLearner
    from openroborl_b200.rollout import MlpGaussianPolicy
    N, T = 128, 16
    pol = MlpGaussianPolicy(_params(), "cuda")
    env = common.make_product("laikago", ["laikago_pace"], N, seed=3, auto_reset=True, max_episode_steps=600)
    lrn = PPOLearner(pol, optim_epochs=2, optim_stepsize=1e-4, optim_batchsize=512, seed=2)
    before = {k: v.detach().clone() for k, v in pol.p.items()}
    logs = []
    hist = lrn.learn(env, total_timesteps=2 * N * T, horizon=T, use_cuda_graph=True, seed=4, log=logs.append)
    assert len(hist) == 2 and hist == logs and lrn.num_timesteps == 2 * N * T
    assert lrn.adam.step == 2 * 2 * (N * T // 512)
    for rec in hist:
        assert all(np.isfinite(rec[k]) for k in ("pol_surr", "vf_loss", "kl", "mean_reward")) and rec["kl"] >= 0
    assert any(not torch.equal(before[k], pol.p[k]) for k in before)              # the policy moved
    assert hist[1]["vf_loss"] < hist[0]["vf_loss"] * 1.5                            # and did not blow up
    # the captured rollout graph reads the updated weights in place: a fresh forward agrees with the graph's last values
    seg_obs = torch.randn((8, 160), device="cuda")
    a, v, _ = pol.step(seg_obs, deterministic=True)
    np.testing.assert_allclose(v.cpu().numpy(), oppo.mlp_forward({k: t.detach().cpu().numpy() for k, t in pol.p.items()}, seg_obs.cpu().numpy(), "vf")[:, 0],
                               rtol=1e-3, atol=1e-3)
    env.close()
