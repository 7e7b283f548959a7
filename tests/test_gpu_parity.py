"""Parity of the CUDA path (through the C ABI) against the oracle and the reference's golden vectors.

Tolerances (north_star): motion-frame indices and termination flags bit-exact; FK and reward terms
<= 1e-5 relative; mass matrix / inverse dynamics <= 1e-4 relative.  Full control steps (33 FP32 substeps
with contact) are compared with the tolerances stated in each test.
"""
import numpy as np
import pytest
import torch

from tests import common

pytestmark = pytest.mark.gpu

KAT = np.load(common.GOLD + "/ref_motion_kat.npz")
PERM = [3, 4, 5, 0, 1, 2] + list(range(6, 18))      # interface [lin, ang, q] <- oracle [ang, lin, q]


ERROR_TABLE = {}


def dump_error_table():
    """CUDA-vs-oracle error distributions of test_control_step_matches_oracle, written where gpurun brings files back."""
    import json, os
    out = os.path.join(common.ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_errors.json"), "w") as f:
            json.dump(ERROR_TABLE, f, indent=1, sort_keys=True)
    except OSError:
        pass


@pytest.fixture(scope="module")
def laikago():
    env = common.make_product("laikago", ["laikago_pace"], 256, seed=5)
    orc = common.make_oracle("laikago", ["laikago_pace"], 256, seed=5, phys=common.oracle_phys_from_product(env.phys))
    yield env, orc
    env.close()


@pytest.fixture(scope="module")
def cheetah():
    env = common.make_product("mini_cheetah", ["minicheetah_trot"], 128, seed=9)
    orc = common.make_oracle("mini_cheetah", ["minicheetah_trot"], 128, seed=9, phys=common.oracle_phys_from_product(env.phys))
    yield env, orc
    env.close()


LIBRARY = ["laikago_pace", "laikago_trot", "laikago_spin", "laikago_hopturn", "laikago_sidesteps"]   # BASELINE.json configs[4]


@pytest.fixture(scope="module")
def library():
    """Config C5: the 5-clip Laikago motion library, random clip and random reference start per episode."""
    env = common.make_product("laikago", LIBRARY, 320, seed=13)
    orc = common.make_oracle("laikago", LIBRARY, 320, seed=13, phys=common.oracle_phys_from_product(env.phys))
    yield env, orc
    env.close()


# ---- (a1-a4) motion tables and frame lookup ---------------------------------------------------------
@pytest.mark.parametrize("name", ["laikago_pace", "laikago_spin", "laikago_hopturn", "minicheetah_trot", "synthetic_pace_clamp"])
def test_motion_lookup_against_reference_golden(name):
    robot = "mini_cheetah" if name.startswith("mini") else "laikago"
    env = common.make_product(robot, [name], 8)
    mt = env.motion_tables()
    np.testing.assert_allclose(mt["frames"], KAT[name + "/frames"], rtol=0, atol=3e-16)    # host post-processing (f64)
    np.testing.assert_allclose(mt["vels"], KAT[name + "/vels"], rtol=0, atol=1e-11)
    assert mt["duration"][0] == float(KAT[name + "/duration"])
    np.testing.assert_allclose(mt["delta_heading"][0], KAT[name + "/delta_heading"], atol=1e-14)
    t = KAT[name + "/times"]
    lk = env.motion_lookup(np.zeros(len(t), dtype=np.int32), t)
    assert np.array_equal(lk["f0"], KAT[name + "/f0"]) and np.array_equal(lk["f1"], KAT[name + "/f1"])   # bit exact
    assert np.array_equal(lk["cycle"], KAT[name + "/cycle"])
    assert np.array_equal(lk["blend"], KAT[name + "/blend"])                                         # FP64, same op order
    scale = np.abs(KAT[name + "/pose"]).max()
    assert np.abs(lk["pose"] - KAT[name + "/pose"]).max() <= 1e-6 * max(scale, 1.0)
    assert np.abs(lk["vel"] - KAT[name + "/vel"]).max() <= 1e-6 * np.abs(KAT[name + "/vel"]).max()
    env.close()


def test_clamp_mode_clip_terminates_when_the_motion_is_over():
    """a10: done = motion_over for Clamp clips (imitation_task.py:532, motion_data.py:265-276); never for Wrap clips."""
    N = 32
    env = common.make_product("laikago", ["synthetic_pace_clamp"], N, seed=4)
    orc = common.make_oracle("laikago", ["synthetic_pace_clamp"], N, seed=4)
    env.reset(); orc.reset()
    dur = float(KAT["synthetic_pace_clamp/duration"])
    rng = np.random.default_rng(0)
    over_seen = False
    for step in range(int(dur / 0.033) + 4):
        common.sync_oracle_from_product(orc, env.get_state())
        act = rng.normal(scale=0.05, size=(N, 12)).astype(np.float32)
        o, r, d = env.step(torch.as_tensor(act, device=env.device))
        oo, ro, do = orc.step(act.astype(np.float64))
        st = env.get_state()
        t_motion = st["substep_counter"] * 1e-3 + st["time_offset"] - np.where(st["warmup"] != 0, 0.25, 0.0)
        over = t_motion >= dur
        d = d.cpu().numpy().astype(bool)
        assert (d | ~over).all()                              # motion over => done
        safe = (orc._term_margin > 1e-4) & (orc._fall_margin > 1e-5)
        assert np.array_equal(d[safe], do[safe])
        over_seen |= bool(over.any())
    assert over_seen and over.mean() > 0.8 and d[over].all()  # all but the warm-up robots (negative start time) are past the end
    env.close()


def test_motion_lookup_library_random_times():
    names = ["laikago_pace", "laikago_trot", "laikago_spin", "laikago_hopturn", "laikago_sidesteps"]
    env = common.make_product("laikago", names, 8)
    orc = common.make_oracle("laikago", names, 8)
    rng = np.random.default_rng(3)
    clip = rng.integers(0, 5, size=20000)
    t = rng.uniform(-0.3, 30.0, size=20000)
    lk = env.motion_lookup(clip, t)
    f0, f1, bl = orc.lib.calc_blend_idx(clip, t)
    assert np.array_equal(lk["f0"], f0) and np.array_equal(lk["f1"], f1) and np.array_equal(lk["blend"], bl)
    assert np.array_equal(lk["cycle"], orc.lib.calc_cycle_count(clip, t))
    pose = orc.lib.calc_frame(clip, t)
    assert np.abs(lk["pose"] - pose).max() <= 1e-6 * np.abs(pose).max()
    env.close()


# ---- (a8, a14) kinematics and dynamics ------------------------------------------------------------------
def test_cuda_fk_against_the_reference_closed_form_leg_model(cheetah):
    """a8 against the one piece of forward kinematics the reference states itself: the analytic 3-link leg of
    R/envs/quadruped_robot/task/motions/trans2minicheetah.m:28-30, 146-161 (the model its mini-cheetah clip was made with).
    Toe position from orl_fk relative to the hip joint, random base poses and joint angles, <= 1e-5 relative."""
    from openroborl_b200 import retarget as rt
    env, orc = cheetah
    rng = np.random.default_rng(7)
    n = 256
    pose, _ = common.random_states(rng, n, "mini_cheetah")
    pose[:, 7:] = rng.uniform(-1.0, 1.0, size=(n, 12)).astype(np.float32)
    p64 = pose.astype(np.float64)
    com = env.fk(pose)                                                  # world COM of every link (getLinkState()[0])
    k = orc.tree.fk(p64[:, 0:3], p64[:, 3:7], p64[:, 7:])               # hip joint origins (fixed in the base) from the URDF tree
    from oracle import xform as xf
    Rb = xf.quat_to_matrix(p64[:, 3:7])
    for l, side in enumerate((-1.0, 1.0, -1.0, 1.0)):                   # fr, fl, hr, hl
        hip, toe = 1 + 4 * l, 4 + 4 * l
        rel_world = com[:, toe - 1] - k["x"][:, hip]                    # the toe link's inertial origin is its own origin in the fixture
        rel = np.einsum("nji,nj->ni", Rb, rel_world)                    # into the base frame
        want = rt.leg_fk(p64[:, 7 + 3 * l:10 + 3 * l] * rt.MINI_SIGNS[3 * l:3 * l + 3], *rt.MINI_CHEETAH_LEG, side)
        assert np.abs(rel - want).max() <= 1e-5 * np.abs(want).max() + 2e-6      # float32 world coordinates up to 3 m from the origin



@pytest.mark.parametrize("which", ["laikago", "cheetah"])
def test_fk_mass_matrix_inverse_dynamics(which, request):
    env, orc = request.getfixturevalue(which)
    rng = np.random.default_rng(1)
    n = 200
    pose, vel = common.random_states(rng, n)
    p64, v64 = pose.astype(np.float64), vel.astype(np.float64)
    com = orc.tree.link_com(p64[:, 0:3], p64[:, 3:7], p64[:, 7:])
    rel = (com - p64[:, None, 0:3])                                  # positions relative to the base
    got = env.fk(pose) - pose[:, None, 0:3]
    assert np.abs(got - rel).max() <= 1e-5 * np.abs(rel).max()       # FK <= 1e-5 relative
    M_o = orc.tree.mass_matrix(p64[:, 0:3], p64[:, 3:7], p64[:, 7:])[:, PERM][:, :, PERM]
    M = env.mass_matrix(pose)
    assert np.abs(M - M_o).max() <= 1e-4 * np.abs(M_o).max()         # <= 1e-4 relative (measured ~1e-7)
    assert np.abs(M - M.transpose(0, 2, 1)).max() == 0.0
    acc = (rng.normal(size=(n, 18)) * 5).astype(np.float32)
    nu = np.concatenate([v64[:, 3:6], v64[:, 0:3], v64[:, 6:]], axis=1)
    nud = np.concatenate([acc[:, 3:6], acc[:, 0:3], acc[:, 6:]], axis=1).astype(np.float64)
    tau_o = orc.tree.inverse_dynamics(p64[:, 0:3], p64[:, 3:7], p64[:, 7:], nu, nud)[:, PERM]
    tau = env.inverse_dynamics(pose, vel, acc)
    assert np.abs(tau - tau_o).max() <= 1e-4 * np.abs(tau_o).max()
    # linearity in the acceleration: ID(a1 + a2) - ID(0) == (ID(a1) - ID(0)) + (ID(a2) - ID(0))
    z = np.zeros_like(acc)
    t0, t1, t2 = env.inverse_dynamics(pose, vel, z), env.inverse_dynamics(pose, vel, acc), env.inverse_dynamics(pose, vel, -acc)
    assert np.abs((t1 - t0) + (t2 - t0)).max() <= 2e-3 * np.abs(t1).max()


@pytest.mark.parametrize("which", ["laikago", "cheetah"])
def test_forward_dynamics_with_contact(which, request):
    env, orc = request.getfixturevalue(which)
    rng = np.random.default_rng(2)
    n = orc.N
    pose, vel = common.random_states(rng, n)
    pose[:, 2] = rng.uniform(0.15, 0.5, size=n)
    pose[:, 3:7] = np.array(env.model.spec.INIT_QUAT) + rng.normal(size=(n, 4)) * 0.1
    pose[:, 3:7] /= np.linalg.norm(pose[:, 3:7], axis=1, keepdims=True)
    tau = (rng.normal(size=(n, 12)) * 5).astype(np.float32)
    orc.pos, orc.quat, orc.q = [pose[:, a:b].astype(np.float64) for a, b in ((0, 3), (3, 7), (7, 19))]
    orc.lin, orc.ang, orc.qd = [vel[:, a:b].astype(np.float64) for a, b in ((0, 3), (3, 6), (6, 18))]
    acc_o = orc.forward_dynamics(tau.astype(np.float64))[:, PERM]
    acc = env.forward_dynamics(pose, vel, tau)
    # stiff contact (kn = 1e5) amplifies FP32 rounding of the penetration depth: 1e-4 of the batch scale
    assert np.abs(acc - acc_o).max() <= 1e-4 * np.abs(acc_o).max()
    # free fall: no contact, no torque, no velocity -> the base accelerates with gravity, bit for bit finite
    pose[:, 2] = 5.0
    ff = env.forward_dynamics(pose, np.zeros_like(vel), np.zeros_like(tau))
    np.testing.assert_allclose(ff[:, 0:3], np.tile([0, 0, -10.0], (n, 1)), atol=2e-4)


# ---- (a5-a7, a16) reset --------------------------------------------------------------------------------
@pytest.mark.parametrize("which", ["laikago", "cheetah", "library"])
def test_reset_matches_oracle(which, request):
    env, orc = request.getfixturevalue(which)
    obs = env.reset().cpu().numpy()
    obs_o = orc.reset()
    st = env.get_state()
    assert np.array_equal(st["clip"], orc.clip) and np.array_equal(st["warmup"].astype(bool), orc.warmup)
    if which == "library":                                             # every clip of the library is drawn, incl. laikago_spin
        assert set(np.unique(st["clip"])) == set(range(len(LIBRARY)))
    assert np.array_equal(st["time_offset"], orc.time_offset) and np.array_equal(st["prev_phase"], orc.prev_phase)   # FP64 exact
    assert orc.warmup.any() and (~orc.warmup).any()                    # both RSI and warm-up episodes are covered
    for k in ("pos", "quat", "q", "lin", "ang", "qd", "ref_pose", "ref_vel", "origin_pos", "origin_rot"):
        assert np.abs(st[k] - getattr(orc, k)).max() <= 2e-6 * max(1.0, np.abs(getattr(orc, k)).max()), k
    assert np.abs(obs - obs_o).max() <= 2e-6 * np.abs(obs_o).max()
    # quirk A.6.1: the first observation shows the DEFAULT pose in the sensor histories
    assert np.array_equal(obs[:, 0:12], np.zeros_like(obs[:, 0:12])) and np.array_equal(obs[:, 12:48], np.zeros_like(obs[:, 12:48]))
    # perfect tracking right after reset: reward terms all 1 (imitation_task.py:350-356)
    assert np.abs(env.reward_terms() - 1.0).max() <= 1e-5


def test_masked_reset_only_touches_masked_robots(laikago):
    env, _ = laikago
    env.reset()
    before = env.get_state()
    mask = torch.zeros(env.num_robot, dtype=torch.uint8, device=env.device)
    mask[::3] = 1
    env.reset(mask)
    after = env.get_state()
    m = mask.cpu().numpy().astype(bool)
    for k in before:
        assert np.array_equal(before[k][~m], after[k][~m]), k
    assert np.array_equal(after["episode"][m], before["episode"][m] + 1)
    assert not np.array_equal(after["time_offset"][m], before["time_offset"][m])


# ---- (a9-a13, a15) full control steps ------------------------------------------------------------------
@pytest.mark.parametrize("which", ["laikago", "cheetah", "library"])
def test_control_step_matches_oracle(which, request):
    env, orc = request.getfixturevalue(which)
    env.reset()
    orc.reset()
    rng = np.random.default_rng(4)
    for step in range(4):
        common.sync_oracle_from_product(orc, env.get_state())
        # reward terms on IDENTICAL states (no physics in between): <= 1e-5 relative
        rt, rt_o = env.reward_terms(), orc.reward_terms()
        assert np.abs(rt - rt_o).max() <= 1e-5 * np.abs(rt_o).max() + 1e-7
        act = rng.normal(scale=0.2, size=(env.num_robot, 12)).astype(np.float32)
        o, r, d = env.step(torch.as_tensor(act, device=env.device))
        oo, ro, do = orc.step(act.astype(np.float64))
        o, r, d = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy().astype(bool)
        st = env.get_state()
        # termination flags: bit exact wherever the oracle is not within 1e-4 of a threshold
        safe = (orc._term_margin > 1e-4) & (orc._fall_margin > 1e-5)
        assert safe.mean() > 0.95 and np.array_equal(d[safe], do[safe])
        assert np.array_equal(st["prev_phase"], orc.prev_phase)                      # FP64 clock, exact
        assert np.array_equal(st["substep_counter"], orc.substep_counter)
        # One control step = 33 FP32 substeps with stiff contact against the FP64 oracle from the SAME start state.  The
        # error distribution per field is recorded (profiles/r02_parity_errors.json, table in DESIGN.md section 6) and the
        # bounds below are the measured ones with a margin: medians at FP32 round-off, 99th percentiles two orders above, the
        # maximum only over robots without a contact that toggles on a knife edge (their joint velocities differ by > 2e-2).
        ok = np.abs(st["qd"] - orc.qd).max(axis=1) < 2e-2
        assert ok.mean() > 0.97
        errs = {"reward": np.abs(r - ro), "obs": np.abs(o - oo).max(axis=1)}
        for k in ("pos", "quat", "q", "lin", "ang", "qd"):
            errs[k] = np.abs(st[k] - getattr(orc, k)).max(axis=1)
        rec = {k: dict(p50=float(np.percentile(v, 50)), p99=float(np.percentile(v, 99)), max_ok=float(v[ok].max()), max=float(v.max()))
               for k, v in errs.items()}
        ERROR_TABLE.setdefault(which, {})["step %d" % (step + 1)] = dict(rec, robots=int(env.num_robot), knife_edge_robots=int((~ok).sum()))
        # (p50, p99, max over robots without a knife-edge contact); measured r02: reward 3e-7 / 4e-6 / 4e-5, obs 2e-5 / 2e-4 / 4e-3,
        # pos 4e-8 / 2e-7 / 4e-6, q 1e-6 / 1e-5 / 2e-4, qd 1e-4 / 1e-3 / 1e-2 (library clips incl. laikago_spin are the worst)
        bound = {"reward": (1e-6, 1e-5, 2e-4), "obs": (5e-5, 5e-4, 2e-2), "pos": (2e-7, 1e-6, 2e-5), "quat": (5e-7, 3e-6, 3e-5),
                 "q": (5e-6, 5e-5, 1e-3), "lin": (1e-5, 3e-5, 1e-3), "ang": (1e-4, 1e-3, 2e-2), "qd": (5e-4, 5e-3, 2e-2)}
        for k, (b50, b99, bmax) in bound.items():
            assert rec[k]["p50"] <= b50 and rec[k]["p99"] <= b99 and rec[k]["max_ok"] <= bmax, (k, rec[k])
        for k in ("act_hist", "filt_x", "filt_y", "prev_filtered", "ref_vel"):
            assert np.abs(st[k] - getattr(orc, k)).max() <= 2e-6 * max(1.0, np.abs(getattr(orc, k)).max()), k
    dump_error_table()


def test_velocity_clamp_matches_oracle_and_state_stays_finite(laikago):
    """The reference's own velocity tables spike (laikago_spin: |w| = 375 rad/s where the clip's quaternions flip
    sign), so reference-state initialisation can start a robot with such a velocity.  Every generalised velocity
    coordinate is clamped to +-100 (btMultiBody) in the oracle and in the kernel."""
    env, orc = laikago
    env.reset(); orc.reset()
    st = env.get_state()
    N = env.num_robot
    st["ang"] = np.tile(np.array([-83.8, -167.7, -255.4], np.float32), (N, 1))
    st["qd"] = np.full((N, 12), 150.0, np.float32) * np.sign(np.random.default_rng(0).normal(size=(N, 12))).astype(np.float32)
    env.set_state(st)
    common.sync_oracle_from_product(orc, env.get_state())
    act = np.zeros((N, 12), np.float32)
    o, r, d = env.step(torch.as_tensor(act, device=env.device))
    orc.step(act.astype(np.float64))
    s2 = env.get_state()
    for k in ("pos", "quat", "q", "lin", "ang", "qd"):
        assert np.isfinite(s2[k]).all(), k
    for k in ("lin", "ang", "qd"):
        assert np.abs(s2[k]).max() <= 100.0 + 1e-3 and np.abs(getattr(orc, k)).max() <= 100.0 + 1e-9
    assert np.isfinite(o.cpu().numpy()).all() and np.isfinite(r.cpu().numpy()).all()
    # violent, chaotic motion: FP32 and FP64 trajectories only agree in the bulk
    assert np.median(np.abs(s2["ang"] - orc.ang)) < 1.0 and np.median(np.abs(s2["pos"] - orc.pos)) < 1e-2


def test_motion_library_with_velocity_spikes_stays_finite():
    clips = ["laikago_pace", "laikago_trot", "laikago_spin", "laikago_hopturn", "laikago_sidesteps"]
    N = 4096
    env = common.make_product("laikago", clips, N, seed=0, auto_reset=True)
    env.reset()
    assert np.abs(env.motion_tables()["vels"][:, 3:6]).max() > 300.0       # the spike is in the reference's data
    gen = torch.Generator(device=env.device); gen.manual_seed(0)
    acts = [torch.randn((N, 12), device=env.device, generator=gen) * 0.125 for _ in range(4)]
    finite = torch.ones((), dtype=torch.bool, device=env.device)
    for i in range(150):
        o, r, d = env.step(acts[i % 4])
        finite &= torch.isfinite(o).all() & torch.isfinite(r).all()
    assert bool(finite)
    st = env.stats()
    assert np.isfinite(st).all() and st[2] > 0
    s = env.get_state()
    assert np.isfinite(s["pos"]).all() and np.isfinite(s["qd"]).all()
    env.close()


def test_non_finite_state_ends_the_episode():
    N = 16
    env = common.make_product("laikago", ["laikago_pace"], N, seed=1)
    env.reset()
    st = env.get_state()
    st["pos"][3, 2] = np.nan
    env.set_state(st)
    o, r, d = env.step(torch.zeros((N, 12), device=env.device))
    d, r = d.cpu().numpy(), r.cpu().numpy()
    assert d[3] == 1 and r[3] == 0.0 and d.sum() == 1 and np.isfinite(r).all()
    env.close()


# ---- (f2) domain randomizer ----------------------------------------------------------------------------------
@pytest.mark.parametrize("robot,clip", [("laikago", "laikago_pace"), ("mini_cheetah", "minicheetah_trot")])
def test_randomizer_reset_values_match_oracle(robot, clip):
    """Same Philox draws -> same sampled parameters (reference arithmetic: oracle/randomizer.py, pinned to the
    reference class by tests/golden/ref_randomizer_kat.npz)."""
    from oracle import randomizer as orand
    N = 256
    env = common.make_product(robot, [clip], N, seed=11, randomizer=True)
    orc = common.make_oracle(robot, [clip], N, seed=11, randomizer=True)
    for episode in range(2):
        env.reset(); orc.reset()
        st = env.get_state()
        rp = st["rand_params"].astype(np.float64)
        v = orc.rand_values
        tol = dict(rtol=2e-7, atol=1e-9)
        assert np.allclose(rp[:, 0], v["mass"][:, 0], **tol) and np.allclose(rp[:, 1], v["mass"][:, 1], **tol)
        assert np.allclose(rp[:, 2], v["inertia"][:, 0], **tol) and np.allclose(rp[:, 3], v["inertia"][:, 1], **tol)
        assert np.allclose(rp[:, 4], v["latency"][:, 0], **tol)
        assert np.allclose(rp[:, 5], v["lateral friction"][:, 0] * env.phys["mu_plane"], **tol)
        assert np.allclose(rp[:, 6:10], v["joint friction"][:, 0:8:2], **tol)          # knee of leg l <- sample 2l
        assert np.allclose(st["motor_strength"], v["motor strength"], **tol)
        # the record maps back onto the oracle's physical parameters
        o2 = common.make_oracle(robot, [clip], N, seed=11)
        common.load_rand_params(o2, rp)
        assert np.allclose(o2.link_mass, orc.link_mass, rtol=3e-7) and np.allclose(o2.inertia_scale, orc.inertia_scale, rtol=3e-7)
        assert np.allclose(o2.joint_friction, orc.joint_friction, rtol=3e-7, atol=1e-9)
        assert np.allclose(o2.foot_friction, orc.foot_friction, rtol=3e-7)
        for name, (lo, hi) in orand.ALL_PARAMS.items():
            assert v[name].min() >= lo and v[name].max() <= hi
    # a second episode draws new values; a disabled randomizer leaves the defaults
    off = common.make_product(robot, [clip], 8, seed=11)
    off.reset()
    r0 = off.get_state()["rand_params"]
    assert np.array_equal(r0[:, :4], np.ones((8, 4), np.float32)) and (r0[:, 5] == -1).all() and (r0[:, 6:10] == 0).all()
    env.close(); off.close()


def test_randomized_dynamics_match_oracle():
    """Mass matrix / inverse dynamics / forward dynamics with per-robot masses, inertias, foot friction."""
    N = 128
    env = common.make_product("laikago", ["laikago_pace"], N, seed=3, randomizer=True)
    orc = common.make_oracle("laikago", ["laikago_pace"], N, seed=3, randomizer=True)
    env.reset(); orc.reset()
    st = env.get_state()
    common.sync_oracle_from_product(orc, st)
    rng = np.random.default_rng(0)
    pose = np.concatenate([st["pos"], st["quat"], st["q"]], axis=1)
    pose[:, 2] -= rng.uniform(0.0, 0.004, size=N).astype(np.float32)         # push the feet into the ground
    vel = rng.normal(scale=0.5, size=(N, 18)).astype(np.float32)
    env.set_dynamics_rand(st["rand_params"])
    M = env.mass_matrix(pose)
    t = orc.tree
    Mo = t.mass_matrix(pose[:, 0:3].astype(np.float64), pose[:, 3:7].astype(np.float64), pose[:, 7:].astype(np.float64),
                       mass=orc.link_mass, inertia_scale=orc.inertia_scale)
    Mo = Mo[:, PERM][:, :, PERM]                                              # oracle order [ang, lin, q] -> interface [lin, ang, q]
    assert np.abs(M - Mo).max() <= 1e-4 * np.abs(Mo).max()
    leg = np.abs(Mo[:, 6:, 6:]).max()
    assert np.abs(M - Mo)[:, 6:, 6:].max() <= 1e-4 * leg                      # the small leg blocks on their own scale
    # and it is NOT the un-randomised matrix
    env.set_dynamics_rand(None)
    assert np.abs(env.mass_matrix(pose) - Mo).max() > 1e-2 * np.abs(Mo).max()
    env.set_dynamics_rand(st["rand_params"])
    # forward dynamics with contact
    tau = rng.normal(scale=3.0, size=(N, 12)).astype(np.float32)
    acc = env.forward_dynamics(pose, vel, tau)
    orc.pos, orc.quat, orc.q = [pose[:, a:b].astype(np.float64) for a, b in ((0, 3), (3, 7), (7, 19))]
    orc.lin, orc.ang, orc.qd = [vel[:, a:b].astype(np.float64) for a, b in ((0, 3), (3, 6), (6, 18))]
    acc_o = orc.forward_dynamics(tau.astype(np.float64))[:, PERM]
    assert np.abs(acc - acc_o).max() <= 1e-4 * np.abs(acc_o).max()          # same criterion as the nominal test
    err = np.abs(acc - acc_o).max(axis=1) / (np.abs(acc_o).max(axis=1) + 1.0)
    assert np.median(err) < 2e-5
    env.set_dynamics_rand(None)
    env.close()


def test_randomized_control_step_matches_oracle():
    N = 128
    env = common.make_product("laikago", ["laikago_pace"], N, seed=5, randomizer=True)
    orc = common.make_oracle("laikago", ["laikago_pace"], N, seed=5, randomizer=True)
    env.reset(); orc.reset()
    rng = np.random.default_rng(2)
    for step in range(3):
        common.sync_oracle_from_product(orc, env.get_state(), rings=env.latency_rings())
        act = rng.normal(scale=0.2, size=(N, 12)).astype(np.float32)
        o, r, d = env.step(torch.as_tensor(act, device=env.device))
        oo, ro, do = orc.step(act.astype(np.float64))
        o, r = o.cpu().numpy(), r.cpu().numpy()
        st = env.get_state()
        ok = np.abs(st["qd"] - orc.qd).max(axis=1) < 2e-2
        assert ok.mean() > 0.95
        assert np.abs(r - ro)[ok].max() <= 1e-4 and np.median(np.abs(r - ro)) <= 5e-6
        assert np.abs(st["q"] - orc.q)[ok].max() <= 2e-4 and np.abs(st["pos"] - orc.pos)[ok].max() <= 1e-4
        assert np.abs(o - oo)[ok].max() <= 2e-2 and np.median(np.abs(o - oo)) <= 1e-5     # incl. the delayed sensor readings
    # the randomised robots really move differently from the nominal ones
    nominal = common.make_product("laikago", ["laikago_pace"], N, seed=5)
    nominal.reset()
    s0 = env.get_state()
    keep = {k: s0[k] for k in ("pos", "quat", "q", "lin", "ang", "qd")}
    nominal.set_state(dict(keep)); env.set_state(dict(keep))
    z = torch.zeros((N, 12), device=env.device)
    nominal.step(z); env.step(z)
    assert np.abs(nominal.get_state()["q"] - env.get_state()["q"]).max() > 1e-3
    env.close(); nominal.close()


@pytest.mark.parametrize("latency", [0.0, 0.0003, 0.001, 0.002, 0.0137, 0.04])
def test_randomised_control_latency_matches_oracle(latency):
    """_get_delay_obs (minitaur.py:336-357) with the latency the randomizer sets: 0, a fraction of a substep, whole
    substeps, the default, 13.7 substeps, and the maximum (40 substeps: longer than one control step, and longer than
    the history right after a reset, where the oldest entry is returned)."""
    N = 64
    env = common.make_product("laikago", ["laikago_pace"], N, seed=9, randomizer=True)
    orc = common.make_oracle("laikago", ["laikago_pace"], N, seed=9, randomizer=True)
    env.reset(); orc.reset()
    st = env.get_state()
    rp = st["rand_params"].copy(); rp[:, 4] = latency
    env.set_state(dict(rand_params=rp))
    # the control observation of the FIRST step was fixed at reset with the sampled latency: redo it for the forced one
    rq, rb = env.latency_rings()
    n, al = int(np.float32(latency) / 1e-3), np.float32(latency) / 1e-3 - int(np.float32(latency) / 1e-3)
    first = (1 - al) * rq[:, 1] + al * rq[:, 0] if n == 0 else rq[:, 0]
    env.set_state(dict(ring2=first.astype(np.float32)))
    rng = np.random.default_rng(1)
    for step in range(4):                          # 4 x 33 substeps: the 40-substep latency reaches past a whole step
        common.sync_oracle_from_product(orc, env.get_state(), rings=env.latency_rings())
        assert np.allclose(orc.ctrl_latency, np.float32(latency))
        act = rng.normal(scale=0.1, size=(N, 12)).astype(np.float32)
        o, r, d = env.step(torch.as_tensor(act, device=env.device))
        oo, ro, do = orc.step(act.astype(np.float64))
        o = o.cpu().numpy()
        st = env.get_state()
        ok = np.abs(st["qd"] - orc.qd).max(axis=1) < 2e-2
        assert ok.mean() > 0.9
        # delayed motor angles and IMU channels of the newest history row, then the whole observation
        assert np.abs(st["ang_hist"][:, 0] - orc.ang_hist[:, 0])[ok].max() <= 2e-4
        assert np.abs(st["imu_hist"][:, 0] - orc.imu_hist[:, 0])[ok].max() <= 5e-3
        assert np.median(np.abs(o - oo)) <= 1e-5
        assert np.abs(st["q"] - orc.q)[ok].max() <= 2e-4
    env.close()


@pytest.mark.parametrize("rand", [False, True])
def test_checkpoint_restore_is_bit_exact_and_results_do_not_depend_on_sharding(rand):
    N = 48
    acts = [torch.as_tensor(np.random.default_rng(i).normal(scale=0.15, size=(N, 12)).astype(np.float32), device="cuda") for i in range(6)]
    a = common.make_product("laikago", ["laikago_pace"], N, seed=21, auto_reset=True, randomizer=rand, max_episode_steps=4)
    a.reset()
    for i in range(3):
        a.step(acts[i])
    ck = a.checkpoint()
    want = [tuple(x.cpu().numpy().copy() for x in a.step(acts[i])) for i in range(3, 6)]
    b = common.make_product("laikago", ["laikago_pace"], N, seed=21, auto_reset=True, randomizer=rand, max_episode_steps=4)
    b.reset()
    b.restore(ck)
    for i, w in zip(range(3, 6), want):
        got = b.step(acts[i])
        for x, y in zip(got, w):
            assert np.array_equal(x.cpu().numpy(), y)
    # two shards keyed by global robot ids reproduce the unsharded batch (incl. the randomizer draws and in-kernel resets)
    lo = common.make_product("laikago", ["laikago_pace"], 20, seed=21, gid0=0, auto_reset=True, randomizer=rand, max_episode_steps=4)
    hi = common.make_product("laikago", ["laikago_pace"], N - 20, seed=21, gid0=20, auto_reset=True, randomizer=rand, max_episode_steps=4)
    full = common.make_product("laikago", ["laikago_pace"], N, seed=21, auto_reset=True, randomizer=rand, max_episode_steps=4)
    ol, oh, of = lo.reset().clone(), hi.reset().clone(), full.reset().clone()
    assert torch.equal(torch.cat([ol, oh]), of)
    for i in range(6):                                   # crosses an episode boundary (time limit 4 steps)
        rl, rh, rf = lo.step(acts[i][:20].contiguous()), hi.step(acts[i][20:].contiguous()), full.step(acts[i])
        for x, y, z in zip(rl, rh, rf):
            assert torch.equal(torch.cat([x, y]), z)
    for e in (a, b, lo, hi, full):
        e.close()


@pytest.mark.parametrize("N", [3, 33])
def test_ragged_batch_with_randomizer_is_deterministic(N):
    """A batch that is not a multiple of the 8 robots of a warp, train mode (randomizer + randomised control latency):
    the idle quads of the last warp work on a copy of the last record and must not touch its observation rings.
    Two runs agree bit for bit, and with the first N robots of a larger batch keyed by the same global ids."""
    M = 64
    acts = [torch.as_tensor(np.random.default_rng(100 + i).normal(scale=0.15, size=(M, 12)).astype(np.float32), device="cuda") for i in range(10)]

    def run(n):
        env = common.make_product("laikago", ["laikago_pace"], n, seed=7, auto_reset=True, randomizer=True, max_episode_steps=4)
        out = [env.reset().clone()]
        for a in acts:
            out.extend(x.clone() for x in env.step(a[:n].contiguous()))
        rings = env.checkpoint()
        env.close()
        return out, rings

    a, ra = run(N)
    b, rb = run(N)
    full, _ = run(M)
    for x, y, z in zip(a, b, full):
        assert torch.equal(x, y)
        assert torch.equal(x, z[:N])
    for k in ra["state"]:
        assert np.array_equal(ra["state"][k], rb["state"][k]), k
    for x, y in zip(ra["latency_rings"], rb["latency_rings"]):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("robot,clips,N,rand", [("laikago", ["laikago_pace", "laikago_spin", "laikago_trot"], 2048 + 5, False),
                                                ("minicheetah", ["minicheetah_trot"], 77, True)])
def test_reference_packets_equal_the_inline_reference_evaluation(robot, clips, N, rand, monkeypatch):
    """The leading blocks of a step launch compute the reference sample of task.update and the target frames of the
    observation ahead of the workers (DevEnv::ahead); ORL_AHEAD=0 makes every worker compute them itself.  Both must give
    the same bits: observations, rewards, done flags every step and the final records -- across in-kernel resets (short
    episodes), cycle wraps of the clips, a ragged last warp, and records edited behind the library's back (a row whose
    key no longer matches its record must be dropped, not used)."""
    K = 24
    acts = [torch.as_tensor(np.random.default_rng(300 + i).normal(scale=0.2, size=(N, 12)).astype(np.float32), device="cuda") for i in range(K)]

    def run(flag):
        monkeypatch.setenv("ORL_AHEAD", flag)
        env = common.make_product(robot, clips, N, seed=11, auto_reset=True, randomizer=rand, max_episode_steps=9)
        out = [env.reset().clone()]
        for i, a in enumerate(acts):
            if i == 10:                      # move the motion clock of every third robot by hand: old packets are now wrong
                st = env.state.cpu().numpy().copy()
                toff = st[:, env.layout["time_offset"]:env.layout["time_offset"] + 2].copy().view(np.float64)
                toff[::3] += 0.0625
                st[:, env.layout["time_offset"]:env.layout["time_offset"] + 2] = toff.view(np.float32)
                env.state.copy_(torch.as_tensor(st, device=env.device))
            out.extend(x.clone() for x in env.step(a))
        out.append(env.state.clone())
        env.close()
        return out

    on, off = run("1"), run("0")
    for i, (x, y) in enumerate(zip(on, off)):
        assert torch.equal(x.view(torch.uint8), y.view(torch.uint8)), i


@pytest.mark.parametrize("N,rand", [(1024 + 3, False), (200, True)])
def test_host_buffer_steps_with_packet_rows_equal_the_device_api(N, rand, monkeypatch):
    """orl_step_host with page-locked buffers ships the observation row in pieces (DevEnv::obs_split) while the reference part
    of it comes out of packet rows computed one launch ahead.  Every step's row must equal the device API's, bit for bit,
    across in-kernel resets (short episodes: a reset robot rewrites its whole row, its next row comes from the prepared
    episode) and a ragged last warp, and equal the same API with ORL_AHEAD=0 (every packet computed inline)."""
    K = 30
    rng = np.random.default_rng(9)
    acts = [rng.normal(scale=0.2, size=(N, 12)).astype(np.float32) for _ in range(K)]

    def run(flag, host):
        monkeypatch.setenv("ORL_AHEAD", flag)
        env = common.make_product("laikago", ["laikago_pace", "laikago_trot"], N, seed=4, auto_reset=True, randomizer=rand, max_episode_steps=7)
        out = []
        if host:
            pa, pout = env.alloc_host_buffers()
            out.append(env.reset_host().copy())
            for a in acts:
                pa[...] = a
                out.extend(np.array(x).copy() for x in env.step_host(pa, pout))
        else:
            out.append(env.reset().cpu().numpy())
            for a in acts:
                out.extend(x.cpu().numpy() for x in env.step(torch.as_tensor(a, device=env.device)))
        env.close()
        return out

    dev, host_on, host_off = run("1", False), run("1", True), run("0", True)
    for i, (x, y, z) in enumerate(zip(dev, host_on, host_off)):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), i
        assert np.array_equal(x.view(np.uint8), z.view(np.uint8)), i


def test_time_limit_and_termination_thresholds(laikago):
    env, orc = laikago
    env.reset(); orc.reset()
    st = env.get_state()
    # push some robots over the root-distance / root-rotation thresholds, keep others just inside
    st["pos"][0:8, 0] += 1.05
    st["pos"][8:16, 0] += 0.90
    s, c = np.sin(0.85), np.cos(0.85)      # 1.7 rad > pi/2 about z
    rot = np.array([0, 0, s, c], dtype=np.float32)
    from oracle import xform as xf
    st["quat"][16:24] = xf.quaternion_multiply(rot, st["quat"][16:24]).astype(np.float32)
    st["env_step_counter"][24:32] = 598
    env.set_state(st)
    common.sync_oracle_from_product(orc, env.get_state())
    act = np.zeros((env.num_robot, 12), dtype=np.float32)
    _, _, d = env.step(torch.as_tensor(act, device=env.device))
    _, _, do = orc.step(act.astype(np.float64))
    d = d.cpu().numpy().astype(bool)
    safe = (orc._term_margin > 1e-4) & (orc._fall_margin > 1e-5)
    assert np.array_equal(d[safe], do[safe])
    # (a cycle wrap inside this step re-anchors the reference origin to the robot and legitimately clears the distance failure)
    assert d[0:8].sum() >= 6 and d[16:24].all() and do[0:8].sum() >= 6
    assert not d[24:32].any() or do[24:32].any()
    _, _, d2 = env.step(torch.as_tensor(act, device=env.device))       # env_step_counter reaches 600
    assert d2.cpu().numpy()[24:32].all()


# ---- size-independent properties at BASELINE.json's full size -------------------------------------------
def test_full_size_properties():
    N = 4096
    env = common.make_product("laikago", ["laikago_pace"], N, seed=21)
    obs0 = env.reset().clone()
    assert np.abs(env.reward_terms() - 1.0).max() <= 1e-5                 # perfect tracking after reset
    lo, hi = env.observation_bounds()
    o = obs0.cpu().numpy()
    # the Box bounds are not enforced by the reference either (warm-up targets hold the default pose, which
    # lies outside the clip's joint range); RSI robots must be inside them
    rsi = env.get_state()["warmup"] == 0
    assert np.isfinite(o).all() and (o[rsi] >= lo - 1e-4).all() and (o[rsi] <= hi + 1e-4).all()
    gen = torch.Generator(device=env.device); gen.manual_seed(0)
    acts = [torch.randn((N, 12), device=env.device, generator=gen) * 0.125 for _ in range(5)]
    outs = []
    for a in acts:
        o, r, d = env.step(a)
        outs.append((o.clone(), r.clone(), d.clone()))
    r_all = torch.stack([x[1] for x in outs]).cpu().numpy()
    assert np.isfinite(r_all).all() and r_all.min() >= 0.0 and r_all.max() <= 1.0 + 1e-6
    # determinism: same seed, same actions -> identical bits
    env2 = common.make_product("laikago", ["laikago_pace"], N, seed=21)
    assert torch.equal(env2.reset(), obs0)
    for a, (o1, r1, d1) in zip(acts, outs):
        o2, r2, d2 = env2.step(a)
        assert torch.equal(o2, o1) and torch.equal(r2, r1) and torch.equal(d2, d1)
    # sharding independence: two half-size handles keyed by global ids reproduce the full batch
    h0 = common.make_product("laikago", ["laikago_pace"], N // 2, seed=21, gid0=0)
    h1 = common.make_product("laikago", ["laikago_pace"], N // 2, seed=21, gid0=N // 2)
    assert torch.equal(torch.cat([h0.reset(), h1.reset()]), obs0)
    oa, ra, _ = h0.step(acts[0][:N // 2].contiguous())
    ob, rb, _ = h1.step(acts[0][N // 2:].contiguous())
    assert torch.equal(torch.cat([oa, ob]), outs[0][0]) and torch.equal(torch.cat([ra, rb]), outs[0][1])
    # the float64 oracle on a slice of the full batch: robots [3000, 3064) keyed by their global ids reset to the same
    # state (same Philox draws) and take the same first step
    k0, k = 3000, 64
    orc = common.make_oracle("laikago", ["laikago_pace"], k, seed=21, gid0=k0)
    oo = orc.reset()
    assert np.abs(oo - obs0.cpu().numpy()[k0:k0 + k]).max() <= 2e-5
    a0 = acts[0].cpu().numpy()[k0:k0 + k].astype(np.float64)
    o1, r1, d1 = orc.step(a0)
    pr, pd = outs[0][1].cpu().numpy()[k0:k0 + k], outs[0][2].cpu().numpy()[k0:k0 + k].astype(bool)
    safe = (orc._term_margin > 1e-4) & (orc._fall_margin > 1e-5)
    assert np.array_equal(pd[safe], d1[safe]) and safe.mean() > 0.9
    assert np.median(np.abs(pr - r1)) <= 1e-5 and np.abs(pr - r1).max() <= 2e-3
    for e in (env, env2, h0, h1):
        e.close()


def test_full_size_minicheetah_against_oracle_slice():
    """Config C3: 16384 mini-cheetahs on one GPU.  Determinism, sharding independence and the float64 oracle on a slice of
    the batch keyed by global robot ids (same Philox draws): reset and two control steps."""
    N = 16384
    env = common.make_product("mini_cheetah", ["minicheetah_trot"], N, seed=33)
    obs0 = env.reset().clone()
    assert np.abs(env.reward_terms() - 1.0).max() <= 1e-5
    gen = torch.Generator(device=env.device); gen.manual_seed(2)
    acts = [torch.randn((N, 12), device=env.device, generator=gen) * 0.1 for _ in range(2)]
    outs = [tuple(x.clone() for x in env.step(a)) for a in acts]
    assert all(torch.isfinite(o).all() and torch.isfinite(r).all() for o, r, _ in outs)
    # two shards reproduce the full batch bit for bit
    h0 = common.make_product("mini_cheetah", ["minicheetah_trot"], 10000, seed=33, gid0=0)
    h1 = common.make_product("mini_cheetah", ["minicheetah_trot"], N - 10000, seed=33, gid0=10000)
    assert torch.equal(torch.cat([h0.reset(), h1.reset()]), obs0)
    for a, (o, r, d) in zip(acts, outs):
        oa, ra, da = h0.step(a[:10000].contiguous())
        ob, rb, db = h1.step(a[10000:].contiguous())
        assert torch.equal(torch.cat([oa, ob]), o) and torch.equal(torch.cat([ra, rb]), r) and torch.equal(torch.cat([da, db]), d)
    # oracle slices at both ends and in the middle of the batch
    for k0 in (0, 9000, N - 64):
        k = 64
        orc = common.make_oracle("mini_cheetah", ["minicheetah_trot"], k, seed=33, gid0=k0, phys=common.oracle_phys_from_product(env.phys))
        oo = orc.reset()
        assert np.abs(oo - obs0.cpu().numpy()[k0:k0 + k]).max() <= 2e-5
        for i, (a, (o, r, d)) in enumerate(zip(acts, outs)):
            o1, r1, d1 = orc.step(a.cpu().numpy()[k0:k0 + k].astype(np.float64))
            pr, pd = r.cpu().numpy()[k0:k0 + k], d.cpu().numpy()[k0:k0 + k].astype(bool)
            safe = (orc._term_margin > 1e-4) & (orc._fall_margin > 1e-5)
            assert np.array_equal(pd[safe], d1[safe]) and safe.mean() > 0.9
            # the second step starts from two slightly different states (FP32 vs FP64 first step): medians only
            assert np.median(np.abs(pr - r1)) <= 2e-5 and (i > 0 or np.abs(pr - r1).max() <= 5e-3)
            assert np.median(np.abs(o.cpu().numpy()[k0:k0 + k] - o1)) <= 5e-5
    for e in (env, h0, h1):
        e.close()


def test_host_buffer_api_equals_device_api_and_ragged_sizes():
    for N in (1, 7, 9, 33):                      # ragged: not a multiple of the 8 robots a warp owns
        a = common.make_product("laikago", ["laikago_pace"], N, seed=2)
        b = common.make_product("laikago", ["laikago_pace"], N, seed=2)
        assert np.array_equal(a.reset().cpu().numpy(), b.reset_host())
        act = np.random.default_rng(N).normal(scale=0.1, size=(N, 12)).astype(np.float32)
        o, r, d = a.step(torch.as_tensor(act, device=a.device))
        oh, rh, dh = b.step_host(act)
        assert np.array_equal(o.cpu().numpy(), oh) and np.array_equal(r.cpu().numpy(), rh) and np.array_equal(d.cpu().numpy(), dh)
        # a batch is a prefix of a larger batch with the same seed (per-robot RNG keyed by global id)
        big = common.make_product("laikago", ["laikago_pace"], 40, seed=2)
        assert np.array_equal(big.reset_host()[:N], common.make_product("laikago", ["laikago_pace"], N, seed=2).reset_host())
        a.close(); b.close(); big.close()


def test_host_buffer_api_pinned_and_pageable_large_batch():
    """orl_step_host: page-locked buffers (direct DMA) and pageable ones (chunked staging) give the same bytes."""
    N = 8200                                      # > 1 MiB of observations: the chunked path; ragged last warp
    a = common.make_product("laikago", ["laikago_pace"], N, seed=5)
    b = common.make_product("laikago", ["laikago_pace"], N, seed=5)
    c = common.make_product("laikago", ["laikago_pace"], N, seed=5)
    pa, pout = b.alloc_host_buffers()
    ref0 = a.reset().cpu().numpy()
    assert np.array_equal(ref0, b.reset_host()) and np.array_equal(ref0, c.reset_host())
    rng = np.random.default_rng(0)
    for _ in range(3):
        act = rng.normal(scale=0.1, size=(N, 12)).astype(np.float32)
        o, r, d = a.step(torch.as_tensor(act, device=a.device))
        pa[...] = act
        op, rp, dp = b.step_host(pa, pout)
        oh, rh, dh = c.step_host(act)
        for x, y, z in ((o, op, oh), (r, rp, rh), (d, dp, dh)):
            assert np.array_equal(x.cpu().numpy(), y) and np.array_equal(y, z)
    a.close(); b.close(); c.close()


def test_auto_reset_and_statistics():
    N = 512
    env = common.make_product("laikago", ["laikago_pace"], N, seed=1, auto_reset=True, max_episode_steps=5)
    env.reset()
    env.stats(reset=True)
    act = torch.zeros((N, 12), device=env.device)
    tot_r = 0.0
    for i in range(5):
        o, r, d = env.step(act)
        tot_r += float(r.double().sum())
    assert d.cpu().numpy().all()                         # time limit after 5 steps
    st = env.get_state()
    assert (st["env_step_counter"] == 0).all() and (st["episode"] >= 2).all() and (st["ep_return"] == 0).all()
    s = env.stats()
    assert s[3] == 5 * N and abs(s[0] - tot_r) < 1e-3 * tot_r and s[2] >= N
    # without auto reset the caller resets; done robots keep stepping (SURVEY.md 8b "episode semantics")
    env2 = common.make_product("laikago", ["laikago_pace"], 16, seed=1, auto_reset=False, max_episode_steps=2)
    env2.reset()
    for i in range(3):
        _, _, d = env2.step(torch.zeros((16, 12), device=env2.device))
    assert d.cpu().numpy().all() and (env2.get_state()["env_step_counter"] == 3).all()
    env.close(); env2.close()


def test_errors_cross_the_boundary_as_python_exceptions():
    with pytest.raises(ValueError, match="wrong robot select"):
        common.make_product("spot", ["laikago_pace"], 4)
    from openroborl_b200.env import BatchedImitationEnv
    with pytest.raises(ValueError, match="No reference motions"):
        BatchedImitationEnv("laikago", [], num_robot=4)
    env = common.make_product("laikago", ["laikago_pace"], 4)
    with pytest.raises(ValueError):
        env.step(torch.zeros((3, 12), device=env.device))
    with pytest.raises(ValueError):
        env.set_max_episode_steps(0)
    # raw C-ABI argument checks: null / unaligned observation buffer, ring access without the randomizer
    import ctypes as C
    lib, h = env._lib, env._h
    act = torch.zeros((4, 12), device=env.device)
    buf = torch.zeros(4 * 160 + 4, device=env.device)
    rew, done = torch.zeros(4, device=env.device), torch.zeros(4, dtype=torch.uint8, device=env.device)
    P = lambda t, off=0: C.c_void_p(t.data_ptr() + off)
    assert lib.orl_step(h, P(act), P(buf, 4), P(rew), P(done), None) == -1          # ORL_ERR_INVALID: 16-byte alignment
    assert b"aligned" in lib.orl_last_error(h)
    assert lib.orl_step(h, P(act), None, P(rew), P(done), None) == -1
    assert lib.orl_step(h, P(act), P(buf), P(rew), P(done), None) == 0
    assert lib.orl_get_latency_rings_host(h, None, None) == -1 and b"randomizer" in lib.orl_last_error(h)
    env.close()
    # more clips than the kernel-parameter tables hold
    clips = common.product_clips(["laikago_pace"]) * 33
    with pytest.raises(ValueError, match="at most 32"):
        BatchedImitationEnv("laikago", clips, num_robot=4)


def test_policy_rollout_statistics():
    """Pretrained reference policy (golden actor MLP) as the action source, 64 robots, 40 steps.

    Trajectories are chaotic, so the FP32 kernel and the FP64 oracle are compared by episode statistics
    (north_star: "contact rollouts ... by episode-reward statistics"): mean reward within 0.02.
    """
    N = 64
    policy = common.load_policy("laikago_pace")
    env = common.make_product("laikago", ["laikago_pace"], N, seed=8)
    orc = common.make_oracle("laikago", ["laikago_pace"], N, seed=8, phys=common.oracle_phys_from_product(env.phys))
    o = env.reset().cpu().numpy()
    oo = orc.reset()
    rs, ros = [], []
    for t in range(40):
        _, r, _ = env.step(torch.as_tensor(policy(o), device=env.device))
        o = env.obs.cpu().numpy()
        oo, ro, _ = orc.step(policy(oo).astype(np.float64))
        rs.append(r.cpu().numpy()); ros.append(ro)
    rs, ros = np.array(rs), np.array(ros)
    assert np.abs(rs[:5] - ros[:5]).max() < 5e-3            # short horizon: still the same trajectory
    assert abs(rs.mean() - ros.mean()) < 0.02
    env.close()


@pytest.mark.parametrize("robot,clip", [("laikago", "laikago_pace"), ("laikago", "laikago_spin"), ("mini_cheetah", "minicheetah_trot")])
def test_cuda_path_against_reference_class_trace(robot, clip):
    """tests/golden/ref_task_trace.npz was produced by the reference's own Minitaur/ImitationTask/LocomotionGymEnv/
    WrapperEnv code (scripts/make_trace.py).  The CUDA env replays the recorded reset draws and actions; trajectories
    are chaotic over a rollout, so the comparison is tight on the first steps and statistical afterwards."""
    TR = np.load(common.GOLD + "/ref_task_trace.npz")
    n = len(TR[clip + "/clip"])
    env = common.make_product(robot, [clip], n)
    orc = common.make_oracle(robot, [clip], n, phys=common.oracle_phys_from_product(env.phys))
    orc.reset(draws=(TR[clip + "/u_rsi"], TR[clip + "/clip"], np.zeros(n)), time_offset=TR[clip + "/time_offset"])
    env.reset()
    common.sync_product_from_oracle_reset(env, orc)
    ref_obs, ref_rew, ref_done = TR[clip + "/obs"], TR[clip + "/reward"], TR[clip + "/done"]
    rews = []
    for t, a in enumerate(TR[clip + "/action"]):
        o, r, d = env.step(torch.as_tensor(a.astype(np.float32), device=env.device))
        o, r, d = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy().astype(bool)
        rews.append(r)
        if t < 4:
            assert np.abs(o - ref_obs[t + 1]).max() < 5e-3 and np.abs(r - ref_rew[t]).max() < 1e-3, t
            assert np.array_equal(d, ref_done[t])
    rews = np.array(rews)
    assert abs(rews.mean() - ref_rew.mean()) < 0.03
    env.close()


@pytest.mark.parametrize("robot,clip,source", [("laikago", "laikago_pace", "policy"), ("laikago", "laikago_pace", "tracking"),
                                               ("mini_cheetah", "minicheetah_trot", "tracking")])
def test_config_c1_single_robot_episode_every_step_against_oracle(robot, clip, source):
    """BASELINE.json configs[0] / SURVEY.md 8(d) C1: ONE robot in test mode (episode length 600, RSI 0.9, warm-up 0.25 s,
    randomizer off).  Action source: the reference's pretrained policy (deterministic mean action; trained against Bullet
    and the real URDF, it loses the clip within ~25 steps on this repo's physics model and synthetic URDF -- DESIGN.md
    section 5) or the open-loop tracking controller of the bench workload (next reference motor angles - INIT).
    The CUDA env runs the episode until it ends; EVERY control step of it is then checked against the float64 oracle in
    one batched oracle step: oracle robot k starts from the recorded state before step k and receives the recorded action k
    (per-step parity traces of observation, reward and done; the free-running trajectories themselves are chaotic)."""
    policy = common.load_policy(clip) if source == "policy" else None
    env = common.make_product(robot, [clip], 1, seed=0, max_episode_steps=600)
    spec, mdl = env.model.spec, env.model
    perm = [mdl.q_index[j] for j in mdl.motor_id_list]
    o = env.reset().cpu().numpy()
    states, acts, obs, rew, done = [], [], [], [], []
    for t in range(600):
        states.append(env.get_state())
        if policy is not None:
            a = policy(o)
        else:
            a = (o[:, 84 + 7:84 + 19][:, perm] - np.asarray(spec.JOINT_OFFSETS)) * np.asarray(spec.JOINT_DIRECTIONS) - np.asarray(spec.INIT_MOTOR_ANGLES)
        a = np.clip(a, -2 * np.pi, 2 * np.pi).astype(np.float32)
        ob, r, d = env.step(torch.as_tensor(a, device=env.device))
        o = ob.cpu().numpy()
        acts.append(a[0]); obs.append(o[0].copy()); rew.append(float(r[0])); done.append(bool(d[0]))
        if done[-1]:
            break
    T = len(acts)
    assert done[-1] and not any(done[:-1])
    assert T >= 10, T
    orc = common.make_oracle(robot, [clip], T, phys=common.oracle_phys_from_product(env.phys), max_episode_steps=600)
    orc.reset()
    common.sync_oracle_from_product(orc, {k: np.concatenate([s[k] for s in states], axis=0) for k in states[0]})
    oo, ro, do = orc.step(np.asarray(acts, dtype=np.float64))
    obs, rew, done = np.asarray(obs), np.asarray(rew), np.asarray(done)
    safe = (orc._term_margin > 1e-4) & (orc._fall_margin > 1e-5)
    assert safe.mean() > 0.9 and np.array_equal(done[safe], do[safe])
    assert do[-1]
    ok = np.abs(obs - oo).max(axis=1) < 2e-2                 # steps without a contact toggling on a knife edge (as in test_control_step_matches_oracle)
    assert ok.mean() > 0.97
    er, eo = np.abs(rew - ro)[ok], np.abs(obs - oo).max(axis=1)[ok]
    key = "c1_single_robot_episode_%s_%s" % (robot, source)
    ERROR_TABLE[key] = dict(steps=int(T), mean_reward=float(rew.mean()), knife_edge_steps=int((~ok).sum()),
                            reward=dict(p50=float(np.percentile(er, 50)), p99=float(np.percentile(er, 99)), max=float(er.max())),
                            obs=dict(p50=float(np.percentile(eo, 50)), p99=float(np.percentile(eo, 99)), max=float(eo.max())))
    dump_error_table()
    assert np.percentile(er, 50) <= 2e-6 and np.percentile(er, 99) <= 5e-5 and er.max() <= 5e-4, ERROR_TABLE[key]
    assert np.percentile(eo, 50) <= 1e-4 and np.percentile(eo, 99) <= 5e-3, ERROR_TABLE[key]
    env.close()
