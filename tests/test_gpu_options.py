"""Task options R/run.py leaves at their defaults -- clip switching (imitation_task.py:740-754, 1057-1101), initial-state
perturbation (:192-195, 1199-1243), target-observation noise (:274-275), reward weight (:356), fixed start time (:1104) --
and the other control modes of the motor model (minitaur_motor.py:104-173): CUDA path against the oracle, which is pinned
to the reference's own classes with these options on (tests/test_oracle_trace.py, ref_task_options_trace.npz)."""
import numpy as np
import pytest
import torch

from tests import common

pytestmark = pytest.mark.gpu

CLIPS = ["laikago_pace", "laikago_trot", "laikago_spin"]
TASK = dict(clip_time_min=0.2, clip_time_max=0.5, perturb_init_state_prob=0.6, tar_obs_noise=[0.1, 0.2], weight=0.7)


def make_pair(N, seed=5, task=TASK, clips=CLIPS):
    env = common.make_product("laikago", clips, N, seed=seed, task=task)
    orc = common.make_oracle("laikago", clips, N, seed=seed, task=task, phys=common.oracle_phys_from_product(env.phys))
    return env, orc


def test_option_reset_matches_oracle():
    N = 512
    env, orc = make_pair(N)
    obs = env.reset().cpu().numpy()
    obs_o = orc.reset()
    st = env.get_state()
    assert np.array_equal(st["clip"], orc.clip) and np.array_equal(st["warmup"].astype(bool), orc.warmup)
    assert np.array_equal(st["time_offset"], orc.time_offset) and np.array_equal(st["prev_phase"], orc.prev_phase)
    assert np.array_equal(st["clip_change_time"], orc.clip_change_time)                       # FP64, bit exact
    assert np.all(st["clip_changes"] == 0)
    dur = orc.clip_change_time - orc.motion_time()
    assert dur.min() >= 0.2 and dur.max() <= 0.5
    # the reference state is NOT perturbed, the simulated one is (about 60 % of the robots)
    for k in ("ref_pose", "ref_vel", "origin_pos", "origin_rot"):
        assert np.abs(st[k] - getattr(orc, k)).max() <= 2e-6 * max(1.0, np.abs(getattr(orc, k)).max()), k
    for k in ("pos", "quat", "q", "lin", "ang", "qd"):
        assert np.abs(st[k] - getattr(orc, k)).max() <= 3e-6 * max(1.0, np.abs(getattr(orc, k)).max()), k
    moved = np.abs(st["q"] - st["ref_pose"][:, 7:]).max(axis=1) > 1e-4
    assert 0.5 < moved.mean() < 0.7
    assert np.array_equal(moved, np.abs(orc.q - orc.ref_pose[:, 7:]).max(axis=1) > 1e-4)      # the same robots
    # noisy heading in the target frames: against the oracle, and different from the noise-free observation
    assert np.abs(obs - obs_o).max() <= 3e-6 * np.abs(obs_o).max()
    quiet_env, _ = make_pair(N, task=dict(TASK, tar_obs_noise=None))
    quiet = quiet_env.reset().cpu().numpy()
    assert np.abs(obs[:, 84:87] - quiet[:, 84:87]).max() > 1e-3 and np.array_equal(obs[:, :84], quiet[:, :84])
    quiet_env.close(); env.close()


def test_option_steps_match_oracle_through_clip_changes():
    N = 256
    env, orc = make_pair(N)
    env.reset(); orc.reset()
    rng = np.random.default_rng(8)
    changes, obs_err, rew_err = 0, [], []
    for step in range(24):
        st0 = env.get_state()
        common.sync_oracle_from_product(orc, st0)
        pol = (st0["ref_pose"][:, 7:][:, [env.model.q_index[j] for j in env.model.motor_id_list]]
               - np.asarray(env.model.spec.JOINT_OFFSETS)) * np.asarray(env.model.spec.JOINT_DIRECTIONS) - np.asarray(env.model.spec.INIT_MOTOR_ANGLES)
        act = (pol + rng.normal(scale=0.1, size=(N, 12))).astype(np.float32)
        o, r, d = env.step(torch.as_tensor(act, device=env.device))
        oo, ro, do = orc.step(act.astype(np.float64))
        o, r, d = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy().astype(bool)
        st = env.get_state()
        # the clock side is FP64 on both sides: decisions, draws and times are bit exact
        assert np.array_equal(st["clip"], orc.clip), step
        assert np.array_equal(st["clip_changes"], orc.clip_changes)
        assert np.array_equal(st["clip_change_time"], orc.clip_change_time)
        assert np.array_equal(st["time_offset"], orc.time_offset)
        assert np.array_equal(st["prev_phase"], orc.prev_phase)
        changed = st["clip_changes"] != st0["clip_changes"]
        changes += int(changed.sum())
        ok = np.abs(st["qd"] - orc.qd).max(axis=1) < 2e-2                                   # no knife-edge contact (test_gpu_parity.py)
        assert ok.mean() > 0.95
        # origin re-sync at a clip change uses the simulated heading: FP32 state vs FP64 oracle from the same start
        assert np.abs(st["origin_rot"] - orc.origin_rot)[ok].max() <= 2e-4
        assert np.abs(st["origin_pos"] - orc.origin_pos)[ok].max() <= 5e-4
        safe = (orc._term_margin > 1e-4) & (orc._fall_margin > 1e-5)
        assert safe.mean() > 0.9 and np.array_equal(d[safe], do[safe])
        obs_err.append(np.abs(o - oo).max(axis=1)[ok]); rew_err.append(np.abs(r - ro)[ok])
    assert changes >= N                                                                       # every robot switched about twice
    obs_err, rew_err = np.concatenate(obs_err), np.concatenate(rew_err)
    assert np.percentile(rew_err, 50) <= 1e-6 and np.percentile(rew_err, 99) <= 2e-5 and rew_err.max() <= 5e-4
    assert np.percentile(obs_err, 50) <= 1e-4 and np.percentile(obs_err, 99) <= 2e-3 and obs_err.max() <= 5e-2
    env.close()


def test_reward_weight_and_fixed_start_time():
    N = 64
    base = common.make_product("laikago", ["laikago_pace"], N, seed=2)
    half = common.make_product("laikago", ["laikago_pace"], N, seed=2, task=dict(weight=0.5))
    fixed = common.make_product("laikago", ["laikago_pace"], N, seed=2, task=dict(enable_rand_init_time=False))
    ob, oh = base.reset().clone(), half.reset().clone()
    fixed.reset()
    assert torch.equal(ob, oh)
    assert np.all(fixed.get_state()["time_offset"] == 0.0) and np.any(base.get_state()["time_offset"] != 0.0)
    act = torch.as_tensor(np.random.default_rng(0).normal(scale=0.1, size=(N, 12)).astype(np.float32), device="cuda")
    _, rb, _ = base.step(act)
    _, rh, _ = half.step(act)
    assert torch.allclose(rh, 0.5 * rb, rtol=1e-6, atol=0)
    for e in (base, half, fixed):
        e.close()


def test_options_with_auto_reset_do_not_depend_on_sharding():
    """In-kernel resets (prepared episode records included) with every option on: two shards keyed by global ids = one batch."""
    N = 48
    acts = [torch.as_tensor(np.random.default_rng(i).normal(scale=0.3, size=(N, 12)).astype(np.float32), device="cuda") for i in range(12)]
    mk = lambda n, g: common.make_product("laikago", CLIPS, n, seed=21, gid0=g, auto_reset=True, task=TASK, max_episode_steps=5)
    lo, hi, full = mk(20, 0), mk(N - 20, 20), mk(N, 0)
    assert torch.equal(torch.cat([lo.reset(), hi.reset()]), full.reset())
    for a in acts:
        rl, rh, rf = lo.step(a[:20].contiguous()), hi.step(a[20:].contiguous()), full.step(a)
        for x, y, z in zip(rl, rh, rf):
            assert torch.equal(torch.cat([x, y]), z)
    st = full.get_state()
    assert st["episode"].min() >= 3 and np.isfinite(st["clip_change_time"]).all()
    for e in (lo, hi, full):
        e.close()


def test_compat_task_accepts_the_options():
    from openroborl_b200 import compat
    files = common.product_clips(CLIPS)
    robot = [compat.Minitaur(name_robot="laikago", robot_index=i, enable_randomizer=False) for i in range(4)]
    task = [compat.ImitationTask(ref_motion_filenames=files, enable_cycle_sync=True, tar_frame_steps=[1, 2, 10, 30], ref_state_init_prob=0.9,
                                 warmup_time=0.25, clip_time_min=0.2, clip_time_max=0.5, perturb_init_state_prob=0.6,
                                 tar_obs_noise=[0.1, 0.2], weight=0.7, clip_velocity=3.0) for _ in range(4)]
    env = compat.WrapperEnv(compat.LocomotionGymEnv(robot, task), episode_length_start=600, episode_length_end=600,
                            curriculum_steps=30000000, num_parallel_envs=1)
    obs = env.reset()
    for _ in range(20):
        obs, rew, done, info = env.step([np.zeros(12) for _ in range(4)])
    assert task[0]._env._cuda.get_state()["clip_changes"].min() >= 1
    assert len(obs) == 4 and len(obs[0]) == 160 and all(0.0 <= r <= 0.7 for r in rew)
    assert 0.0 <= task[0].reward() <= 0.7                  # reward of the CURRENT state against the updated reference, x weight
    env.close()


# ---- motor model, all control modes -----------------------------------------------------------------------
def motor_reference(mode, cmd, q, qd, kp, kd, strength, limits):
    """minitaur_motor.py:104-173 restated in NumPy float64 (pinned to the reference class in tests/test_oracle_golden.py)."""
    from oracle import motor
    return motor.convert_to_torque(mode, cmd, q, qd, kp, kd, strength, limits)


@pytest.mark.parametrize("mode", ["position", "torque", "hybrid"])
@pytest.mark.parametrize("limits", [False, True])
def test_motor_model_modes(mode, limits):
    from openroborl_b200 import env as E
    rng = np.random.default_rng(3)
    n = 300
    q, qd = rng.normal(size=(n, 12)), rng.normal(scale=5, size=(n, 12))
    kp, kd = rng.uniform(50, 250, size=12), rng.uniform(0.5, 3, size=12)
    strength = rng.uniform(0.8, 1.2, size=(n, 12))
    lim = rng.uniform(5, 40, size=12) if limits else None
    if mode == "hybrid":
        cmd = rng.normal(size=(n, 60)); cmd[:, 1::5] = rng.uniform(50, 250, size=(n, 12)); cmd[:, 3::5] = rng.uniform(0.5, 3, size=(n, 12))
    else:
        cmd = rng.normal(scale=3.0, size=(n, 12))
    t = lambda a: None if a is None else torch.as_tensor(np.asarray(a, np.float32), device="cuda").contiguous()
    code = {"position": E.MOTOR_POSITION, "torque": E.MOTOR_TORQUE, "hybrid": E.MOTOR_HYBRID}[mode]
    out = E.motor_convert_to_torque(code, t(cmd), t(q), t(qd), t(kp), t(kd), t(strength), t(lim)).cpu().numpy()
    f32 = lambda a: None if a is None else np.asarray(a, np.float32).astype(np.float64)
    want = motor_reference(mode, f32(cmd), f32(q), f32(qd), f32(kp), f32(kd), f32(strength), f32(lim))
    assert np.abs(out - want).max() <= 2e-6 * np.abs(want).max()
    if limits and mode != "torque":
        assert (np.abs(out) <= np.asarray(lim, np.float32)[None] * (1 + 1e-6)).all() and (np.abs(out) >= lim.min() * 0.999).any()


def test_motor_model_rejects_unknown_mode():
    from openroborl_b200 import env as E
    x = torch.zeros((2, 12), device="cuda")
    with pytest.raises(ValueError):
        E.motor_convert_to_torque(4, x, x, x, x[0], x[0])                  # PWM: "not a supported motor control mode"
    with pytest.raises(AssertionError):
        E.motor_convert_to_torque(E.MOTOR_HYBRID, x, x, x)                 # 60 commands per robot expected
