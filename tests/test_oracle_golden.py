"""The oracle against the reference's own outputs (tests/golden/ref_motion_kat.npz, made by
scripts/make_golden.py from the UNMODIFIED reference modules) and SURVEY.md Appendix B."""
import numpy as np
import pytest

from tests import common
from oracle import xform as xf
from oracle.motion import MotionLib
from oracle.env import BUTTER_A, BUTTER_B

KAT = np.load(common.GOLD + "/ref_motion_kat.npz")
CLIPS = ["laikago_pace", "laikago_spin", "laikago_hopturn", "minicheetah_trot", "synthetic_pace_clamp"]   # the last one: LoopMode Clamp, generated with the reference module


@pytest.mark.parametrize("name", CLIPS)
def test_motion_tables_match_reference(name):
    lib = MotionLib(common.golden_clips_raw([name]))
    assert np.array_equal(lib.frames, KAT[name + "/frames"])                 # _postprocess_frames, bit exact
    np.testing.assert_allclose(lib.vels, KAT[name + "/vels"], rtol=0, atol=1e-12)
    assert lib.duration[0] == float(KAT[name + "/duration"])
    np.testing.assert_allclose(lib.delta_pos[0], KAT[name + "/delta_pos"], atol=0)
    np.testing.assert_allclose(lib.delta_heading[0], KAT[name + "/delta_heading"], atol=1e-15)


@pytest.mark.parametrize("name", CLIPS)
def test_frame_lookup_matches_reference(name):
    lib = MotionLib(common.golden_clips_raw([name]))
    t = KAT[name + "/times"]
    clip = np.zeros(len(t), dtype=np.int64)
    f0, f1, blend = lib.calc_blend_idx(clip, t)
    assert np.array_equal(f0, KAT[name + "/f0"]) and np.array_equal(f1, KAT[name + "/f1"])       # bit exact indices
    assert np.array_equal(blend, KAT[name + "/blend"])
    assert np.array_equal(lib.calc_cycle_count(clip, t), KAT[name + "/cycle"])
    assert np.array_equal(lib.calc_phase(clip, t), KAT[name + "/phase"])
    np.testing.assert_allclose(lib.calc_frame(clip, t), KAT[name + "/pose"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(lib.calc_frame_vel(clip, t), KAT[name + "/vel"], rtol=0, atol=1e-12)


def test_appendix_b_known_answers():
    """SURVEY.md Appendix B rows computed with the reference's motion_data.py."""
    lib = MotionLib(common.golden_clips_raw(["laikago_pace", "laikago_spin"]))
    assert lib.num_frames[0] == 39 and abs(lib.duration[0] - 0.63346) < 1e-12
    np.testing.assert_allclose(lib.delta_pos[0], [0.68773, 0, 0], atol=1e-12)
    assert abs(lib.delta_heading[0] - (-0.05141235558665544)) < 1e-12
    t = np.array([0.0, 0.033, 0.5, 0.7, 1.2345, -0.1])
    f0, f1, bl = lib.calc_blend_idx(np.zeros(6, dtype=int), t)
    assert list(f0) == [0, 1, 29, 3, 36, 32] and list(f1) == [1, 2, 30, 4, 37, 33]
    np.testing.assert_allclose(bl, [0.0, 0.9796040791841631, 0.9940011997600486, 0.9916016796640604,
                                    0.055188962207553496, 0.0011997600479916044], atol=1e-12)
    assert list(lib.calc_cycle_count(np.zeros(6, dtype=int), t)) == [0, 0, 0, 1, 1, -1]
    pose = lib.calc_frame(np.zeros(6, dtype=int), t)
    np.testing.assert_allclose(pose[3, :8], [0.75346672, 0.00881513, 0.44208698, 0.47097854, 0.53858454, 0.51301071,
                                             0.47426354, -0.12087756], atol=1e-8)
    assert abs(lib.duration[1] - 0.75015) < 1e-12 and abs(lib.delta_heading[1] - 1.7393955112255681) < 1e-12
    spin = lib.calc_frame(np.array([1]), np.array([2.0]))[0, :7]
    np.testing.assert_allclose(spin, [0.0114991, -0.00821536, 0.44762465, 0.68102652, 0.03075894, 0.03614657, 0.73071896], atol=1e-8)


def test_pose3d_helpers_match_reference():
    q, pts = KAT["pose3d/q"], KAT["pose3d/pts"]
    np.testing.assert_allclose(xf.calc_heading(q), KAT["pose3d/heading"], atol=1e-14)
    np.testing.assert_allclose(xf.QuaternionRotatePoint(pts, q), KAT["pose3d/rotate"], atol=1e-14)
    axis, angle = xf.QuaternionToAxisAngle(q)
    np.testing.assert_allclose(axis, KAT["pose3d/axis"], atol=1e-14)
    np.testing.assert_allclose(angle, KAT["pose3d/angle"], atol=1e-14)
    np.testing.assert_allclose(xf.normalize_rotation_angle(KAT["pose3d/ang_in"]), KAT["pose3d/norm_angle"], atol=1e-14)
    np.testing.assert_allclose(xf.MapToMinusPiToPi(KAT["pose3d/ang_in"]), KAT["pose3d/map_pi"], atol=1e-14)
    np.testing.assert_allclose(xf.calc_heading_rot(q), KAT["pose3d/heading_rot"], atol=1e-14)
    assert abs(xf.calc_heading(np.array([0.5, 0.5, 0.5, 0.5])) - np.pi / 2) < 1e-15
    ax, an = xf.QuaternionToAxisAngle(np.array([0, 0, np.sin(0.3), np.cos(0.3)]))
    np.testing.assert_allclose(ax, [0, 0, 1]); assert abs(an - 0.6) < 1e-15
    with pytest.raises(ValueError):
        xf.QuaternionToAxisAngle(np.array([0.0, 0.0, 0.0, 2.0]))              # pose3d.py:165-167


def test_butterworth_matches_reference():
    np.testing.assert_allclose(BUTTER_B, KAT["filter/b"], atol=1e-16)
    np.testing.assert_allclose(BUTTER_A, KAT["filter/a"], atol=1e-16)
    env = common.make_oracle("laikago", ["laikago_pace"], 1)
    x0 = KAT["filter/x0"]
    env.filt_x[:] = x0; env.filt_y[:] = x0
    env.substep_counter[:] = 1      # history already initialised
    ys = []
    for x in KAT["filter/x"]:
        env.set_act((x - env.spec.init_motor)[None])
        ys.append(env.action[0].copy())
    np.testing.assert_allclose(ys, KAT["filter/y"], atol=1e-14)


def test_pd_motor_model_matches_reference():
    env = common.make_oracle("laikago", ["laikago_pace"], 8)
    s = env.spec
    tau = -1 * (s.kp * (KAT["motor/q"] - KAT["motor/cmd"])) - s.kd * KAT["motor/qd"]
    np.testing.assert_allclose(tau, KAT["motor/tau"], atol=1e-12)
    # Appendix B: q* - q = 0.05, qdot = 0.5 -> [10.85, 10.0, 10.0]
    np.testing.assert_allclose((-s.kp * (-0.05) - s.kd * 0.5)[:3], [10.85, 10.0, 10.0], atol=1e-12)


def test_observation_bounds_match_policy_zip():
    """The bounds serialised in the reference's policy zips pin the observation layout."""
    for robot, clip, pol in (("laikago", "laikago_pace", "laikago_pace"), ("mini_cheetah", "minicheetah_trot", "minicheetah_trot")):
        z = np.load(common.GOLD + "/policy_%s.npz" % pol)
        if "obs_low" not in z.files:
            pytest.skip("bounds not stored")
        env = common.make_oracle(robot, [clip], 1)
        low, high = env.observation_bounds()
        np.testing.assert_allclose(low, z["obs_low"], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(high, z["obs_high"], rtol=1e-6, atol=1e-6)


def test_clamp_mode_clip_ends():
    """LoopMode Clamp (motion_data.py:265-276, 682-718): indices clamp at both ends, the cycle count to [0, 1], is_over."""
    name = "synthetic_pace_clamp"
    lib = MotionLib(common.golden_clips_raw([name]))
    t = KAT[name + "/times"]
    clip = np.zeros(len(t), dtype=np.int64)
    assert np.array_equal(lib.is_over(clip, t), KAT[name + "/is_over"])
    assert KAT[name + "/is_over"].any() and not KAT[name + "/is_over"].all()
    assert set(np.unique(KAT[name + "/cycle"])) == {0, 1}


def test_motor_model_all_modes_against_reference_class():
    """minitaur_motor.py:104-173 with strength ratios and torque limits: POSITION, TORQUE (unclipped), HYBRID."""
    from oracle import motor
    q, qd, cmd = KAT["motor/q"], KAT["motor/qd"], KAT["motor/cmd"]
    kp, kd = np.full(12, 220.0), np.array([0.3, 2.0, 2.0] * 4)
    lim, sr, hyb = KAT["motor/limits"], KAT["motor/strength"], KAT["motor/hybrid_cmd"]
    np.testing.assert_allclose(motor.convert_to_torque("position", cmd, q, qd, kp, kd), KAT["motor/tau"], rtol=0, atol=1e-12)
    np.testing.assert_array_equal(motor.convert_to_torque("position", cmd, q, qd, kp, kd, sr, lim), KAT["motor/tau_position_limited"])
    np.testing.assert_array_equal(motor.convert_to_torque("torque", cmd, q, qd, kp, kd, sr, lim), KAT["motor/tau_torque"])
    np.testing.assert_array_equal(motor.convert_to_torque("hybrid", hyb, q, qd, None, None, sr, lim), KAT["motor/tau_hybrid"])
    assert (np.abs(KAT["motor/tau_position_limited"]) == lim[None]).any()          # the limits do clip in the fixture
    assert (np.abs(KAT["motor/tau_torque"]) > lim[None]).any()                     # TORQUE mode is returned before the clip
    with pytest.raises(ValueError):
        motor.convert_to_torque("pwm", cmd, q, qd, kp, kd)
