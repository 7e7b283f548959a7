"""Host logic: URDF compiler, link-id bookkeeping, motion-file reader, error behaviour."""
import json
import os

import numpy as np
import pytest

from tests import common
from openroborl_b200.model import RobotModel
from openroborl_b200.motion import MotionClip, load_clip_library
from openroborl_b200 import robots


def _rz(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def compiled_link_com(model, pos, R_base, q):
    """numpy FK straight from the compiled ctypes model (what the CUDA leg_fk does)."""
    c = model.c
    out = np.zeros((c.n_links, 3))
    for l in range(4):
        L = c.leg[l]
        R, p = R_base.copy(), np.zeros(3)
        frames = []
        for k in range(3):
            p = p + R @ np.array(L.p_fix[k])
            R = R @ np.array(L.R_fix[k]).reshape(3, 3) @ _rz(q[3 * l + k])
            frames.append((R.copy(), p.copy()))
        for i in range(c.n_links):
            if c.link_leg[i] == l:
                Rk, pk = frames[c.link_body[i]]
                out[i] = pos + pk + Rk @ np.array(c.link_com[i])
    return out


@pytest.mark.parametrize("robot", ["laikago", "mini_cheetah"])
def test_compiled_model_fk_matches_textbook_tree(robot):
    from oracle.rbd import UrdfTree
    from oracle import xform as xf
    model = RobotModel(robot)
    tree = UrdfTree(common.URDF[robot])
    rng = np.random.default_rng(0)
    for _ in range(5):
        quat = rng.normal(size=4); quat /= np.linalg.norm(quat)
        pos, q = rng.normal(size=3), rng.uniform(-1, 1, size=12)
        ref = tree.link_com(pos[None], quat[None], q[None])[0]
        got = compiled_link_com(model, pos, xf.quat_to_matrix(quat), q)
        np.testing.assert_allclose(got, ref, atol=1e-6)       # the compiled model is stored in float32
    assert abs(model.mass_total - tree.mass.sum()) < 1e-5


def test_link_id_sets_follow_reference_regexes():
    """Laikago ids for URDF order FR,FL,RR,RL x (hip, upper, lower, toe)."""
    m = RobotModel("laikago")
    assert m.chassis_link_ids == [-1, 0, 4, 8, 12]
    assert m.motor_link_ids == [1, 5, 9, 13]
    assert m.knee_link_ids == [2, 6, 10, 14]
    assert m.foot_link_ids == [2, 3, 6, 7, 10, 11, 14, 15]
    assert m.leg_link_ids == [2, 3, 6, 7, 10, 11, 14, 15]
    assert m.motor_id_list == [0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14]
    mc = RobotModel("mini_cheetah")
    # motor order fl, hl, fr, hr is NOT the URDF order: mapping by joint name (minitaur.py:860-864)
    assert [mc.joint_names[i] for i in mc.motor_id_list] == robots.mini_cheetah.MOTOR_NAMES
    assert sorted(sum([list(mc.c.leg[l].motor) for l in range(4)], [])) == list(range(12))
    for l in range(4):
        assert mc.c.leg[l].n_ee == 2 and m.c.leg[l].n_ee == 2


def test_robot_name_check_and_both_minicheetah_spellings():
    with pytest.raises(ValueError, match="wrong robot select"):       # minitaur.py:96-97
        robots.get("spot")
    assert robots.get("minicheetah") is robots.get("mini_cheetah")    # run.py:57 spelling accepted


def test_default_pose_and_constants():
    m = RobotModel("laikago")
    np.testing.assert_allclose(m.default_joint_pose(), [0, 0.07, -0.59] * 4, atol=1e-12)
    assert robots.laikago.NUM_ACTION_REPEAT == 33 and robots.laikago.T_STEP == 0.001
    assert list(robots.laikago.JOINT_DIRECTIONS) == [-1, 1, 1, 1, 1, 1, -1, 1, 1, 1, 1, 1]


def test_motion_json_round_trip(tmp_path):
    clips = load_clip_library(common.CLIPS)
    assert len(clips) == 12                      # the reference's 11 clips + the synthetic Clamp-mode one
    by = {c.name: c for c in clips}
    assert by["laikago_pace"].frames.shape == (39, 19) and by["laikago_sidesteps"].frames.shape == (146, 19)
    assert by["laikago_spin"].cycle_rot and not by["laikago_pace"].cycle_rot
    assert all(c.loop_mode == "Wrap" and c.cycle_pos for c in clips if not c.name.startswith("synthetic"))
    assert by["synthetic_pace_clamp"].loop_mode == "Clamp" and not by["synthetic_pace_clamp"].cycle_pos
    p = tmp_path / "clip.txt"
    p.write_text(by["laikago_spin"].to_json())
    back = MotionClip.load(str(p))
    assert np.array_equal(back.frames, by["laikago_spin"].frames) and back.flags == by["laikago_spin"].flags == 7
    assert back.frame_duration == by["laikago_spin"].frame_duration


def test_motion_reader_errors():
    good = dict(LoopMode="Wrap", FrameDuration=0.1, Frames=[[0.0] * 3 + [0, 0, 0, 1] + [0.0] * 12])
    MotionClip.from_json(json.dumps(good))
    with pytest.raises(KeyError):
        MotionClip.from_json(json.dumps(dict(good, LoopMode="Bounce")))
    with pytest.raises(AssertionError):
        MotionClip.from_json(json.dumps(dict(good, FrameDuration=0.0)))
    with pytest.raises(AssertionError):
        MotionClip.from_json(json.dumps(dict(good, Frames=[[0.0] * 7])))
    with pytest.raises(KeyError):
        MotionClip.from_json(json.dumps({"FrameDuration": 0.1, "Frames": good["Frames"]}))


def test_unsupported_topology_is_rejected(tmp_path):
    text = open(common.URDF["laikago"]).read().replace('name="jtoeFR" type="fixed"', 'name="jtoeFR" type="prismatic"')
    d = tmp_path / "laikago"; d.mkdir()
    (d / "laikago_toes_limits.urdf").write_text(text)
    with pytest.raises(ValueError, match="Only support 1D and 0D joints"):
        RobotModel("laikago", urdf_search_path=str(tmp_path))
