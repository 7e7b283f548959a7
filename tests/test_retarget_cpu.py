"""Host-side tooling and the reference-held geometry.

* `openroborl_b200.retarget` against the reference's own artefact: retargeting `laikago_trot` reproduces the shipped
  `minicheetah_trot.txt` (the output of R/envs/quadruped_robot/task/motions/trans2minicheetah.m) to the 5 decimals it was
  printed with, all 19 columns.
* The analytic 3-link leg model the reference holds (trans2minicheetah.m:3-5, 28-30, 146-161; trans_data.py:36-65) against
  the URDF fixtures: hip -> toe of the mini-cheetah fixture is that model exactly; femur and tibia of the Laikago fixture
  agree with its link lengths to a millimetre.  This is the one piece of FK the reference states in closed form, i.e. the
  available substitute for `getLinkState` while PyBullet cannot be installed.
* An inertial frame with a non-identity `rpy` goes through both URDF readers (real pybullet_data files have them).
"""
import os
import re

import numpy as np
import pytest

from oracle.rbd import UrdfTree
from tests import common
from openroborl_b200 import retarget as rt
from openroborl_b200.motion import load_clip_library

SIDE = [-1.0, 1.0, -1.0, 1.0]            # FR, FL, RR, RL (trans2minicheetah.m:13-17)


def test_retargeting_reproduces_the_shipped_minicheetah_clip():
    src, want = load_clip_library(common.CLIPS, ["laikago_trot", "minicheetah_trot"])
    got = rt.retarget_clip(src)
    assert got.frames.shape == want.frames.shape == (33, 19)
    assert got.frame_duration == want.frame_duration and got.loop_mode == want.loop_mode
    assert np.abs(got.frames - want.frames).max() <= 5.1e-6          # the reference printed "%6.5f"
    q = got.frames[:, 3:7]
    np.testing.assert_allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-6)


def test_leg_ik_inverts_leg_fk():
    """The self-check loop of trans_data.py:96-109 (angles in the first quadrant, knee bent backwards)."""
    rng = np.random.default_rng(0)
    for coxa, femur, tibia in (rt.LAIKAGO_LEG, rt.MINI_CHEETAH_LEG, (0.0671, 0.206, 0.185)):
        ang = rng.uniform(0.05, 1.5, size=(2000, 3)) * np.array([1.0, 1.0, -1.0])
        for side in (-1.0, 1.0):
            back = rt.leg_ik(rt.leg_fk(ang, coxa, femur, tibia, side), coxa, femur, tibia, side)
            assert np.abs(back - ang).max() < 1e-7


def test_minicheetah_fixture_is_the_reference_leg_model():
    tree = UrdfTree(common.URDF["mini_cheetah"])
    rng = np.random.default_rng(1)
    n = 200
    q = rng.uniform(-1.0, 1.0, size=(n, 12))
    k = tree.fk(np.zeros((n, 3)), np.tile([0.0, 0.0, 0.0, 1.0], (n, 1)), q)
    for l in range(4):
        hip, toe = 1 + 4 * l, 4 + 4 * l                              # abduct_*, toe_* (URDF order fr, fl, hr, hl)
        rel = k["x"][:, toe] - k["x"][:, hip]
        ang = q[:, 3 * l:3 * l + 3] * rt.MINI_SIGNS[3 * l:3 * l + 3]
        want = rt.leg_fk(ang, *rt.MINI_CHEETAH_LEG, SIDE[l])
        assert np.abs(rel - want).max() <= 1e-9


def test_laikago_fixture_link_lengths_match_the_reference_leg_model():
    tree = UrdfTree(common.URDF["laikago"])
    k = tree.fk(np.zeros((1, 3)), np.array([[0.0, 0.0, 0.0, 1.0]]), np.zeros((1, 12)))
    x = k["x"][0]
    coxa, femur, tibia = rt.LAIKAGO_LEG
    for l in range(4):
        thigh, shank, toe = 2 + 4 * l, 3 + 4 * l, 4 + 4 * l
        assert abs(np.linalg.norm(x[shank] - x[thigh]) - femur) < 1.5e-3
        assert abs(np.linalg.norm(x[toe] - x[shank]) - tibia) < 1.5e-3


def _rotated_inertial_urdf(src, dst, link, rpy):
    """Copy of a URDF in which `link`'s inertial frame is rotated by `rpy` and its tensor re-expressed in that frame:
    the same physical link written the way pybullet_data files write theirs."""
    from oracle import xform as xf
    text = open(src).read()
    m = re.search(r'(<link name="%s">.*?<inertial><origin rpy=")([^"]*)(" xyz="[^"]*"/><mass value="[^"]*"/>\s*<inertia )([^/]*)(/>)' % link, text, re.S)
    assert m, "fixture layout changed"
    vals = dict(re.findall(r'(i[xyz]{2})="([^"]*)"', m.group(4)))
    I = np.array([[float(vals["ixx"]), float(vals["ixy"]), float(vals["ixz"])],
                  [float(vals["ixy"]), float(vals["iyy"]), float(vals["iyz"])],
                  [float(vals["ixz"]), float(vals["iyz"]), float(vals["izz"])]])
    R = xf.rpy_to_matrix(np.asarray(rpy, dtype=np.float64))
    Ir = R.T @ I @ R                                  # tensor in the rotated inertial frame
    new = 'ixx="%.12g" ixy="%.12g" ixz="%.12g" iyy="%.12g" iyz="%.12g" izz="%.12g"' % (Ir[0, 0], Ir[0, 1], Ir[0, 2], Ir[1, 1], Ir[1, 2], Ir[2, 2])
    text = text[:m.start()] + m.group(1) + "%.12g %.12g %.12g" % tuple(rpy) + m.group(3) + new + m.group(5) + text[m.end():]
    open(dst, "w").write(text)


@pytest.mark.parametrize("link", ["thigh_fr", "shank_hl"])
def test_inertial_rpy_is_honoured_by_both_urdf_readers(tmp_path, link):
    from openroborl_b200.model import RobotModel
    from openroborl_b200 import robots
    src = common.URDF["mini_cheetah"]
    dst = str(tmp_path / robots.get("mini_cheetah").URDF_FILENAME)      # pybullet_data-relative name under the search path
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    _rotated_inertial_urdf(src, dst, link, (0.4, -0.7, 1.1))
    a, b = UrdfTree(src), UrdfTree(dst)
    rng = np.random.default_rng(2)
    q = rng.uniform(-1, 1, size=(4, 12))
    quat = rng.normal(size=(4, 4)); quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    Ma = a.mass_matrix(np.zeros((4, 3)), quat, q)
    Mb = b.mass_matrix(np.zeros((4, 3)), quat, q)
    assert np.abs(Ma - Mb).max() <= 1e-9 * np.abs(Ma).max()           # the same physical robot
    ma, mb = RobotModel("mini_cheetah"), RobotModel("mini_cheetah", urdf_search_path=str(tmp_path))
    assert os.path.samefile(mb.urdf_path, dst)
    for l in range(4):
        np.testing.assert_allclose(np.array(mb.c.leg[l].inertia), np.array(ma.c.leg[l].inertia), atol=1e-8)
        np.testing.assert_allclose(np.array(mb.c.leg[l].com), np.array(ma.c.leg[l].com), atol=1e-9)
