"""Self-checks of the oracle's rigid-body model (no Bullet available: physical invariants instead)."""
import numpy as np

from tests import common
from oracle import philox
from oracle.rbd import UrdfTree


def _rand(rng, n):
    quat = rng.normal(size=(n, 4)); quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    return rng.normal(size=(n, 3)), quat, rng.uniform(-1, 1, size=(n, 12)), rng.normal(size=(n, 18))


def test_mass_matrix_is_spd_and_consistent_with_inverse_dynamics():
    tree = UrdfTree(common.URDF["laikago"])
    rng = np.random.default_rng(0)
    pos, quat, q, nu = _rand(rng, 3)
    M = tree.mass_matrix(pos, quat, q)
    assert np.abs(M - M.transpose(0, 2, 1)).max() < 1e-12 and np.linalg.eigvalsh(M).min() > 0
    assert abs(M[0, 3, 3] - tree.mass.sum()) < 1e-9        # total mass on the linear diagonal
    zero = np.zeros((3, 18))
    base = tree.inverse_dynamics(pos, quat, q, zero, zero)
    for j in range(18):
        e = np.zeros((3, 18)); e[:, j] = 1.0
        np.testing.assert_allclose(tree.inverse_dynamics(pos, quat, q, zero, e) - base, M[:, :, j], atol=1e-10)


def test_free_flight_energy_drift_is_first_order():
    from oracle import xform as xf
    tree = UrdfTree(common.URDF["mini_cheetah"])
    rng = np.random.default_rng(1)
    pos, quat, q, nu = _rand(rng, 2)
    drifts = []
    for h in (1e-3, 2.5e-4):
        s = [pos.copy(), quat.copy(), q.copy(), nu.copy() * 0.5]
        E0 = tree.kinetic_energy(*s) + tree.potential_energy(*s[:3])
        for _ in range(int(0.02 / h)):
            M = tree.mass_matrix(*s[:3]); C = tree.bias(s[0], s[1], s[2], s[3])
            s[3] = s[3] + h * np.linalg.solve(M, -C[..., None])[..., 0]
            dq = xf.quaternion_about_axis(np.linalg.norm(s[3][:, 0:3], axis=1) * h, s[3][:, 0:3])
            s[0] = s[0] + h * s[3][:, 3:6]
            s[1] = xf.quaternion_multiply(dq, s[1]); s[1] /= np.linalg.norm(s[1], axis=1, keepdims=True)
            s[2] = s[2] + h * s[3][:, 6:]
        E1 = tree.kinetic_energy(*s) + tree.potential_energy(*s[:3])
        drifts.append(np.abs((E1 - E0) / E0).max())
    assert drifts[0] < 5e-3 and drifts[1] < drifts[0] / 3


def test_perfect_tracking_gives_unit_reward():
    """imitation_task.py:45-49,350-356: all five exp(0) terms -> 0.5+0.05+0.2+0.15+0.1 = 1.0."""
    env = common.make_oracle("laikago", ["laikago_pace"], 4)
    env.reset()
    assert np.abs(env.reward() - 1.0).max() < 1e-12


def test_standing_is_stable_with_zero_action():
    env = common.make_oracle("laikago", ["laikago_pace"], 1, task=dict(ref_state_init_prob=0.0))
    env.reset()
    for _ in range(12):
        env.step(np.zeros((1, 12)))
    assert 0.44 < env.pos[0, 2] < 0.49 and np.abs(env.qd).max() < 1.0 and np.abs(env.lin).max() < 0.3


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        got = philox.philox4x32(np.array([c], dtype=np.uint32), np.array([k], dtype=np.uint32))[0]
        assert tuple(int(x) for x in got) == want


def test_reset_draws_do_not_depend_on_sharding():
    a = philox.reset_draws(7, np.arange(8), np.zeros(8, dtype=int), 5)
    b = philox.reset_draws(7, np.arange(4, 8), np.zeros(4, dtype=int), 5)
    for x, y in zip(a, b):
        assert np.array_equal(x[4:], y)
    full = common.make_oracle("laikago", ["laikago_pace"], 6, seed=11)
    part = common.make_oracle("laikago", ["laikago_pace"], 3, seed=11, gid0=3)
    assert np.array_equal(full.reset()[3:], part.reset())


def _contact_point_heights(orc, frames):
    k = orc.tree.fk(frames[:, 0:3], frames[:, 3:7], frames[:, 7:])
    hs = []
    for c in range(len(orc.cp_body)):
        b = orc.cp_body[c]
        P = k["x"][:, b] + np.einsum("nij,j->ni", k["R"][:, b], orc.cp_local[c])
        hs.append(P[:, 2] - orc.cp_r[c])
    return np.array(hs)                                    # (points, frames)


def test_reference_clips_put_the_feet_of_the_urdf_fixtures_on_the_ground():
    """A reference-held cross-check of the kinematic model (the URDFs of pybullet_data are not on
    disk, openroborl_b200/assets holds fixtures written from the published geometry): the reference's motion clips were
    retargeted onto the REAL robots, so forward kinematics of their frames with the right link lengths, joint offsets /
    directions and toe radii puts the lowest collision point of every frame on the ground plane, never a body that is
    not a foot link.  Every shipped clip (R/envs/quadruped_robot/task/motions/*.txt), all frames."""
    import os
    z = np.load(os.path.join(common.ROOT, "openroborl_b200", "assets", "motions", "clips.npz"), allow_pickle=True)
    names = sorted(set(k.split("/")[0] for k in z.files) - {"synthetic_pace_clamp"})
    assert len(names) == 11
    for clip in names:
        robot = "mini_cheetah" if clip.startswith("mini") else "laikago"
        orc = common.make_oracle(robot, [clip], 1)
        hs = _contact_point_heights(orc, orc.lib.frames)
        low = hs.min(axis=0)                               # lowest collision point of each frame
        foot = orc.cp_foot
        assert foot[hs.argmin(axis=0)].all(), clip         # it is always a point of a foot link
        assert hs[~foot].min() > 0.05, (clip, hs[~foot].min())          # chassis, hips, thighs stay clear of the ground
        assert abs(np.median(low)) < 0.012, (clip, np.median(low))      # stance feet sit on the plane to ~1 cm
        if clip != "laikago_hopturn":                                   # (the hop has a flight phase: 0.2 m)
            assert low.max() < 0.035, (clip, low.max())
        assert low.min() > -0.09, (clip, low.min())                     # worst frame of the 11 clips: laikago_sidesteps
    # the symmetric gaits used by the bench workloads: within [-1 mm, +20 mm] over the whole clip
    for clip in ("laikago_pace", "laikago_trot"):
        orc = common.make_oracle("laikago", [clip], 1)
        low = _contact_point_heights(orc, orc.lib.frames).min(axis=0)
        assert -0.002 < low.min() and low.max() < 0.021, (clip, low.min(), low.max())


def test_reference_clips_keep_stance_feet_in_place():
    """Second reference-held check of the kinematic model (leg assignment, joint signs, JOINT_DIRECTIONS / JOINT_OFFSETS):
    while a foot of the retargeted clip stands on the ground it must not travel with the trunk.  With two legs swapped or
    a joint axis flipped the stance feet would move at least as fast as the root."""
    for clip, robot in (("laikago_pace", "laikago"), ("laikago_trot", "laikago"), ("minicheetah_trot", "mini_cheetah")):
        orc = common.make_oracle(robot, [clip], 1)
        fr = orc.lib.frames
        k = orc.tree.fk(fr[:, 0:3], fr[:, 3:7], fr[:, 7:])
        P, H = [], []
        for c in np.nonzero(orc.cp_foot)[0]:
            b = orc.cp_body[c]
            p = k["x"][:, b] + np.einsum("nij,j->ni", k["R"][:, b], orc.cp_local[c])
            P.append(p); H.append(p[:, 2] - orc.cp_r[c])
        P, H = np.array(P), np.array(H)
        step = np.linalg.norm(np.diff(P[:, :, :2], axis=1), axis=-1)             # horizontal travel per frame
        stance = (H[:, :-1] < 0.012) & (H[:, 1:] < 0.012)
        root = np.median(np.linalg.norm(np.diff(fr[:, :2], axis=0), axis=-1))
        assert stance.sum() >= 20, clip
        assert np.median(step[stance]) < 0.45 * root, (clip, np.median(step[stance]), root)
        assert np.median(step[~stance]) > 0.7 * root, (clip, np.median(step[~stance]), root)
