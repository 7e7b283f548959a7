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
