"""Quaternion / angle helpers of the reference hot path, float64, batched.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Quaternions are xyzw, shape (..., 4).

Two sources are restated here:

* `pybullet_utils.transformations` (pybullet==2.7.3, third party, NOT under
  /root/reference; call sites listed in SURVEY.md section 8c): quaternion_multiply,
  quaternion_conjugate, quaternion_inverse, quaternion_about_axis,
  quaternion_from_euler (static xyz), quaternion_slerp (shortest path, `_EPS =
  4*eps` guards).  These follow the published transformations.py algorithms.
* `OpenRoboRL/envs/utilities/pose3d.py` of the reference (line numbers cited per
  function).

Bullet's euler helpers (`getEulerFromQuaternion`, `getQuaternionFromEuler`) are
restated from knowledge of Bullet3; unpinned.
"""
import numpy as np

_EPS = np.finfo(float).eps * 4.0


def _a(x):
    return np.asarray(x, dtype=np.float64)


# ----------------------------------------------------------------------------------------------
# pybullet_utils.transformations (restated)
# ----------------------------------------------------------------------------------------------
def quaternion_multiply(q1, q0):
    """Hamilton product q1 (x) q0, xyzw."""
    q1 = _a(q1)
    q0 = _a(q0)
    x0, y0, z0, w0 = q0[..., 0], q0[..., 1], q0[..., 2], q0[..., 3]
    x1, y1, z1, w1 = q1[..., 0], q1[..., 1], q1[..., 2], q1[..., 3]
    return np.stack([
        x1 * w0 + y1 * z0 - z1 * y0 + w1 * x0,
        -x1 * z0 + y1 * w0 + z1 * x0 + w1 * y0,
        x1 * y0 - y1 * x0 + z1 * w0 + w1 * z0,
        -x1 * x0 - y1 * y0 - z1 * z0 + w1 * w0], axis=-1)


def quaternion_conjugate(q):
    q = _a(q)
    return np.concatenate([-q[..., :3], q[..., 3:]], axis=-1)


def quaternion_inverse(q):
    q = _a(q)
    return quaternion_conjugate(q) / np.sum(q * q, axis=-1, keepdims=True)


def quaternion_about_axis(angle, axis):
    angle = _a(angle)
    axis = _a(axis)
    axis = np.broadcast_to(axis, angle.shape + (3,))
    n = np.sqrt(np.sum(axis * axis, axis=-1, keepdims=True))
    v = axis / n * np.sin(angle / 2.0)[..., None]
    return np.concatenate([v, np.cos(angle / 2.0)[..., None]], axis=-1)


def quaternion_from_euler(ai, aj, ak):
    """Static-frame xyz euler (roll about x, then pitch about y, then yaw about z)."""
    ai, aj, ak = _a(ai) / 2.0, _a(aj) / 2.0, _a(ak) / 2.0
    ci, si = np.cos(ai), np.sin(ai)
    cj, sj = np.cos(aj), np.sin(aj)
    ck, sk = np.cos(ak), np.sin(ak)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    return np.stack([cj * sc - sj * cs, cj * ss + sj * cc, cj * cs - sj * sc, cj * cc + sj * ss], axis=-1)


def quaternion_slerp(q0, q1, fraction):
    """transformations.quaternion_slerp(quat0, quat1, fraction, spin=0, shortestpath=True)."""
    q0 = _a(q0)
    q1 = _a(q1)
    fraction = _a(fraction)
    q0 = q0 / np.sqrt(np.sum(q0 * q0, axis=-1, keepdims=True))
    q1 = q1 / np.sqrt(np.sum(q1 * q1, axis=-1, keepdims=True))
    shape = np.broadcast_shapes(q0.shape[:-1], q1.shape[:-1], fraction.shape)
    q0 = np.broadcast_to(q0, shape + (4,)).copy()
    q1 = np.broadcast_to(q1, shape + (4,)).copy()
    fraction = np.broadcast_to(fraction, shape)
    d = np.sum(q0 * q1, axis=-1)
    ret_q0 = (fraction == 0.0)
    ret_q1 = (fraction == 1.0) & ~ret_q0
    ret_q0 = ret_q0 | ((np.abs(np.abs(d) - 1.0) < _EPS) & ~ret_q1)
    neg = d < 0.0
    d = np.where(neg, -d, d)
    q1 = np.where(neg[..., None], -q1, q1)
    angle = np.arccos(np.clip(d, -1.0, 1.0))
    ret_q0 = ret_q0 | ((np.abs(angle) < _EPS) & ~ret_q1)
    safe = np.where(np.abs(angle) < _EPS, 1.0, angle)
    isin = 1.0 / np.sin(safe)
    out = q0 * (np.sin((1.0 - fraction) * safe) * isin)[..., None] \
        + q1 * (np.sin(fraction * safe) * isin)[..., None]
    # the early returns give the *normalised* inputs (q1 before any sign flip is impossible to
    # distinguish: fraction == 1 returns q1 as normalised, un-negated)
    q1_plain = np.where(neg[..., None], -q1, q1)
    out = np.where(ret_q0[..., None], q0, out)
    out = np.where(ret_q1[..., None], q1_plain, out)
    return out


# ----------------------------------------------------------------------------------------------
# reference pose3d.py
# ----------------------------------------------------------------------------------------------
def QuaternionNormalize(q):
    """pose3d.py:85-108."""
    q = _a(q)
    n = np.sqrt(np.sum(q * q, axis=-1, keepdims=True))
    if np.any(np.isclose(n, 0.0)):
        raise ValueError('Quaternion may not be zero in QuaternionNormalize')
    return q / n


def QuaternionRotatePoint(point, quat):
    """pose3d.py:213-231: (q * [p,0]) * q^-1, q^-1 = conj / |q|^2."""
    point = _a(point)
    quat = _a(quat)
    shape = np.broadcast_shapes(point.shape[:-1], quat.shape[:-1])
    point = np.broadcast_to(point, shape + (3,))
    quat = np.broadcast_to(quat, shape + (4,))
    qp = np.concatenate([point, np.zeros(shape + (1,))], axis=-1)
    qi = quaternion_inverse(quat)
    return quaternion_multiply(quaternion_multiply(quat, qp), qi)[..., :3]


def QuaternionToAxisAngle(quat):
    """pose3d.py:139-187 (default_axis = z, no direction axis)."""
    quat = _a(quat)
    nrm = np.sqrt(np.sum(quat * quat, axis=-1))
    if not np.all(np.isclose(1.0, nrm)):
        raise ValueError('Quaternion should have unit length')
    axis = quat[..., :3].copy()
    axis_norm = np.sqrt(np.sum(axis * axis, axis=-1))
    small = axis_norm < 1e-8
    z = np.zeros_like(axis)
    z[..., 2] = 1.0
    axis = np.where(small[..., None], z, axis / np.where(small, 1.0, axis_norm)[..., None])
    angle = np.arctan2(axis_norm, quat[..., 3]) * 2
    return axis, angle


def standardize_quaternion(q):
    """pose3d.py:289-301."""
    q = _a(q)
    return np.where((q[..., 3] < 0)[..., None], -q, q)


def normalize_rotation_angle(theta):
    """pose3d.py:304-322."""
    theta = _a(theta)
    big = np.abs(theta) > np.pi
    f = np.fmod(theta, 2 * np.pi)
    f = np.where(f >= 0, f + -2 * np.pi, f + 2 * np.pi)
    return np.where(big, f, theta)


def calc_heading(q):
    """pose3d.py:325-341."""
    rot_dir = QuaternionRotatePoint(np.array([1.0, 0.0, 0.0]), q)
    return np.arctan2(rot_dir[..., 1], rot_dir[..., 0])


def calc_heading_rot(q):
    """pose3d.py:344-356."""
    return quaternion_about_axis(calc_heading(q), [0, 0, 1])


def MapToMinusPiToPi(angles):
    """pose3d.py:358-374: fmod(a, 2pi); >= pi -> -2pi; < -pi -> +2pi."""
    angles = _a(angles)
    m = np.fmod(angles, 2 * np.pi)
    m = np.where(m >= np.pi, m - 2 * np.pi, np.where(m < -np.pi, m + 2 * np.pi, m))
    return m


# ----------------------------------------------------------------------------------------------
# Bullet euler helpers (knowledge of Bullet3; unpinned)
# ----------------------------------------------------------------------------------------------
def bullet_euler_from_quaternion(q):
    """pybullet.getEulerFromQuaternion: ZYX, gimbal branches at |sarg| >= 0.99999."""
    q = _a(q)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    sqx, sqy, sqz, squ = x * x, y * y, z * z, w * w
    sarg = -2.0 * (x * z - w * y)
    lo = sarg <= -0.99999
    hi = sarg >= 0.99999
    pitch = np.arcsin(np.clip(sarg, -1.0, 1.0))
    roll = np.arctan2(2 * (y * z + w * x), squ - sqx - sqy + sqz)
    yaw = np.arctan2(2 * (x * y + w * z), squ + sqx - sqy - sqz)
    pitch = np.where(lo, -0.5 * np.pi, np.where(hi, 0.5 * np.pi, pitch))
    roll = np.where(lo | hi, 0.0, roll)
    yaw = np.where(lo, 2 * np.arctan2(x, -y), np.where(hi, 2 * np.arctan2(-x, y), yaw))
    return np.stack([roll, pitch, yaw], axis=-1)


def bullet_quaternion_from_euler(rpy):
    """pybullet.getQuaternionFromEuler: btQuaternion::setEulerZYX(yaw, pitch, roll), normalised."""
    rpy = _a(rpy)
    hr, hp, hy = rpy[..., 0] * 0.5, rpy[..., 1] * 0.5, rpy[..., 2] * 0.5
    cr, sr = np.cos(hr), np.sin(hr)
    cp, sp = np.cos(hp), np.sin(hp)
    cy, sy = np.cos(hy), np.sin(hy)
    q = np.stack([sr * cp * cy - cr * sp * sy,
                  cr * sp * cy + sr * cp * sy,
                  cr * cp * sy - sr * sp * cy,
                  cr * cp * cy + sr * sp * sy], axis=-1)
    return q / np.sqrt(np.sum(q * q, axis=-1, keepdims=True))


def quat_to_matrix(q):
    q = _a(q)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z)
    R[..., 0, 1] = 2 * (x * y - z * w)
    R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w)
    R[..., 1, 1] = 1 - 2 * (x * x + z * z)
    R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w)
    R[..., 2, 1] = 2 * (y * z + x * w)
    R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def rpy_to_matrix(rpy):
    """URDF rpy: R = Rz(yaw) Ry(pitch) Rx(roll)."""
    r, p, y = [float(v) for v in rpy]
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def axis_angle_matrix(axis, angle):
    """Rodrigues; axis (3,), angle (...,) -> (..., 3, 3)."""
    axis = _a(axis)
    axis = axis / np.linalg.norm(axis)
    angle = _a(angle)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    s = np.sin(angle)[..., None, None]
    c = np.cos(angle)[..., None, None]
    return np.eye(3) + s * K + (1 - c) * (K @ K)
