"""Laikago -> mini-cheetah motion retargeting (host-side tooling).

Restates the reference's MATLAB script `envs/quadruped_robot/task/motions/trans2minicheetah.m` (the tool that produced
the shipped `minicheetah_trot.txt` from `laikago_trot.txt`) and the analytic leg model of `trans_data.py:36-65`:
* toe positions of the source clip by the closed-form 3-link leg kinematics (coxa / femur / tibia, :146-161) with the
  Laikago joint signs and offsets (:7-8),
* hip spacing corrected by the difference of the coxa lengths, everything scaled by the leg-length ratio (:30, :52-54),
* inverse kinematics on the target leg (:131-144), target joint signs (:31),
* root position scaled and lowered by 0.045 m (:48-49), root orientation re-expressed from the y-up Laikago URDF frame in
  the z-up mini-cheetah frame (`inv * R`, :32-40).
Pinned against the reference's own artefact: `retarget_clip(laikago_trot)` reproduces `minicheetah_trot.txt`
(tests/test_model_cpu.py).
"""
import math

import numpy as np

from .motion import MotionClip

LAIKAGO_LEG = (0.032875, 0.25223, 0.251)          # coxa, femur, tibia  (trans2minicheetah.m:3-5)
MINI_CHEETAH_LEG = (0.062, 0.209, 0.18)           # (:28-30)
LAIKAGO_SIGNS = np.array([-1, 1, 1, 1, 1, 1, -1, 1, 1, 1, 1, 1], dtype=np.float64)     # laikago2KinematicsOffset (:7)
LAIKAGO_OFFSET = np.array([0.0, 0.6, -0.66])                                           # offsetAngle (:8)
MINI_SIGNS = np.array([1, -1, -1, 1, -1, -1, 1, -1, -1, 1, -1, -1], dtype=np.float64)  # kinematics2Mini (:31)
AXIS_SWAP = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [1.0, 0.0, 0.0]])              # inv (:32-34)
ROOT_DROP = 0.045                                                                      # (:49)


def leg_fk(angle, coxa, femur, tibia, side):
    """Toe position relative to the hip (trans2minicheetah.m:146-161); side = +1 / -1."""
    s1, s2, s3 = np.sin(angle[..., 0]), np.sin(angle[..., 1]), np.sin(angle[..., 2])
    c1, c2, c3 = np.cos(angle[..., 0]), np.cos(angle[..., 1]), np.cos(angle[..., 2])
    c23 = c2 * c3 - s2 * s3
    s23 = s2 * c3 + c2 * s3
    p0 = tibia * s23 + femur * s2
    p1 = coxa * side * c1 + tibia * (s1 * c23) + femur * c2 * s1
    p2 = coxa * side * s1 - tibia * (c1 * c23) - femur * c1 * c2
    return np.stack([-p0, p1, p2], axis=-1)


def leg_ik(coord, coxa, femur, tibia, side):
    """Inverse of leg_fk (trans2minicheetah.m:131-144, knee bent backwards)."""
    x, y, z = coord[..., 0], coord[..., 1], coord[..., 2]
    D = (y ** 2 + z ** 2 - coxa ** 2 + x ** 2 - femur ** 2 - tibia ** 2) / (2 * tibia * femur)
    D = np.clip(D, -0.99999999999999, 0.99999999999999)              # checkdomain (:117-129)
    gamma = np.arctan2(-np.sqrt(1 - D ** 2), D)
    reach = np.sqrt(np.maximum(y ** 2 + z ** 2 - coxa ** 2, 0.0))
    tetta = -np.arctan2(z, y) - np.arctan2(reach, side * coxa)
    tetta = np.where(tetta > math.pi, tetta - 2 * math.pi, tetta)
    tetta = np.where(tetta < -math.pi, tetta + 2 * math.pi, tetta)
    alpha = np.arctan2(-x, reach) - np.arctan2(tibia * np.sin(gamma), femur + tibia * np.cos(gamma))
    return np.stack([-tetta, alpha, gamma], axis=-1)


def _quat_to_matrix(q):          # xyzw -> 3x3
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    n = x * x + y * y + z * z + w * w
    s = 2.0 / n
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - s * (y * y + z * z); R[..., 0, 1] = s * (x * y - z * w); R[..., 0, 2] = s * (x * z + y * w)
    R[..., 1, 0] = s * (x * y + z * w); R[..., 1, 1] = 1 - s * (x * x + z * z); R[..., 1, 2] = s * (y * z - x * w)
    R[..., 2, 0] = s * (x * z - y * w); R[..., 2, 1] = s * (y * z + x * w); R[..., 2, 2] = 1 - s * (x * x + y * y)
    return R


def _matrix_to_quat(R):          # 3x3 -> xyzw with w >= 0 (MATLAB rotm2quat convention)
    out = np.empty(R.shape[:-2] + (4,))
    for idx in np.ndindex(R.shape[:-2]):
        m = R[idx]
        K = np.array([[m[0, 0] - m[1, 1] - m[2, 2], 0, 0, 0],
                      [m[0, 1] + m[1, 0], m[1, 1] - m[0, 0] - m[2, 2], 0, 0],
                      [m[0, 2] + m[2, 0], m[1, 2] + m[2, 1], m[2, 2] - m[0, 0] - m[1, 1], 0],
                      [m[2, 1] - m[1, 2], m[0, 2] - m[2, 0], m[1, 0] - m[0, 1], m[0, 0] + m[1, 1] + m[2, 2]]]) / 3.0
        w, v = np.linalg.eigh(K + K.T - np.diag(np.diag(K)))
        q = v[:, np.argmax(w)]               # xyzw
        out[idx] = q if q[3] >= 0 else -q
    return out


def retarget_frames(frames, src_leg=LAIKAGO_LEG, dst_leg=MINI_CHEETAH_LEG):
    """[F, 19] Laikago frames (x y z qx qy qz qw + 12 joints, legs FR FL RR RL) -> mini-cheetah frames."""
    frames = np.asarray(frames, dtype=np.float64)
    F = frames.shape[0]
    c1, f1, t1 = src_leg
    c2, f2, t2 = dst_leg
    coff = (f2 + t2) / (f1 + t1)
    side = np.array([-1.0, 1.0, -1.0, 1.0])                 # legs 2 and 4 (FL, RL) use +1 (:13-17)
    joints = frames[:, 7:19].reshape(F, 4, 3)
    ang = (joints + LAIKAGO_OFFSET) * LAIKAGO_SIGNS.reshape(4, 3)
    pos = np.stack([leg_fk(ang[:, l], c1, f1, t1, side[l]) for l in range(4)], axis=1)          # [F, 4, 3]
    pos[:, [0, 2], 1] -= (c2 - c1)                          # pos(:,[2,8]) (:52)
    pos[:, [1, 3], 1] += (c2 - c1)                          # pos(:,[5,11]) (:53)
    pos = pos * coff
    out_j = np.stack([leg_ik(pos[:, l], c2, f2, t2, side[l]) for l in range(4)], axis=1)
    out_j = out_j.reshape(F, 12) * MINI_SIGNS
    out = frames.copy()
    out[:, 0:3] = frames[:, 0:3] * coff
    out[:, 2] -= ROOT_DROP
    # The script hands the clip's xyzw quaternions to quat2rotm, which reads them as wxyz (:35-36): reproduced, the shipped
    # minicheetah_trot.txt is the output of exactly that (w' = qx, x' = qy, y' = qz, z' = qw).
    q = frames[:, 3:7]
    R = _quat_to_matrix(np.stack([q[:, 1], q[:, 2], q[:, 3], q[:, 0]], axis=1))
    out[:, 3:7] = _matrix_to_quat(AXIS_SWAP @ R)
    out[:, 7:19] = out_j
    return out


def retarget_clip(clip, name="retargeted"):
    """MotionClip of a Laikago motion -> MotionClip for the mini-cheetah (same timing and loop flags)."""
    return MotionClip(retarget_frames(clip.frames), clip.frame_duration, clip.loop_mode, clip.cycle_pos, clip.cycle_rot, name=name)
