"""Robot constant sets of the two quadrupeds the reference supports.

The reference keeps them as module-level names in `OpenRoboRL/envs/quadruped_robot/robots/laikago.py:23-71` and
`mini_cheetah.py:23-72`; the classes that consume them only use attribute access, so here each set is one
`RobotSpec` instance built from a compact table (`robots.laikago`, `robots.mini_cheetah`, `robots.get(name)`).
Attribute names are the reference's, because `compat.Minitaur` and user code read them by name.
"""
import math
import re

import numpy as np

LEGS = 4
JOINTS_PER_LEG = 3


class RobotSpec(object):
    """Attribute bag with the reference's constant names."""

    def __init__(self, urdf, motor_names, link_class_patterns, stand_height, stand_quat, direction, offset_per_leg,
                 init_angles_per_leg, kp, kd_per_leg):
        per_leg = lambda triple: np.array(list(triple) * LEGS, dtype=np.float64)
        self.URDF_FILENAME = urdf
        # one control step = 33 physics steps of 1 ms; the controller sees motor angles 2 ms late
        self.T_STEP, self.NUM_ACTION_REPEAT, self.CTRL_LATENCY = 1e-3, 33, 2e-3
        self.NUM_LEGS, self.DOFS_PER_LEG, self.NUM_MOTORS = LEGS, JOINTS_PER_LEG, LEGS * JOINTS_PER_LEG
        self.MOTOR_NAMES = list(motor_names)
        self.PATTERN = [re.compile(p) for p in link_class_patterns]      # chassis / hip motor / knee / toe (minitaur.py:825-849)
        self.INIT_POSITION = [0, 0, stand_height]
        self.INIT_QUAT = list(stand_quat)
        self.JOINT_DIRECTIONS = np.array(direction)
        self.JOINT_OFFSETS = per_leg(offset_per_leg)
        self.INIT_MOTOR_ANGLES = per_leg(init_angles_per_leg)
        self.motor_kp = [float(kp)] * self.NUM_MOTORS
        self.motor_kd = [float(v) for v in kd_per_leg] * LEGS
        self.OVERHEAT_SHUTDOWN_TORQUE, self.OVERHEAT_SHUTDOWN_TIME = 2.45, 1.0       # unused: overheat protection is off
        self.MAX_MOTOR_ANGLE_CHANGE_PER_STEP = 0.2
        self.ACTION_BOUND = 2 * math.pi


def _laikago_motors():
    parts = ("hip_motor_2_chassis", "upper_leg_2_hip_motor", "lower_leg_2_upper_leg")
    return ["%s_%s_joint" % (leg, part) for leg in ("FR", "FL", "RR", "RL") for part in parts]


def _cheetah_motors():
    # motor order is fl, hl, fr, hr -- NOT the URDF order; motors are mapped by joint name (minitaur.py:860-864)
    out = []
    for leg in ("fl", "hl", "fr", "hr"):
        out += ["torso_to_abduct_%s_j" % leg, "abduct_%s_to_thigh_%s_j" % (leg, leg), "thigh_%s_to_knee_%s_j" % (leg, leg)]
    return out


# y-up URDF: the standing orientation is the quaternion (0.5, 0.5, 0.5, 0.5); right-side hips turn the other way
laikago = RobotSpec(
    urdf="laikago/laikago_toes_limits.urdf", motor_names=_laikago_motors(),
    link_class_patterns=(r"\w+_chassis_\w+", r"\w+_hip_motor_\w+", r"\w+_lower_leg_\w+", r"jtoe\d*"),
    stand_height=0.48, stand_quat=(0.5, 0.5, 0.5, 0.5),
    direction=[-1 if (m % 6) == 0 else 1 for m in range(12)],
    offset_per_leg=(0.0, -0.6, 0.66), init_angles_per_leg=(0, 0.67, -1.25), kp=220.0, kd_per_leg=(0.3, 2.0, 2.0))

mini_cheetah = RobotSpec(
    urdf="mini_cheetah/mini_cheetah.urdf", motor_names=_cheetah_motors(),
    link_class_patterns=(r"torso_", r"abduct_", r"thigh_", r"toe_"),
    stand_height=0.28, stand_quat=(0.0, 0.0, 0.0, 1.0),
    direction=[1] * 12,
    offset_per_leg=(0.0, 0.0, 0.0), init_angles_per_leg=(0, -0.78, 1.74), kp=80.0, kd_per_leg=(0.1, 1.0, 1.0))


def get(name_robot):
    """Minitaur.__init__ robot selection (reference minitaur.py:92-97).

    Also accepts "minicheetah": run.py:57 derives that spelling from the task name and the reference
    then fails with ValueError("wrong robot select").
    """
    if name_robot == "laikago":
        return laikago
    if name_robot in ("mini_cheetah", "minicheetah"):
        return mini_cheetah
    raise ValueError("wrong robot select")
