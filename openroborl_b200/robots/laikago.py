"""Laikago constants (reference OpenRoboRL/envs/quadruped_robot/robots/laikago.py:23-71)."""
import math
import re
import numpy as np

URDF_FILENAME = "laikago/laikago_toes_limits.urdf"

T_STEP = 0.001
NUM_ACTION_REPEAT = 33
CTRL_LATENCY = 0.002

NUM_MOTORS = 12
NUM_LEGS = 4
DOFS_PER_LEG = 3
MOTOR_NAMES = ["%s_%s_joint" % (leg, part)
               for leg in ("FR", "FL", "RR", "RL")
               for part in ("hip_motor_2_chassis", "upper_leg_2_hip_motor", "lower_leg_2_upper_leg")]
# chassis / motor(hip) / knee / foot joint-name classes (minitaur.py:825-849)
PATTERN = [re.compile(r"\w+_chassis_\w+"), re.compile(r"\w+_hip_motor_\w+"),
           re.compile(r"\w+_lower_leg_\w+"), re.compile(r"jtoe\d*")]

INIT_POSITION = [0, 0, 0.48]
INIT_QUAT = [0.5, 0.5, 0.5, 0.5]
JOINT_DIRECTIONS = np.array([-1, 1, 1, 1, 1, 1, -1, 1, 1, 1, 1, 1])
JOINT_OFFSETS = np.array([0.0, -0.6, 0.66] * 4)
INIT_MOTOR_ANGLES = np.array([0, 0.67, -1.25] * NUM_LEGS)

motor_kp = [220.0, 220.0, 220.0] * NUM_LEGS
motor_kd = [0.3, 2.0, 2.0] * NUM_LEGS

OVERHEAT_SHUTDOWN_TORQUE = 2.45
OVERHEAT_SHUTDOWN_TIME = 1.0
MAX_MOTOR_ANGLE_CHANGE_PER_STEP = 0.2
ACTION_BOUND = 2 * math.pi
