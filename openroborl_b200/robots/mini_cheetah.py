"""Mini-cheetah constants (reference OpenRoboRL/envs/quadruped_robot/robots/mini_cheetah.py:23-72)."""
import math
import re
import numpy as np

URDF_FILENAME = "mini_cheetah/mini_cheetah.urdf"

T_STEP = 0.001
NUM_ACTION_REPEAT = 33
CTRL_LATENCY = 0.002

NUM_MOTORS = 12
NUM_LEGS = 4
DOFS_PER_LEG = 3
# motor order is fl, hl, fr, hr -- not the URDF order; mapped by joint name (minitaur.py:860-864)
MOTOR_NAMES = [fmt % (leg, leg) if fmt.count("%s") == 2 else fmt % leg
               for leg in ("fl", "hl", "fr", "hr")
               for fmt in ("torso_to_abduct_%s_j", "abduct_%s_to_thigh_%s_j", "thigh_%s_to_knee_%s_j")]
PATTERN = [re.compile(r"torso_"), re.compile(r"abduct_"), re.compile(r"thigh_"), re.compile(r"toe_")]

INIT_POSITION = [0, 0, 0.28]
INIT_QUAT = [0.0, 0.0, 0.0, 1.0]
JOINT_DIRECTIONS = np.array([1] * 12)
JOINT_OFFSETS = np.array([0.0, 0.0, 0.0] * 4)
INIT_MOTOR_ANGLES = np.array([0, -0.78, 1.74] * NUM_LEGS)

motor_kp = [80.0, 80.0, 80.0] * NUM_LEGS
motor_kd = [0.1, 1.0, 1.0] * NUM_LEGS

OVERHEAT_SHUTDOWN_TORQUE = 2.45
OVERHEAT_SHUTDOWN_TIME = 1.0
MAX_MOTOR_ANGLE_CHANGE_PER_STEP = 0.2
ACTION_BOUND = 2 * math.pi
