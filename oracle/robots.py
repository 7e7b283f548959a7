"""Per-robot constants, restated from the reference robot modules.

TEST INFRASTRUCTURE (see oracle/__init__.py).
Sources: OpenRoboRL/envs/quadruped_robot/robots/laikago.py:23-71,
OpenRoboRL/envs/quadruped_robot/robots/mini_cheetah.py:23-72,
OpenRoboRL/envs/quadruped_robot/robots/minitaur.py:92-97 (name check), :812-851 (link id sets).
"""
import re
import numpy as np


class RobotSpec(object):
    pass


def _laikago():
    s = RobotSpec()
    s.name = "laikago"
    s.urdf = "laikago/laikago_toes_limits.urdf"
    s.time_step = 0.001
    s.action_repeat = 33
    s.ctrl_latency = 0.002
    s.motor_names = [
        "FR_hip_motor_2_chassis_joint", "FR_upper_leg_2_hip_motor_joint", "FR_lower_leg_2_upper_leg_joint",
        "FL_hip_motor_2_chassis_joint", "FL_upper_leg_2_hip_motor_joint", "FL_lower_leg_2_upper_leg_joint",
        "RR_hip_motor_2_chassis_joint", "RR_upper_leg_2_hip_motor_joint", "RR_lower_leg_2_upper_leg_joint",
        "RL_hip_motor_2_chassis_joint", "RL_upper_leg_2_hip_motor_joint", "RL_lower_leg_2_upper_leg_joint"]
    s.pattern = [re.compile(r"\w+_chassis_\w+"), re.compile(r"\w+_hip_motor_\w+"),
                 re.compile(r"\w+_lower_leg_\w+"), re.compile(r"jtoe\d*")]
    s.init_pos = np.array([0, 0, 0.48])
    s.init_quat = np.array([0.5, 0.5, 0.5, 0.5])
    s.joint_dir = np.array([-1, 1, 1, 1, 1, 1, -1, 1, 1, 1, 1, 1], dtype=np.float64)
    s.joint_offset = np.array([0.0, -0.6, 0.66] * 4)
    s.init_motor = np.array([0, 0.67, -1.25] * 4)
    s.kp = np.array([220.0, 220.0, 220.0] * 4)
    s.kd = np.array([0.3, 2.0, 2.0] * 4)
    s.max_angle_change = 0.2
    return s


def _mini_cheetah():
    s = RobotSpec()
    s.name = "mini_cheetah"
    s.urdf = "mini_cheetah/mini_cheetah.urdf"
    s.time_step = 0.001
    s.action_repeat = 33
    s.ctrl_latency = 0.002
    s.motor_names = [
        "torso_to_abduct_fl_j", "abduct_fl_to_thigh_fl_j", "thigh_fl_to_knee_fl_j",
        "torso_to_abduct_hl_j", "abduct_hl_to_thigh_hl_j", "thigh_hl_to_knee_hl_j",
        "torso_to_abduct_fr_j", "abduct_fr_to_thigh_fr_j", "thigh_fr_to_knee_fr_j",
        "torso_to_abduct_hr_j", "abduct_hr_to_thigh_hr_j", "thigh_hr_to_knee_hr_j"]
    s.pattern = [re.compile(r"torso_"), re.compile(r"abduct_"), re.compile(r"thigh_"), re.compile(r"toe_")]
    s.init_pos = np.array([0, 0, 0.28])
    s.init_quat = np.array([0.0, 0.0, 0.0, 1.0])
    s.joint_dir = np.ones(12)
    s.joint_offset = np.zeros(12)
    s.init_motor = np.array([0, -0.78, 1.74] * 4)
    s.kp = np.array([80.0, 80.0, 80.0] * 4)
    s.kd = np.array([0.1, 1.0, 1.0] * 4)
    s.max_angle_change = 0.2
    return s


def get_spec(name_robot):
    if name_robot == "laikago":
        return _laikago()
    if name_robot == "mini_cheetah":
        return _mini_cheetah()
    raise ValueError("wrong robot select")  # minitaur.py:96-97


def classify_links(spec, joint_names):
    """minitaur.py:812-851.  Returns dict of sorted id lists."""
    chassis, motor, knee, foot = [-1], [], [], []
    for i, name in enumerate(joint_names):
        if spec.pattern[0].match(name):
            chassis.append(i)
        elif spec.pattern[1].match(name):
            motor.append(i)
        elif spec.pattern[2].match(name):
            knee.append(i)
        elif spec.pattern[3].match(name):
            foot.append(i)
        else:
            raise ValueError("Unknown category of joint %s" % name)
    leg = sorted(knee + foot)
    foot = sorted(foot + knee)
    return dict(chassis=sorted(chassis), motor=sorted(motor), knee=knee, foot=foot, leg=leg)
