"""MotorModel.convert_to_torque of the reference, all control modes, batched NumPy float64.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows R/envs/quadruped_robot/robots/minitaur_motor.py:104-173; pinned to the
reference class by tests/golden/ref_motion_kat.npz (motor/* arrays, scripts/make_golden.py).
"""
import numpy as np

NUM_MOTORS = 12
MOTOR_COMMAND_DIMENSION = 5                                                             # :24
POSITION_INDEX, POSITION_GAIN_INDEX, VELOCITY_INDEX, VELOCITY_GAIN_INDEX, TORQUE_INDEX = 0, 1, 2, 3, 4   # :27-31


def convert_to_torque(mode, commands, motor_angle, motor_velocity, kp, kd, strength=None, torque_limits=None):
    """mode: 'position' | 'torque' | 'hybrid'; commands (n, 12) or (n, 60); returns the motor torques (n, 12)."""
    commands = np.asarray(commands, dtype=np.float64)
    strength = np.ones(NUM_MOTORS) if strength is None else np.asarray(strength, dtype=np.float64)
    if mode == "pwm":
        raise ValueError("pwm is not a supported motor control mode")                    # :133-135
    if mode == "torque":                                                                 # :137-140: no gains, no clipping
        assert commands.shape[-1] == NUM_MOTORS
        return strength * commands
    additional = np.full(NUM_MOTORS, 0)
    if mode == "position":                                                               # :147-152
        assert commands.shape[-1] == NUM_MOTORS
        des_q, des_qd = commands, np.full(NUM_MOTORS, 0)
    elif mode == "hybrid":                                                               # :153-162
        assert commands.shape[-1] == MOTOR_COMMAND_DIMENSION * NUM_MOTORS
        kp = commands[..., POSITION_GAIN_INDEX::MOTOR_COMMAND_DIMENSION]
        kd = commands[..., VELOCITY_GAIN_INDEX::MOTOR_COMMAND_DIMENSION]
        des_q = commands[..., POSITION_INDEX::MOTOR_COMMAND_DIMENSION]
        des_qd = commands[..., VELOCITY_INDEX::MOTOR_COMMAND_DIMENSION]
        additional = commands[..., TORQUE_INDEX::MOTOR_COMMAND_DIMENSION]
    else:
        raise ValueError(mode)
    tau = -1 * (kp * (motor_angle - des_q)) - kd * (motor_velocity - des_qd) + additional   # :163-164
    tau = strength * tau                                                                  # :165
    if torque_limits is not None:                                                         # :166-171
        lim = np.asarray(torque_limits, dtype=np.float64)
        tau = np.clip(tau, -1.0 * lim, lim)
    return tau
