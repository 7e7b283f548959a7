"""ctypes mirror of include/orl.h and the loader of the CUDA shared library.

There is deliberately no fallback: if `csrc/libopenroborl_b200.so` has not been built
(`python -c "import __graft_entry__ as g; g.build()"`), importing the product fails loudly.
"""
import ctypes as C
import os

ORL_ABI_VERSION = 3
NUM_LEGS, LEG_DOF, NUM_MOTORS = 4, 3, 12
POSE_SIZE, VEL_SIZE, OBS_SIZE, NUM_TAR_FRAMES = 19, 18, 160, 4
MAX_CP, MAX_LINKS, CYCLE_TABLE = 24, 32, 128

MOTOR_POSITION, MOTOR_TORQUE, MOTOR_HYBRID = 1, 2, 3
ORL_OK, ORL_ERR_INVALID, ORL_ERR_CUDA, ORL_ERR_STATE, ORL_ERR_ASSERT = 0, -1, -2, -3, -4


class OrlConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("num_robot", C.c_int32), ("device", C.c_int32), ("gid0", C.c_int32),
        ("seed", C.c_uint64),
        ("action_repeat", C.c_int32), ("max_episode_steps", C.c_int32), ("auto_reset", C.c_int32),
        ("enable_cycle_sync", C.c_int32),
        ("sim_time_step", C.c_double), ("env_time_step", C.c_double), ("warmup_time", C.c_double),
        ("ref_state_init_prob", C.c_double),
        ("tar_frame_steps", C.c_int32 * NUM_TAR_FRAMES),
        ("pose_weight", C.c_float), ("velocity_weight", C.c_float), ("end_effector_weight", C.c_float),
        ("root_pose_weight", C.c_float), ("root_velocity_weight", C.c_float),
        ("pose_err_scale", C.c_float), ("velocity_err_scale", C.c_float), ("end_effector_err_scale", C.c_float),
        ("end_effector_height_err_scale", C.c_float), ("root_pose_err_scale", C.c_float),
        ("root_velocity_err_scale", C.c_float),
        ("dist_fail_threshold", C.c_float), ("rot_fail_threshold", C.c_float),
        ("gravity_z", C.c_float),
        ("contact_kn", C.c_float), ("contact_bn", C.c_float), ("contact_d0", C.c_float), ("contact_bt", C.c_float),
        ("limit_k", C.c_float), ("limit_d", C.c_float), ("limit_e0", C.c_float),
        ("max_coordinate_velocity", C.c_float),
        ("enable_randomizer", C.c_int32), ("rand_lo", C.c_float * 8), ("rand_hi", C.c_float * 8),
        ("joint_friction_k", C.c_float), ("mu_plane", C.c_float),
        ("reward_weight", C.c_float), ("enable_rand_init_time", C.c_int32),
        ("clip_time_min", C.c_double), ("clip_time_max", C.c_double), ("perturb_init_state_prob", C.c_double),
        ("enable_tar_obs_noise", C.c_int32), ("tar_obs_noise", C.c_float),
    ]


class OrlContactPoint(C.Structure):
    _fields_ = [("local", C.c_float * 3), ("radius", C.c_float), ("mu", C.c_float), ("margin", C.c_float),
                ("body", C.c_int32), ("is_foot", C.c_int32)]


class OrlLegModel(C.Structure):
    _fields_ = [
        ("R_fix", (C.c_float * 9) * LEG_DOF), ("p_fix", (C.c_float * 3) * LEG_DOF),
        ("mass", C.c_float * LEG_DOF), ("com", (C.c_float * 3) * LEG_DOF), ("inertia", (C.c_float * 6) * LEG_DOF),
        ("inertia_par", (C.c_float * 6) * LEG_DOF), ("rand_class", C.c_int32 * LEG_DOF), ("friction_draw", C.c_int32 * LEG_DOF),
        ("lower", C.c_float * LEG_DOF), ("upper", C.c_float * LEG_DOF), ("damping", C.c_float * LEG_DOF),
        ("dir", C.c_float * LEG_DOF), ("offset", C.c_float * LEG_DOF), ("kp", C.c_float * LEG_DOF),
        ("kd", C.c_float * LEG_DOF), ("init_angle", C.c_float * LEG_DOF),
        ("motor", C.c_int32 * LEG_DOF),
        ("n_ee", C.c_int32), ("ee_body", C.c_int32 * 2), ("ee_local", (C.c_float * 3) * 2),
        ("n_cp", C.c_int32), ("cp", OrlContactPoint * MAX_CP),
    ]


class OrlRobotModel(C.Structure):
    _fields_ = [
        ("base_mass", C.c_float), ("base_inertia", C.c_float * 6), ("init_pos", C.c_float * 3),
        ("init_quat", C.c_float * 4), ("max_angle_change", C.c_float),
        ("leg", OrlLegModel * NUM_LEGS),
        ("n_links", C.c_int32), ("link_leg", C.c_int32 * MAX_LINKS), ("link_body", C.c_int32 * MAX_LINKS),
        ("link_com", (C.c_float * 3) * MAX_LINKS),
    ]


class OrlStateLayout(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "stride", "pos", "quat", "q", "lin", "ang", "qd", "filt_x", "filt_y", "prev_filtered", "ring1", "ring2",
        "imu_hist", "act_hist", "ang_hist", "ref_pose", "ref_vel", "origin_pos", "origin_rot", "time_offset",
        "prev_phase", "clip", "warmup", "substep_counter", "step_counter", "env_step_counter", "episode",
        "ep_return", "obs_quat", "motor_strength", "rand_params", "clip_change_time", "clip_changes")]


# every symbol include/orl.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "orl_create": (C.c_int, [C.POINTER(OrlConfig), C.POINTER(OrlRobotModel), C.POINTER(_P)]),
    "orl_destroy": (C.c_int, [_P]),
    "orl_last_error": (C.c_char_p, [_P]),
    "orl_abi_version": (C.c_int, []),
    "orl_load_motions": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32]),
    "orl_get_motion_tables": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "orl_reset": (C.c_int, [_P, _P, _P, _P]),
    "orl_step": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "orl_reset_host": (C.c_int, [_P, _P, _P]),
    "orl_step_host": (C.c_int, [_P, _P, _P, _P, _P]),
    "orl_set_max_episode_steps": (C.c_int, [_P, C.c_int32]),
    "orl_set_ref_state_init_prob": (C.c_int, [_P, C.c_double]),
    "orl_state_layout_get": (C.c_int, [_P, C.POINTER(OrlStateLayout)]),
    "orl_state_ptr": (C.c_int, [_P, C.POINTER(_P)]),
    "orl_get_state_host": (C.c_int, [_P, _P]),
    "orl_set_state_host": (C.c_int, [_P, _P]),
    "orl_stats": (C.c_int, [_P, _P, C.c_int32]),
    "orl_motion_lookup": (C.c_int, [_P, _P, _P, C.c_int32, _P, _P, _P, _P, _P, _P, _P]),
    "orl_fk": (C.c_int, [_P, _P, C.c_int32, _P, _P]),
    "orl_get_latency_rings_host": (C.c_int, [_P, _P, _P]),
    "orl_set_latency_rings_host": (C.c_int, [_P, _P, _P]),
    "orl_set_dynamics_rand": (C.c_int, [_P, _P]),
    "orl_mass_matrix": (C.c_int, [_P, _P, C.c_int32, _P, _P]),
    "orl_inverse_dynamics": (C.c_int, [_P, _P, _P, _P, C.c_int32, _P, _P]),
    "orl_forward_dynamics": (C.c_int, [_P, _P, _P, _P, C.c_int32, _P, _P]),
    "orl_reward_terms": (C.c_int, [_P, _P, _P]),
    "orl_launch_count": (C.c_int64, [_P]),
    "orl_fp32_peak": (C.c_int, [C.c_int32, C.c_int32, _P]),
    "orl_motor_convert_to_torque": (C.c_int, [C.c_int32, _P, _P, _P, _P, _P, _P, _P, C.c_int32, _P, _P]),
    "orl_gae": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32, C.c_int32, _P, _P, _P]),
}

# ORL_LIB lets experiments load an alternative build of the same library (kernel-variant A/B runs)
LIB_PATH = os.environ.get("ORL_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libopenroborl_b200.so")
_lib = None


def load_library():
    """Load the CUDA library; raise (never fall back) when it is missing or a symbol is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "openroborl_b200: %s is missing. Build it with `python -c \"import __graft_entry__ as g; g.build()\"`; "
            "there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.orl_abi_version() != ORL_ABI_VERSION:
        raise ImportError("openroborl_b200: ABI mismatch (%d != %d)" % (lib.orl_abi_version(), ORL_ABI_VERSION))
    _lib = lib
    return lib
