"""Tensor-native batched motion-imitation environment (thin host wrapper over the C ABI).

PyTorch is used only for device memory, streams and zero-copy tensor handoff; all arithmetic of the
env step happens in the CUDA library (include/orl.h).  The list-based drop-in for the reference's
`WrapperEnv` lives in `openroborl_b200.compat`.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _abi
from .model import RobotModel
from .motion import MotionClip

# reference defaults: imitation_task.py:45-58 as configured by run.py:58-64
TASK_DEFAULTS = dict(
    pose_weight=0.5, velocity_weight=0.05, end_effector_weight=0.2, root_pose_weight=0.15, root_velocity_weight=0.1,
    pose_err_scale=5.0, velocity_err_scale=0.1, end_effector_err_scale=40.0, end_effector_height_err_scale=3.0,
    root_pose_err_scale=20.0, root_velocity_err_scale=2.0, tar_frame_steps=(1, 2, 10, 30), ref_state_init_prob=0.9,
    warmup_time=0.25, enable_cycle_sync=True, dist_fail_threshold=1.0, rot_fail_threshold=0.5 * math.pi,
    # options run.py leaves at their defaults (imitation_task.py:36-57); any other value selects the option kernels
    weight=1.0, enable_rand_init_time=True, clip_time_min=math.inf, clip_time_max=math.inf, perturb_init_state_prob=0.0,
    tar_obs_noise=None)
# physics model of this repo (DESIGN.md "physics model"); gravity from quadruped_gym_env.py:200
PHYS_DEFAULTS = dict(gravity_z=-10.0, contact_kn=1.0e5, contact_bn=500.0, contact_d0=0.002, contact_bt=3000.0,
                     limit_k=1000.0, limit_d=2.0, limit_e0=0.01, mu_plane=1.0, margin_factor=0.02,
                     max_coordinate_velocity=100.0, joint_friction_k=0.5)
# minitaur_env_randomizer_config.py:22-36 ("all_params"), keys in the reference's sorted sampling order
RANDOMIZER_DEFAULTS = {"battery": (14.0, 16.8), "inertia": (0.5, 1.5), "joint friction": (0.0, 0.05), "latency": (0.0, 0.04),
                       "lateral friction": (0.5, 1.25), "mass": (0.8, 1.2), "motor friction": (0.0, 0.05),
                       "motor strength": (0.8, 1.2)}

_INT_FIELDS = ("clip", "warmup", "substep_counter", "step_counter", "env_step_counter", "episode", "clip_changes")
_F64_FIELDS = ("time_offset", "prev_phase", "clip_change_time")
_FIELD_SHAPES = dict(
    pos=(3,), quat=(4,), q=(12,), lin=(3,), ang=(3,), qd=(12,), filt_x=(2, 12), filt_y=(2, 12), prev_filtered=(12,),
    ring1=(12,), ring2=(12,), imu_hist=(3, 4), act_hist=(3, 12), ang_hist=(3, 12), ref_pose=(19,), ref_vel=(18,),
    origin_pos=(3,), origin_rot=(4,), time_offset=(), prev_phase=(), clip=(), warmup=(), substep_counter=(),
    step_counter=(), env_step_counter=(), episode=(), ep_return=(), obs_quat=(4,), motor_strength=(12,),
    rand_params=(10,), clip_change_time=(), clip_changes=())


class OrlError(RuntimeError):
    pass


class _DevView(object):
    """__cuda_array_interface__ holder so torch can wrap library-owned device memory without a copy."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=3, strides=None)
        self._owner = owner


class BatchedImitationEnv(object):
    def __init__(self, robot="laikago", motions=None, num_robot=1, device=0, seed=0, gid0=0, auto_reset=False,
                 max_episode_steps=600, task=None, phys=None, urdf_search_path=None, action_repeat=None,
                 enable_randomizer=False, randomizer_ranges=None):
        if not motions:
            raise ValueError("No reference motions specified.")           # imitation_task.py:583-584
        self._lib = _abi.load_library()
        self.num_robot = int(num_robot)
        self.device = torch.device("cuda", int(device))
        self.task = dict(TASK_DEFAULTS, **(task or {}))
        self.phys = dict(PHYS_DEFAULTS, **(phys or {}))
        self.model = RobotModel(robot, urdf_search_path, self.phys["mu_plane"], self.phys["margin_factor"])
        self.clips = [m if isinstance(m, MotionClip) else MotionClip.load(m) for m in motions]
        spec = self.model.spec
        cfg = _abi.OrlConfig()
        cfg.abi_version = _abi.ORL_ABI_VERSION
        cfg.num_robot, cfg.device, cfg.gid0, cfg.seed = self.num_robot, int(device), int(gid0), int(seed)
        # SimulationParameters.num_action_repeat (locomotion_gym_config.py); the robot classes fix it to 33
        self.action_repeat = int(action_repeat or spec.NUM_ACTION_REPEAT)
        cfg.action_repeat = self.action_repeat
        cfg.max_episode_steps = int(max_episode_steps)
        cfg.auto_reset = int(bool(auto_reset))
        self.auto_reset = bool(auto_reset)
        self.config_version = 0
        cfg.enable_cycle_sync = int(bool(self.task["enable_cycle_sync"]))
        cfg.sim_time_step = spec.T_STEP
        cfg.env_time_step = self.action_repeat * spec.T_STEP       # quadruped_gym_env.py:177
        cfg.warmup_time = self.task["warmup_time"]
        cfg.ref_state_init_prob = self.task["ref_state_init_prob"]
        if len(self.task["tar_frame_steps"]) != _abi.NUM_TAR_FRAMES:
            raise ValueError("tar_frame_steps must have %d entries" % _abi.NUM_TAR_FRAMES)
        cfg.tar_frame_steps[:] = [int(v) for v in self.task["tar_frame_steps"]]
        for k in ("pose_weight", "velocity_weight", "end_effector_weight", "root_pose_weight", "root_velocity_weight",
                  "pose_err_scale", "velocity_err_scale", "end_effector_err_scale", "end_effector_height_err_scale",
                  "root_pose_err_scale", "root_velocity_err_scale", "dist_fail_threshold", "rot_fail_threshold"):
            setattr(cfg, k, float(self.task[k]))
        for k in ("gravity_z", "contact_kn", "contact_bn", "contact_d0", "contact_bt", "limit_k", "limit_d", "limit_e0",
                  "max_coordinate_velocity", "joint_friction_k", "mu_plane"):
            setattr(cfg, k, float(self.phys[k]))
        cfg.reward_weight = float(self.task["weight"])
        cfg.enable_rand_init_time = int(bool(self.task["enable_rand_init_time"]))
        cfg.clip_time_min, cfg.clip_time_max = float(self.task["clip_time_min"]), float(self.task["clip_time_max"])
        cfg.perturb_init_state_prob = float(self.task["perturb_init_state_prob"])
        noise = self.task["tar_obs_noise"]
        cfg.enable_tar_obs_noise = int(noise is not None)
        cfg.tar_obs_noise = float(noise[0]) if noise is not None else 0.0
        # domain randomizer (run.py:205-211 enables it in train mode only)
        self.enable_randomizer = bool(enable_randomizer)
        self.randomizer_ranges = dict(RANDOMIZER_DEFAULTS, **(randomizer_ranges or {}))
        if sorted(self.randomizer_ranges) != sorted(RANDOMIZER_DEFAULTS):
            raise ValueError("unknown randomizer parameter: %s" % sorted(set(self.randomizer_ranges) - set(RANDOMIZER_DEFAULTS)))
        cfg.enable_randomizer = int(self.enable_randomizer)
        for i, name in enumerate(sorted(RANDOMIZER_DEFAULTS)):
            cfg.rand_lo[i], cfg.rand_hi[i] = [float(v) for v in self.randomizer_ranges[name]]
        self.cfg = cfg
        self.env_time_step = cfg.env_time_step
        self._h = C.c_void_p()
        rc = self._lib.orl_create(C.byref(cfg), C.byref(self.model.c), C.byref(self._h))
        if rc != 0:
            msg = self._lib.orl_last_error(self._h).decode() if self._h else "orl_create failed"
            if self._h:
                self._lib.orl_destroy(self._h)
                self._h = C.c_void_p()
            raise (ValueError if rc == _abi.ORL_ERR_INVALID else OrlError)(msg)
        frames = np.concatenate([c.frames for c in self.clips], axis=0)
        if frames.shape[1] != _abi.POSE_SIZE:
            raise AssertionError("motion frame size %d != pose size %d" % (frames.shape[1], _abi.POSE_SIZE))  # imitation_task.py:689
        nf = np.array([c.frames.shape[0] for c in self.clips], dtype=np.int32)
        fd = np.array([c.frame_duration for c in self.clips], dtype=np.float64)
        fl = np.array([c.flags for c in self.clips], dtype=np.int32)
        frames = np.ascontiguousarray(frames, dtype=np.float64)
        self._check(self._lib.orl_load_motions(self._h, frames.ctypes.data, nf.ctypes.data, fd.ctypes.data, fl.ctypes.data,
                                               len(self.clips)))
        self._total_frames = int(nf.sum())
        lay = _abi.OrlStateLayout()
        self._check(self._lib.orl_state_layout_get(self._h, C.byref(lay)))
        self.layout = {n: getattr(lay, n) for n, _ in lay._fields_}
        N = self.num_robot
        with torch.cuda.device(self.device):
            self.obs = torch.empty((N, _abi.OBS_SIZE), dtype=torch.float32, device=self.device)
            self.reward = torch.empty((N,), dtype=torch.float32, device=self.device)
            self.done = torch.empty((N,), dtype=torch.uint8, device=self.device)
        p = C.c_void_p()
        self._check(self._lib.orl_state_ptr(self._h, C.byref(p)))
        self.state = torch.as_tensor(_DevView(p.value, (N, self.layout["stride"]), "<f4", self), device=self.device)

    # ---- plumbing ------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            msg = self._lib.orl_last_error(self._h).decode()
            if rc == _abi.ORL_ERR_INVALID:
                raise ValueError(msg)
            if rc == _abi.ORL_ERR_ASSERT:
                raise AssertionError(msg)
            raise OrlError("orl error %d: %s" % (rc, msg))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.orl_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- hot path (device tensors) ---------------------------------------------------------------
    def reset(self, mask=None):
        """mask: optional uint8/bool CUDA tensor [N]; returns the observation tensor [N,160] (reused buffer)."""
        mp = None
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            mp = C.c_void_p(mask.data_ptr())
        self._check(self._lib.orl_reset(self._h, mp, C.c_void_p(self.obs.data_ptr()), self._stream()))
        return self.obs

    def step(self, actions):
        """actions: float32 CUDA tensor [N,12] (policy residuals).  Returns (obs, reward, done) device tensors."""
        if actions.dtype != torch.float32 or not actions.is_cuda or not actions.is_contiguous() \
                or tuple(actions.shape) != (self.num_robot, _abi.NUM_MOTORS):
            raise ValueError("actions must be a contiguous float32 CUDA tensor of shape (%d, 12)" % self.num_robot)
        self._check(self._lib.orl_step(self._h, C.c_void_p(actions.data_ptr()), C.c_void_p(self.obs.data_ptr()),
                                       C.c_void_p(self.reward.data_ptr()), C.c_void_p(self.done.data_ptr()), self._stream()))
        return self.obs, self.reward, self.done

    # ---- hot path through host buffers (what the list-based reference API uses) -------------------
    def reset_host(self, mask=None):
        obs = np.empty((self.num_robot, _abi.OBS_SIZE), dtype=np.float32)
        mp = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
            mp = mask.ctypes.data
        self._check(self._lib.orl_reset_host(self._h, mp, obs.ctypes.data))
        return obs

    def alloc_host_buffers(self):
        """Page-locked host buffers (actions, (obs, reward, done)) as NumPy views: `step_host` DMAs straight
        from/to them; ordinary NumPy arrays work too but go through the library's pinned staging copy."""
        t = [torch.empty((self.num_robot, _abi.NUM_MOTORS), dtype=torch.float32).pin_memory(),
             torch.empty((self.num_robot, _abi.OBS_SIZE), dtype=torch.float32).pin_memory(),
             torch.empty(self.num_robot, dtype=torch.float32).pin_memory(),
             torch.empty(self.num_robot, dtype=torch.uint8).pin_memory()]
        self._pinned = getattr(self, "_pinned", []) + t          # keep the allocations alive with the env
        a, o, r, d = [x.numpy() for x in t]
        return a, (o, r, d)

    def step_host(self, actions, out=None):
        # The same validated buffers as in the previous call (the usual loop): their addresses are known.  Building four
        # `.ctypes.data` objects and re-validating took 5 of the ~125 us of a 4096-robot call.
        c = getattr(self, "_host_call", None)
        if c is not None and c[0] is actions and c[1] is out:
            rc = self._lib.orl_step_host(self._h, *c[2])
            if rc != 0:
                self._check(rc)
            return out
        acts = np.ascontiguousarray(actions, dtype=np.float32)
        if acts.shape != (self.num_robot, _abi.NUM_MOTORS):
            raise ValueError("actions must have shape (%d, 12)" % self.num_robot)
        given = out
        if out is None:
            out = (np.empty((self.num_robot, _abi.OBS_SIZE), dtype=np.float32), np.empty(self.num_robot, dtype=np.float32),
                   np.empty(self.num_robot, dtype=np.uint8))
        obs, rew, done = out
        shapes = ((self.num_robot, _abi.OBS_SIZE), (self.num_robot,), (self.num_robot,))
        for a, dt, sh in zip(out, (np.float32, np.float32, np.uint8), shapes):
            if not isinstance(a, np.ndarray) or a.dtype != dt or a.shape != sh or not a.flags.c_contiguous or not a.flags.writeable:
                raise ValueError("output buffers must be writable C-contiguous arrays: obs float32 (N, 160), reward float32 (N,), done uint8 (N,)")
        ptrs = (acts.ctypes.data, obs.ctypes.data, rew.ctypes.data, done.ctypes.data)
        self._check(self._lib.orl_step_host(self._h, *ptrs))
        # remember the call only when the caller's own arrays were used as they are (no converted copy of the actions)
        self._host_call = (actions, given, ptrs) if (acts is actions and given is not None and isinstance(given, tuple)) else None
        return obs, rew, done

    def set_max_episode_steps(self, steps):
        self._check(self._lib.orl_set_max_episode_steps(self._h, int(steps)))
        self.config_version += 1          # kernel arguments travel by value: captured CUDA graphs are stale now

    def set_ref_state_init_prob(self, prob):
        self._check(self._lib.orl_set_ref_state_init_prob(self._h, float(prob)))
        self.config_version += 1

    def stats(self, reset=True):
        out = np.zeros(4, dtype=np.float64)
        self._check(self._lib.orl_stats(self._h, out.ctypes.data, int(reset)))
        return out

    @property
    def launch_count(self):
        return int(self._lib.orl_launch_count(self._h))

    # ---- state import / export (checkpointing, parity tests) ------------------------------------------
    def get_state(self):
        N, stride = self.num_robot, self.layout["stride"]
        raw = np.empty((N, stride), dtype=np.float32)
        self._check(self._lib.orl_get_state_host(self._h, raw.ctypes.data))
        out = {}
        for name, shape in _FIELD_SHAPES.items():
            off = self.layout[name]
            if name in _F64_FIELDS:
                out[name] = raw[:, off:off + 2].copy().view(np.float64)[:, 0]
            elif name in _INT_FIELDS:
                out[name] = raw[:, off].copy().view(np.int32)
            else:
                n = int(np.prod(shape)) if shape else 1
                out[name] = raw[:, off:off + n].reshape((N,) + shape).copy()
        return out

    def set_state(self, st):
        N, stride = self.num_robot, self.layout["stride"]
        raw = np.empty((N, stride), dtype=np.float32)
        self._check(self._lib.orl_get_state_host(self._h, raw.ctypes.data))
        for name, val in st.items():
            off = self.layout[name]
            if name in _F64_FIELDS:
                raw[:, off:off + 2] = np.ascontiguousarray(val, dtype=np.float64).reshape(N, 1).view(np.float32)
            elif name in _INT_FIELDS:
                raw[:, off] = np.ascontiguousarray(val, dtype=np.int32).view(np.float32)
            else:
                shape = _FIELD_SHAPES[name]
                n = int(np.prod(shape)) if shape else 1
                raw[:, off:off + n] = np.asarray(val, dtype=np.float32).reshape(N, n)
        self._check(self._lib.orl_set_state_host(self._h, raw.ctypes.data))

    # ---- parity entry points ------------------------------------------------------------------------
    def motion_tables(self):
        F, K = self._total_frames, len(self.clips)
        frames, vels = np.empty((F, 19)), np.empty((F, 18))
        dur, dpos, dh = np.empty(K), np.empty((K, 3)), np.empty(K)
        self._check(self._lib.orl_get_motion_tables(self._h, frames.ctypes.data, vels.ctypes.data, dur.ctypes.data,
                                                    dpos.ctypes.data, dh.ctypes.data))
        return dict(frames=frames, vels=vels, duration=dur, delta_pos=dpos, delta_heading=dh)

    def motion_lookup(self, clip, time):
        clip = torch.as_tensor(np.asarray(clip, dtype=np.int32), device=self.device)
        time = torch.as_tensor(np.asarray(time, dtype=np.float64), device=self.device)
        n = clip.numel()
        i32 = lambda: torch.empty(n, dtype=torch.int32, device=self.device)
        f0, f1, cyc = i32(), i32(), i32()
        blend = torch.empty(n, dtype=torch.float64, device=self.device)
        pose = torch.empty((n, 19), dtype=torch.float32, device=self.device)
        vel = torch.empty((n, 18), dtype=torch.float32, device=self.device)
        P = lambda t: C.c_void_p(t.data_ptr())
        self._check(self._lib.orl_motion_lookup(self._h, P(clip), P(time), n, P(f0), P(f1), P(cyc), P(blend), P(pose), P(vel),
                                                self._stream()))
        return dict(f0=f0.cpu().numpy(), f1=f1.cpu().numpy(), cycle=cyc.cpu().numpy(), blend=blend.cpu().numpy(),
                    pose=pose.cpu().numpy(), vel=vel.cpu().numpy())

    def _dev(self, a, cols):
        t = torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device=self.device)
        assert t.shape[1] == cols
        return t

    def fk(self, pose):
        pose = self._dev(pose, 19)
        n = pose.shape[0]
        out = torch.zeros((n, self.model.num_joints, 3), dtype=torch.float32, device=self.device)
        self._check(self._lib.orl_fk(self._h, C.c_void_p(pose.data_ptr()), n, C.c_void_p(out.data_ptr()), self._stream()))
        return out.cpu().numpy()

    def latency_rings(self):
        """Observation history behind the randomised control latency: (motor angles [N,42,12], base [N,42,8])."""
        R = 42
        rq = np.empty((self.num_robot, R, 12), dtype=np.float32)
        rb = np.empty((self.num_robot, R, 8), dtype=np.float32)
        self._check(self._lib.orl_get_latency_rings_host(self._h, rq.ctypes.data, rb.ctypes.data))
        return rq, rb

    def set_latency_rings(self, ring_q, ring_b):
        rq = np.ascontiguousarray(ring_q, dtype=np.float32); rb = np.ascontiguousarray(ring_b, dtype=np.float32)
        if rq.shape != (self.num_robot, 42, 12) or rb.shape != (self.num_robot, 42, 8):
            raise ValueError("ring shapes must be (N, 42, 12) and (N, 42, 8)")
        self._check(self._lib.orl_set_latency_rings_host(self._h, rq.ctypes.data, rb.ctypes.data))

    def checkpoint(self):
        """Everything needed to resume bit-exactly (SURVEY.md section 5 "checkpoint / resume"): the state records and, in
        randomizer mode, the observation history of the randomised control latency."""
        ck = dict(state=self.get_state())
        if self.enable_randomizer:
            ck["latency_rings"] = self.latency_rings()
        return ck

    def restore(self, ck):
        self.set_state(ck["state"])
        if self.enable_randomizer:
            self.set_latency_rings(*ck["latency_rings"])

    def set_dynamics_rand(self, rand=None):
        """Randomizer values [n][12] (`rand_params` layout) for the next mass_matrix / inverse_dynamics /
        forward_dynamics calls; None restores the model's own parameters."""
        if rand is not None:
            rand = np.asarray(rand, dtype=np.float32)
            if rand.shape[1] < 12:              # the state field has 10 values; the entry point strides by 12
                rand = np.concatenate([rand, np.zeros((rand.shape[0], 12 - rand.shape[1]), np.float32)], axis=1)
        self._parity_rand = None if rand is None else self._dev(rand, 12)
        ptr = None if rand is None else C.c_void_p(self._parity_rand.data_ptr())
        self._check(self._lib.orl_set_dynamics_rand(self._h, ptr))

    def mass_matrix(self, pose):
        pose = self._dev(pose, 19)
        n = pose.shape[0]
        out = torch.empty((n, 18, 18), dtype=torch.float32, device=self.device)
        self._check(self._lib.orl_mass_matrix(self._h, C.c_void_p(pose.data_ptr()), n, C.c_void_p(out.data_ptr()), self._stream()))
        return out.cpu().numpy()

    def inverse_dynamics(self, pose, vel, acc):
        pose, vel, acc = self._dev(pose, 19), self._dev(vel, 18), self._dev(acc, 18)
        n = pose.shape[0]
        out = torch.empty((n, 18), dtype=torch.float32, device=self.device)
        self._check(self._lib.orl_inverse_dynamics(self._h, C.c_void_p(pose.data_ptr()), C.c_void_p(vel.data_ptr()),
                                                   C.c_void_p(acc.data_ptr()), n, C.c_void_p(out.data_ptr()), self._stream()))
        return out.cpu().numpy()

    def forward_dynamics(self, pose, vel, tau):
        pose, vel, tau = self._dev(pose, 19), self._dev(vel, 18), self._dev(tau, 12)
        n = pose.shape[0]
        out = torch.empty((n, 18), dtype=torch.float32, device=self.device)
        self._check(self._lib.orl_forward_dynamics(self._h, C.c_void_p(pose.data_ptr()), C.c_void_p(vel.data_ptr()),
                                                   C.c_void_p(tau.data_ptr()), n, C.c_void_p(out.data_ptr()), self._stream()))
        return out.cpu().numpy()

    def reward_terms(self):
        out = torch.empty((self.num_robot, 5), dtype=torch.float32, device=self.device)
        self._check(self._lib.orl_reward_terms(self._h, C.c_void_p(out.data_ptr()), self._stream()))
        return out.cpu().numpy()

    # ---- spaces (wrapper_env.py:127-145, minitaur.py:145-155, imitation_task.py:303-335) ---------------
    def observation_bounds(self):
        imu = np.array([2 * np.pi, 2 * np.pi, 2000 * np.pi, 2000 * np.pi] * 3)
        hi = np.concatenate([imu, np.ones(36), np.pi * np.ones(36)])
        fr = self.motion_tables()["frames"]
        low, high = fr.min(axis=0), fr.max(axis=0)
        low[0:3], high[0:3], low[3:7], high[3:7] = -2, 2, -1, 1
        n = _abi.NUM_TAR_FRAMES
        return np.concatenate([-hi, np.tile(low, n)]), np.concatenate([hi, np.tile(high, n)])


# minitaur_motor.py:24-37: per-motor layout of a HYBRID command and the control-mode constants
MOTOR_COMMAND_DIMENSION = 5
POSITION_INDEX, POSITION_GAIN_INDEX, VELOCITY_INDEX, VELOCITY_GAIN_INDEX, TORQUE_INDEX = 0, 1, 2, 3, 4
MOTOR_POSITION, MOTOR_TORQUE, MOTOR_HYBRID = _abi.MOTOR_POSITION, _abi.MOTOR_TORQUE, _abi.MOTOR_HYBRID


def motor_convert_to_torque(mode, commands, motor_angle=None, motor_velocity=None, kp=None, kd=None, strength=None,
                            torque_limits=None):
    """Batched MotorModel.convert_to_torque (minitaur_motor.py:104-173) on the device (orl_motor_convert_to_torque).

    commands: CUDA float32 [n, 12] (POSITION: desired angles, TORQUE: torques) or [n, 60] (HYBRID); the other arguments
    CUDA float32 tensors ([n, 12], kp / kd / torque_limits [12]) or None.  Returns the motor torques [n, 12]; the env step
    itself is POSITION only, like the reference's (minitaur.py:113)."""
    lib = _abi.load_library()
    n = int(commands.shape[0])
    width = _abi.NUM_MOTORS * (MOTOR_COMMAND_DIMENSION if mode == MOTOR_HYBRID else 1)
    if tuple(commands.shape) != (n, width):
        raise AssertionError("motor commands must have shape (n, %d)" % width)          # :139, 148, 155
    ts = [commands, motor_angle, motor_velocity, kp, kd, strength, torque_limits]
    for t in ts:
        if t is not None and (not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous()):
            raise ValueError("contiguous float32 CUDA tensors expected")
    out = torch.empty((n, _abi.NUM_MOTORS), dtype=torch.float32, device=commands.device)
    ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    with torch.cuda.device(commands.device):
        rc = lib.orl_motor_convert_to_torque(int(mode), *[ptr(t) for t in ts], n, C.c_void_p(out.data_ptr()),
                                             C.c_void_p(torch.cuda.current_stream(commands.device).cuda_stream))
    if rc != 0:
        raise ValueError("orl_motor_convert_to_torque: unsupported motor control mode or missing argument")    # :133-135
    return out

