"""Drop-in mirror of the reference's env classes on top of the CUDA environment.

Same class names, constructor arguments, return types and quirks as the reference, so that
`run.py:build_env` works unchanged when it imports these instead of the PyBullet-backed modules
(INTEGRATION.md shows the three import lines to change):

    reference class                                                    mirrored here
    envs/quadruped_robot/robots/minitaur.py          Minitaur          Minitaur
    envs/quadruped_robot/task/imitation_task.py      ImitationTask     ImitationTask
    envs/quadruped_robot/quadruped_gym_env.py        LocomotionGymEnv  LocomotionGymEnv
    envs/quadruped_robot/wrapper_env.py              WrapperEnv        WrapperEnv

All arithmetic of reset()/step() runs in the fused CUDA kernels through the C ABI (host-buffer entry
points orl_reset_host / orl_step_host); these classes only adapt lists <-> arrays.  Per-object methods
of the reference that assume one Python object per robot stepping one Bullet world (Minitaur.robot_step,
ImitationTask.update, ...) have no meaning in a fused batched step and raise NotImplementedError.
Every constructor option of the reference's ImitationTask is accepted (clip switching, initial-state perturbation,
target-observation noise, weight, enable_rand_init_time: the option instantiation of the kernels).
"""
import math

import numpy as np

from . import robots as robots_pkg
from .env import BatchedImitationEnv, TASK_DEFAULTS
from .motion import MotionClip

try:  # the reference uses gym.spaces.Box; fall back to a minimal look-alike when gym is absent
    from gym.spaces import Box as _GymBox   # pragma: no cover
except Exception:  # noqa: BLE001
    _GymBox = None


class Box(object):
    """Minimal gym.spaces.Box look-alike (low/high/shape/dtype/sample/contains)."""

    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low = np.array(low, dtype=dtype)
        self.high = np.array(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)
        self._rng = np.random.RandomState()

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def seed(self, seed=None):
        self._rng = np.random.RandomState(seed)
        return [seed]


def _box(low, high, dtype=np.float32):
    if _GymBox is not None:
        return _GymBox(np.array(low, dtype=dtype), np.array(high, dtype=dtype), dtype=dtype)
    return Box(low, high, dtype=dtype)


class Minitaur(object):
    """Robot description (reference minitaur.py:39-199): constants, spaces, id sets."""

    def __init__(self, name_robot, robot_index=0, enable_randomizer=True):
        robot = robots_pkg.get(name_robot)                      # raises ValueError("wrong robot select")
        self.name_robot = name_robot
        self._robot_index = robot_index
        self._enable_randomizer = enable_randomizer
        self._action_repeat = robot.NUM_ACTION_REPEAT
        self.time_step = robot.T_STEP
        self._urdf_file = robot.URDF_FILENAME
        self.num_motors = robot.NUM_MOTORS
        self.name_motor = robot.MOTOR_NAMES
        self.pattern = robot.PATTERN
        self._init_pos = robot.INIT_POSITION
        self._init_quat = robot.INIT_QUAT
        self._init_motor_angle = robot.INIT_MOTOR_ANGLES
        self._motor_direction = robot.JOINT_DIRECTIONS
        self._motor_offset = robot.JOINT_OFFSETS
        self._control_latency = robot.CTRL_LATENCY
        self.num_legs = self.num_motors // robot.DOFS_PER_LEG
        self.action_space = _box([-2 * math.pi] * 12, [2 * math.pi] * 12)        # minitaur.py:145-148
        self._env = None

    # -- bound by LocomotionGymEnv
    def _bind(self, env):
        self._env = env
        m = env._cuda.model
        self._chassis_link_ids, self._motor_link_ids = m.chassis_link_ids, m.motor_link_ids
        self._foot_link_ids, self._leg_link_ids = m.foot_link_ids, m.leg_link_ids
        self._motor_id_list = m.motor_id_list

    def _state(self):
        return self._env._state()

    @property
    def action_repeat(self):
        return self._action_repeat

    @property
    def step_counter(self):
        return int(self._state()["step_counter"][self._robot_index])

    @property
    def last_action(self):
        return self._state()["act_hist"][self._robot_index, 0].astype(np.float64)

    @property
    def chassis_link_ids(self):
        return self._chassis_link_ids

    def get_foot_link_ids(self):
        return self._foot_link_ids

    def get_urdf_file(self):
        return self._urdf_file

    def get_default_init_pos(self):
        return self._init_pos

    def get_default_init_ori(self):
        return self._init_quat

    def get_default_init_joint_pos(self):
        return (self._init_motor_angle + self._motor_offset) * self._motor_direction       # minitaur.py:1190-1194

    def get_time_since_reset(self):
        return int(self._state()["substep_counter"][self._robot_index]) * self.time_step    # minitaur.py:397-398

    def get_base_pos(self):
        """World position incl. the reference's grid slot (minitaur.py:246-249)."""
        p = self._state()["pos"][self._robot_index].astype(np.float64)
        p[0] -= int(self._robot_index / 4) * 2
        p[1] += (self._robot_index % 4) * 2
        return p

    def get_true_motor_angles(self):
        q = self._state()["q"][self._robot_index].astype(np.float64)
        m = self._env._cuda.model
        urdf = np.array([q[m.q_index[j]] for j in m.motor_id_list])
        return (urdf - self._motor_offset) * self._motor_direction                            # minitaur.py:549-553

    def get_action_dim(self):
        return self.num_motors

    def terminate(self):
        pass

    def robot_step(self, i):
        raise NotImplementedError("the 33 substeps are fused into orl_step; there is no per-robot substep call")


class ImitationTask(object):
    """Task configuration + the imitation-task query API (reference imitation_task.py:31-339)."""

    def __init__(self, weight=1.0, ref_motion_filenames=None, enable_cycle_sync=True, clip_velocity=None,
                 tar_frame_steps=None, clip_time_min=np.inf, clip_time_max=np.inf, ref_state_init_prob=1.0,
                 enable_rand_init_time=True, warmup_time=0.0, pose_weight=0.5, velocity_weight=0.05,
                 end_effector_weight=0.2, root_pose_weight=0.15, root_velocity_weight=0.1, pose_err_scale=5.0,
                 velocity_err_scale=0.1, end_effector_err_scale=40, end_effector_height_err_scale=3.0,
                 root_pose_err_scale=20, root_velocity_err_scale=2, perturb_init_state_prob=0.0, tar_obs_noise=None,
                 draw_ref_model_alpha=0.5):
        assert ref_motion_filenames is not None                                    # imitation_task.py:109
        # clip_velocity is stored and never read by the reference either (imitation_task.py:106)
        self._clip_velocity = clip_velocity
        self._ref_motion_filenames = ref_motion_filenames
        self._tar_frame_steps = [1, 2] if tar_frame_steps is None else list(tar_frame_steps)   # :118-119
        self.params = dict(
            enable_cycle_sync=enable_cycle_sync, tar_frame_steps=tuple(self._tar_frame_steps),
            ref_state_init_prob=ref_state_init_prob, warmup_time=warmup_time, pose_weight=pose_weight,
            velocity_weight=velocity_weight, end_effector_weight=end_effector_weight, root_pose_weight=root_pose_weight,
            root_velocity_weight=root_velocity_weight, pose_err_scale=pose_err_scale, velocity_err_scale=velocity_err_scale,
            end_effector_err_scale=end_effector_err_scale, end_effector_height_err_scale=end_effector_height_err_scale,
            root_pose_err_scale=root_pose_err_scale, root_velocity_err_scale=root_velocity_err_scale,
            # options run.py leaves off: the option instantiation of the CUDA kernels (orl.h "task options")
            weight=float(weight), enable_rand_init_time=bool(enable_rand_init_time), clip_time_min=float(clip_time_min),
            clip_time_max=float(clip_time_max), perturb_init_state_prob=float(perturb_init_state_prob),
            tar_obs_noise=None if tar_obs_noise is None else tuple(float(v) for v in tar_obs_noise))
        self._env, self._index = None, 0

    def _bind(self, env, index):
        self._env, self._index = env, index

    # ---- query API ----
    def get_num_motions(self):
        return len(self._env._cuda.clips)

    def get_num_tar_frames(self):
        return len(self._tar_frame_steps)

    def get_active_motion(self):
        return self._env._cuda.clips[int(self._env._state()["clip"][self._index])]

    def is_motion_over(self):
        st = self._env._state()
        clip = self.get_active_motion()
        t = self._motion_time(st)
        return (clip.loop_mode != "Wrap") and t >= clip.frame_duration * (clip.frames.shape[0] - 1)

    def _motion_time(self, st):
        t = int(st["substep_counter"][self._index]) * self._env._cuda.model.spec.T_STEP + float(st["time_offset"][self._index])
        return t - self.params["warmup_time"] if st["warmup"][self._index] else t

    def reward(self):
        """imitation_task.py:341-356 for the CURRENT state (device kernel orl_reward_terms)."""
        p = self.params
        t = self._env._cuda.reward_terms()[self._index].astype(np.float64)
        return float(p["pose_weight"] * t[0] + p["velocity_weight"] * t[1] + p["end_effector_weight"] * t[2]
                     + p["root_pose_weight"] * t[3] + p["root_velocity_weight"] * t[4]) * p["weight"]      # :356

    cal_reward = reward

    def done(self):
        return bool(self._env._last_done[self._index])

    def build_target_obs(self):
        return self._env._last_obs[self._index, 84:].astype(np.float64)

    def get_target_obs_bounds(self):
        low, high = self._env._cuda.observation_bounds()
        return low[84:], high[84:]

    def set_ref_state_init_prob(self, prob):
        self.params["ref_state_init_prob"] = prob
        self._env._cuda.set_ref_state_init_prob(prob)

    def reset(self):
        raise NotImplementedError("per-task reset is fused into LocomotionGymEnv.reset (orl_reset)")

    def update(self):
        raise NotImplementedError("the reference-motion update is fused into LocomotionGymEnv.step (orl_step)")


class LocomotionGymEnv(object):
    """reference quadruped_gym_env.py:29-348 (DIRECT mode only; rendering is out of scope)."""

    def __init__(self, robot=None, task=None, device=0, seed=0, urdf_search_path=None):
        if robot is None or task is None:
            raise ValueError("no input robot or task")                              # quadruped_gym_env.py:50-51
        if len(robot) != len(task) or len(robot) == 0:
            raise ValueError("robot and task lists must have the same, non-zero length")
        self._robot, self._task = robot, task
        self.num_robot = len(robot)
        names = {r.name_robot for r in robot}
        if len(names) != 1:
            raise ValueError("all robots of one env must be of the same kind")
        p0 = task[0].params
        if any(t.params != p0 or t._ref_motion_filenames != task[0]._ref_motion_filenames for t in task):
            raise ValueError("all tasks of one env must share their configuration")
        sim = self._load_sim_params()
        if abs(sim["sim_time_step_s"] - robot[0].time_step) > 1e-12:
            raise ValueError("pybullet_sim_param.yaml sim_time_step_s differs from the robot's T_STEP")
        clips = [m if isinstance(m, MotionClip) else MotionClip.load(m) for m in task[0]._ref_motion_filenames]
        tp = dict(TASK_DEFAULTS, **p0)
        self._cuda = BatchedImitationEnv(robot[0].name_robot, clips, num_robot=self.num_robot, device=device, seed=seed,
                                         auto_reset=False, max_episode_steps=2 ** 30, task=tp,
                                         urdf_search_path=urdf_search_path,
                                         enable_randomizer=bool(robot[0]._enable_randomizer))   # minitaur.py:273-276
        self._num_action_repeat = robot[0].action_repeat
        self._sim_time_step = sim["sim_time_step_s"]
        self._env_time_step = self._num_action_repeat * self._sim_time_step          # :177
        self._env_step_counter = 0
        self._is_render = False
        self._state_cache = None
        for i, (r, t) in enumerate(zip(robot, task)):
            r._robot_index = i
            r._bind(self)
            t._bind(self, i)
        self.action_space = robot[0].action_space
        imu = [2 * np.pi, 2 * np.pi, 2000 * np.pi, 2000 * np.pi] * 3
        hi = np.concatenate([imu, np.ones(36), np.pi * np.ones(36)])
        self.observation_space = _box(-hi, hi)                                     # :289-320 flattened sensor spaces
        self._last_obs = np.zeros((self.num_robot, 160), dtype=np.float32)
        self._last_done = np.zeros(self.num_robot, dtype=np.uint8)
        self._out = (np.empty((self.num_robot, 160), np.float32), np.empty(self.num_robot, np.float32),
                     np.empty(self.num_robot, np.uint8))
        self.reset()                                                               # :211

    @staticmethod
    def _load_sim_params():
        try:                                                                       # cwd-relative like the reference, :159
            import yaml
            with open("OpenRoboRL/config/pybullet_sim_param.yaml") as f:
                d = yaml.safe_load(f)
            if "quadruped_robot" not in d:
                raise ValueError("Hyperparameters not found for pybullet_sim_config.yaml")
            return d["quadruped_robot"]
        except (IOError, OSError, ImportError):
            return dict(enable_rendering=False, sim_time_step_s=0.001, enable_rendering_gui=True, num_sim_iter_step=300)

    def _state(self):
        if self._state_cache is None:
            self._state_cache = self._cuda.get_state()
        return self._state_cache

    def seed(self, seed=None):
        return [seed]

    def reset(self, reset_visualization_camera=False):
        self._last_obs = self._cuda.reset_host()
        self._last_done[:] = 0
        self._env_step_counter = 0
        self._state_cache = None
        return [self._last_obs[i, :84].astype(np.float64) for i in range(self.num_robot)]

    def step(self, action):
        init = self._robot[0]._init_motor_angle
        a = np.empty((self.num_robot, 12), dtype=np.float32)
        for i in range(self.num_robot):
            a[i] = action[i]
            try:
                action[i] += init                  # the reference mutates the caller's array (minitaur.py:281)
            except TypeError:
                pass
        obs, rew, done = self._cuda.step_host(a, self._out)
        self._last_obs, self._last_done = obs, done
        self._env_step_counter += 1
        self._state_cache = None
        return ([obs[i, :84].astype(np.float64) for i in range(self.num_robot)],
                [float(r) for r in rew], [bool(d) for d in done], {})

    def close(self):
        self._cuda.close()

    def get_ground(self):
        return 0

    @property
    def rendering_enabled(self):
        return self._is_render

    @property
    def pybullet_client(self):
        return None

    @property
    def env_step_counter(self):
        return self._env_step_counter

    @property
    def env_time_step(self):
        return self._env_time_step

    @property
    def task(self):
        return self._task

    @property
    def robot(self):
        return self._robot


class WrapperEnv(object):
    """reference wrapper_env.py:26-159: target observations, episode time limit, curriculum."""

    def __init__(self, gym_env, episode_length_start=1000, episode_length_end=1000, curriculum_steps=0, num_parallel_envs=1):
        self._gym_env = gym_env
        self.observation_space = self._build_observation_space()
        self._episode_length_start = episode_length_start
        self._episode_length_end = episode_length_end
        self._curriculum_steps = int(np.ceil(curriculum_steps / num_parallel_envs))
        self._total_step_count = 0
        if self._enable_curriculum():
            self._update_time_limit()
        self.seed()

    def __getattr__(self, attr):
        return getattr(self._gym_env, attr)

    def step(self, action):
        g = self._gym_env
        original_observation, reward, done, _ = g.step(action)
        observation = self._modify_observation(original_observation)
        terminated = done
        num_robot = g.num_robot
        info = [0 for _ in range(num_robot)]
        for i in range(num_robot):
            # the time limit is already folded into `done` by the kernel with the same rule:
            # done[i] |= env_step_counter >= max_episode_steps   (wrapper_env.py:78-79)
            info[i] = {"terminated": terminated}                    # the SAME list object (quirk A.6.6)
        if np.sum(done) > 0:
            self._total_step_count += num_robot                     # quirk A.6.7
        return observation, reward, done, info

    def reset(self, initial_motor_angles=None, reset_duration=0.0):
        original_observation = self._gym_env.reset()
        observation = self._modify_observation(original_observation)
        if self._enable_curriculum():
            self._update_time_limit()
        return observation

    def _modify_observation(self, original_observation):
        g = self._gym_env
        return [np.concatenate([original_observation[i], g._last_obs[i, 84:].astype(np.float64)], axis=-1)
                for i in range(len(original_observation))]

    def _build_observation_space(self):
        low, high = self._gym_env._cuda.observation_bounds()
        return _box(low, high)

    def _enable_curriculum(self):
        return self._curriculum_steps > 0

    def _update_time_limit(self):
        t = float(self._total_step_count) / self._curriculum_steps
        t = np.clip(t, 0.0, 1.0)
        t = np.power(t, 3.0)
        new_steps = int((1.0 - t) * self._episode_length_start + t * self._episode_length_end)
        self._max_episode_steps = new_steps
        self._gym_env._cuda.set_max_episode_steps(max(new_steps, 1))
        return


def build_env(name_task, num_robot, num_parallel_envs, mode, enable_randomizer, motion_files=None, device=0, seed=0):
    """Mirror of reference run.py:49-76 for the two imitation tasks."""
    if name_task not in ("imitation_learning_laikago", "imitation_learning_minicheetah"):
        raise ValueError("task not supported by the CUDA env: %s" % name_task)
    start, end = 20, 600
    name_robot = name_task.replace("imitation_learning_", "")
    robot = [Minitaur(name_robot=name_robot, robot_index=i, enable_randomizer=enable_randomizer) for i in range(num_robot)]
    task = [ImitationTask(ref_motion_filenames=motion_files, enable_cycle_sync=True, tar_frame_steps=[1, 2, 10, 30],
                          ref_state_init_prob=0.9, warmup_time=0.25) for _ in range(num_robot)]
    env = LocomotionGymEnv(robot, task, device=device, seed=seed)
    if mode == "test":
        start = end
    return WrapperEnv(env, episode_length_start=start, episode_length_end=end, curriculum_steps=30000000,
                      num_parallel_envs=num_parallel_envs)
