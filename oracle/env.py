"""Batched float64 restatement of one `WrapperEnv.step()` / `reset()` of the reference.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every method cites the reference lines it follows
(R/ = OpenRoboRL/).  The rigid-body substep (`stepSimulation`, third party PyBullet) is replaced
by this repo's own documented model (DESIGN.md "physics model"), evaluated here with the
independent textbook formulation of oracle/rbd.py.

Coordinates: each robot lives in its own world whose xy origin is its reset slot (the reference's
grid offset minitaur.py:246-249 only translates the world and cancels in every reward/observation).
"""
import numpy as np

from . import xform as xf
from . import philox
from .motion import MotionLib
from .rbd import UrdfTree
from .robots import get_spec, classify_links

DEFAULT_PHYS = dict(
    gravity=(0.0, 0.0, -10.0),   # quadruped_gym_env.py:200
    kn=1.0e5, bn=500.0, d0=0.002, bt=3000.0, mu_plane=1.0,
    k_lim=1000.0, d_lim=2.0, e0=0.01, margin_factor=0.02, max_coordinate_velocity=100.0,
    joint_friction_k=0.5)

# imitation_task.py:45-58 as configured by run.py:58-64
DEFAULT_TASK = dict(
    pose_weight=0.5, velocity_weight=0.05, end_effector_weight=0.2, root_pose_weight=0.15,
    root_velocity_weight=0.1, pose_err_scale=5.0, velocity_err_scale=0.1, end_effector_err_scale=40.0,
    end_effector_height_err_scale=3.0, root_pose_err_scale=20.0, root_velocity_err_scale=2.0,
    tar_frame_steps=(1, 2, 10, 30), ref_state_init_prob=0.9, warmup_time=0.25, enable_cycle_sync=True,
    dist_fail_threshold=1.0, rot_fail_threshold=0.5 * np.pi,
    # options run.py leaves at their defaults (imitation_task.py:36-57)
    weight=1.0, enable_rand_init_time=True, clip_time_min=np.inf, clip_time_max=np.inf,
    perturb_init_state_prob=0.0, tar_obs_noise=None)

# action_filter.py:42-44 + scipy.signal.butter(2, [4/(0.5*fs)]) at fs = 1/(0.001*33); SURVEY.md App. B
BUTTER_B = np.array([0.10669309218914798, 0.21338618437829596, 0.10669309218914798])
BUTTER_A = np.array([1.0, -0.8877194537361395, 0.31449182249273155])

HIST = 44         # ring entries kept: latency <= 0.04 s = 40 substeps needs entries 0..41 (minitaur.py:336-357; the
                  # reference keeps 100)
TRUE_OBS = 43     # minitaur.py:304-311


class OracleEnv(object):
    def __init__(self, robot_name, urdf_path, clips, num_robot, phys=None, task=None, seed=0,
                 gid0=0, max_episode_steps=600, randomizer=False, randomizer_ranges=None):
        self.spec = get_spec(robot_name)
        self.tree = UrdfTree(urdf_path)
        self.lib = MotionLib(clips)
        self.N = int(num_robot)
        self.phys = dict(DEFAULT_PHYS, **(phys or {}))
        self.task = dict(DEFAULT_TASK, **(task or {}))
        self.seed = seed
        self.gid = np.arange(self.N, dtype=np.int64) + gid0
        self.max_episode_steps = max_episode_steps
        from .randomizer import ALL_PARAMS
        self.randomizer = bool(randomizer)
        self.randomizer_ranges = dict(ALL_PARAMS, **(randomizer_ranges or {}))
        t, s = self.tree, self.spec
        assert t.n_dof == 12 and self.lib.frames.shape[1] == 7 + t.n_dof  # imitation_task.py:689-690
        ids = classify_links(s, t.joint_names)
        self.ids = ids
        self.foot_links = np.array(ids["foot"])
        # motor m -> revolute index (q order = URDF link order), minitaur.py:860-864
        self.motor_q = np.array([t.qidx[t.joint_names.index(n)] for n in s.motor_names])
        assert np.all(self.motor_q >= 0)
        self.revolute_links = np.nonzero(t.movable)[0]
        pts = t.collision_points()
        self.cp_body = np.array([p[0] for p in pts])
        self.cp_local = np.stack([p[1] for p in pts])
        self.cp_r = np.array([p[2] for p in pts])
        self.cp_mu = t.mu[self.cp_body] * self.phys["mu_plane"]
        self.cp_foot = np.array([(b - 1) in ids["foot"] for b in self.cp_body])
        # Bullet contact-breaking threshold ~ 0.02 * angular-motion disc of the shape (A.10, unpinned)
        marg = []
        for b, cols in enumerate(t.cols):
            for kind, prm, cxyz, crpy in cols:
                if kind == "sphere":
                    rad, n = prm[0], 1
                elif kind == "box":
                    rad, n = 0.5 * np.linalg.norm(prm), 8
                elif kind in ("cylinder", "capsule"):
                    rad, n = np.hypot(prm[0], prm[1] / 2) if kind == "cylinder" else prm[0] + prm[1] / 2, 2
                else:
                    continue
                disc = rad + np.linalg.norm(cxyz - t.com[b])
                marg += [self.phys["margin_factor"] * disc] * n
        self.cp_margin = np.array(marg)
        N = self.N
        self.pos = np.zeros((N, 3)); self.quat = np.zeros((N, 4)); self.q = np.zeros((N, 12))
        self.lin = np.zeros((N, 3)); self.ang = np.zeros((N, 3)); self.qd = np.zeros((N, 12))
        self.substep_counter = np.zeros(N, dtype=np.int64)
        self.step_counter = np.zeros(N, dtype=np.int64)
        self.env_step_counter = np.zeros(N, dtype=np.int64)
        self.episode = np.zeros(N, dtype=np.int64)
        self.filt_x = np.zeros((N, 2, 12)); self.filt_y = np.zeros((N, 2, 12))
        self.prev_filtered = np.zeros((N, 12)); self.has_prev = np.zeros(N, dtype=bool)
        self.action = np.zeros((N, 12)); self.last_action = np.zeros((N, 12))
        self.hist = np.zeros((N, HIST, TRUE_OBS)); self.hist_len = np.zeros(N, dtype=np.int64)
        self.ctrl_obs = np.zeros((N, TRUE_OBS))
        self.obs_torque = np.zeros((N, 12))
        self.imu_hist = np.zeros((N, 3, 4)); self.act_hist = np.zeros((N, 3, 12)); self.ang_hist = np.zeros((N, 3, 12))
        self.clip = np.zeros(N, dtype=np.int64); self.time_offset = np.zeros(N); self.warmup = np.zeros(N, dtype=bool)
        self.origin_pos = np.zeros((N, 3)); self.origin_rot = np.tile(np.array([0.0, 0, 0, 1]), (N, 1))
        self.prev_phase = np.zeros(N)
        self.clip_change_time = np.full(N, np.inf); self.clip_changes = np.zeros(N, dtype=np.int64)
        # replay of draws recorded from the reference's classes (tests): per robot a list of (kind, value) in the
        # reference's call order; None: the keyed Philox draws of the product's contract (oracle/philox.py)
        self.replay = None
        self.ref_pose = np.zeros((N, 19)); self.ref_vel = np.zeros((N, 18))
        self.motor_strength = np.ones((N, 12))
        # domain randomizer state (oracle/randomizer.py); identity when it is off
        nb = t.n_links + 1
        self.link_mass = np.tile(np.asarray(t.mass, dtype=np.float64), (N, 1))
        self.inertia_scale = np.ones((N, nb))
        self.foot_friction = np.full(N, -1.0)            # < 0: the URDF value
        self.joint_friction = np.zeros((N, 12))          # Coulomb torque limit per URDF joint coordinate
        self.ctrl_latency = np.full(N, float(s.ctrl_latency))   # set_ctrl_latency (minitaur.py:377-385)
        self.rand_values = None
        self.default_pose = np.concatenate([s.init_pos, s.init_quat,
                                            (s.init_motor + s.joint_offset) * s.joint_dir])  # minitaur.py:1190-1194
        self._urdf_from_motor = np.argsort(self.motor_q)

    # ---- random draws of the task options -------------------------------------------------
    def _draw(self, kind, values, mask):
        """`values`: the keyed Philox draw of every robot, already mapped to the reference call's result
        (np.random.uniform(lo, hi) = lo + (hi - lo) u, randn * std + mean, randint).  In replay mode the robots in
        `mask` take the next recorded reference draw instead (kind 'u' / 'i' / 'n' must match the record)."""
        if self.replay is None:
            return values
        out = np.array(values, dtype=np.int64 if kind == "i" else np.float64)
        for i in np.nonzero(mask)[0]:
            k, v = self.replay[i].pop(0)
            assert k == kind, (i, k, kind)
            out[i] = v
        return out

    # ---- space conversions (minitaur.py:549-553, 575-578, 755-756) ------------------------
    def motor_angles_true(self):
        s = self.spec
        return (self.q[:, self.motor_q] - s.joint_offset) * s.joint_dir

    def motor_vels_true(self):
        return self.qd[:, self.motor_q] * self.spec.joint_dir

    # ---- observation ring (minitaur.py:304-357) -------------------------------------------
    def _true_obs(self):
        s = self.spec
        rel = xf.quaternion_multiply(self.quat, xf.quaternion_inverse(s.init_quat))   # :325-331
        w_local = xf.QuaternionRotatePoint(self.ang, xf.quaternion_inverse(rel))      # :652-672
        return np.concatenate([self.motor_angles_true(), self.motor_vels_true(), self.obs_torque, rel, w_local], axis=-1)

    def _delay_obs(self, latency):
        """minitaur.py:336-357 with a per-robot latency (the randomizer sets it, :377-385)."""
        dt = self.spec.time_step
        latency = np.broadcast_to(np.asarray(latency, dtype=np.float64), (self.N,))
        newest = self.hist[:, 0]
        n = (latency / dt).astype(np.int64)                      # int(latency / self.time_step)
        assert np.all(n + 1 < HIST)
        rem = latency - n * dt
        alpha = (rem / dt)[:, None]
        r = np.arange(self.N)
        blend = (1.0 - alpha) * self.hist[r, n] + alpha * self.hist[r, n + 1]
        oldest = self.hist[r, np.clip(self.hist_len - 1, 0, HIST - 1)]
        out = np.where((n + 1 >= self.hist_len)[:, None], oldest, blend)
        return np.where(((self.hist_len == 1) | (latency <= 0))[:, None], newest, out)

    def receive_obs(self, mask=None):
        """minitaur.py:313-334."""
        m = np.ones(self.N, dtype=bool) if mask is None else mask
        new = self._true_obs()
        shifted = np.concatenate([new[:, None], self.hist[:, :-1]], axis=1)
        self.hist = np.where(m[:, None, None], shifted, self.hist)
        self.hist_len = np.where(m, np.minimum(self.hist_len + 1, 100), self.hist_len)
        self.ctrl_obs = np.where(m[:, None], self._delay_obs(self.ctrl_latency), self.ctrl_obs)

    def get_motor_angles(self):
        return xf.MapToMinusPiToPi(self.ctrl_obs[:, 0:12])           # minitaur.py:555-567

    def get_base_rpy(self):
        return xf.bullet_euler_from_quaternion(self.ctrl_obs[:, 36:40])   # minitaur.py:513-527

    def get_base_orientation(self):
        return xf.bullet_quaternion_from_euler(self.get_base_rpy())       # minitaur.py:630-638

    # ---- motion clock and reference state (imitation_task.py) -----------------------------
    def motion_time(self):
        """imitation_task.py:831-848 (episode_start_time_offset is always -0.0)."""
        time = self.substep_counter * self.spec.time_step
        time = time + self.time_offset
        time = time + (-0.0)
        return np.where(self.warmup, time - self.task["warmup_time"], time)

    def _warmup_pose(self):
        """imitation_task.py:985-1009."""
        pose0 = self.lib.calc_frame(self.clip, np.zeros(self.N))
        wp = np.tile(self.default_pose, (self.N, 1))
        dh = xf.calc_heading(pose0[:, 3:7]) - xf.calc_heading(wp[:, 3:7])
        rot = xf.quaternion_about_axis(dh, [0, 0, 1])
        wp[:, 0:3] = xf.QuaternionRotatePoint(wp[:, 0:3], rot)
        wp[:, 3:7] = xf.quaternion_multiply(rot, wp[:, 3:7])
        return wp

    def calc_ref_pose(self, time, apply_origin_offset=True):
        """imitation_task.py:920-953."""
        wt = self.task["warmup_time"]
        use_w = self.warmup & (time >= -wt) & (time < 0.0)
        pose = self.lib.calc_frame(self.clip, time)
        if np.any(use_w):
            pose = np.where(use_w[:, None], self._warmup_pose(), pose)
        if apply_origin_offset:
            rot = xf.quaternion_multiply(self.origin_rot, pose[:, 3:7])
            p = xf.QuaternionRotatePoint(pose[:, 0:3], self.origin_rot) + self.origin_pos
            pose = np.concatenate([p, rot, pose[:, 7:]], axis=-1)
        return pose

    def calc_ref_vel(self, time):
        """imitation_task.py:955-983, 1011-1018."""
        wt = self.task["warmup_time"]
        use_w = self.warmup & (time >= -wt) & (time < 0.0)
        vel = self.lib.calc_frame_vel(self.clip, time)
        vel = np.where(use_w[:, None], 0.0, vel)
        lin = xf.QuaternionRotatePoint(vel[:, 0:3], self.origin_rot)
        ang = xf.QuaternionRotatePoint(vel[:, 3:6], self.origin_rot)
        return np.concatenate([lin, ang, vel[:, 6:]], axis=-1)

    def _heading_rel_default(self, rot):
        """imitation_task.py:1143-1166."""
        inv_def = xf.quaternion_conjugate(self.default_pose[3:7])
        return xf.calc_heading(xf.quaternion_multiply(rot, inv_def))

    # ---- reset (quadruped_gym_env.py:63-104, minitaur.py:232-278, imitation_task.py:166-199) ----
    def reset(self, mask=None, draws=None, time_offset=None, rand_uniforms=None):
        N, s = self.N, self.spec
        m = np.ones(N, dtype=bool) if mask is None else np.asarray(mask, dtype=bool)
        mm = m[:, None]
        # -- Minitaur.reset
        self.pos = np.where(mm, s.init_pos, self.pos)
        self.quat = np.where(mm, s.init_quat, self.quat)
        self.lin = np.where(mm, 0.0, self.lin); self.ang = np.where(mm, 0.0, self.ang)
        q_def = np.zeros(12)
        q_def[self.motor_q] = s.init_motor + s.joint_offset            # minitaur.py:480-483 (no direction sign)
        self.q = np.where(mm, q_def, self.q); self.qd = np.where(mm, 0.0, self.qd)
        self.hist_len = np.where(m, 0, self.hist_len)
        self.step_counter = np.where(m, 0, self.step_counter)
        self.substep_counter = np.where(m, 0, self.substep_counter)
        self.env_step_counter = np.where(m, 0, self.env_step_counter)
        self.has_prev = np.where(m, False, self.has_prev)
        self.last_action = np.where(mm, 0.0, self.last_action)
        self.obs_torque = np.where(mm, 0.0, self.obs_torque)   # NOTE: reference keeps the stale torque; unobserved
        self.receive_obs(m)
        self.filt_x = np.where(m[:, None, None], 0.0, self.filt_x)     # action_filter.py:103-109
        self.filt_y = np.where(m[:, None, None], 0.0, self.filt_y)
        # sensors.on_reset: 3 copies of the current reading (sensor_wrappers.py:122-129)
        imu, la, ma = self._sensor_readings()
        self.imu_hist = np.where(m[:, None, None], imu[:, None], self.imu_hist)
        self.act_hist = np.where(m[:, None, None], la[:, None], self.act_hist)
        self.ang_hist = np.where(m[:, None, None], ma[:, None], self.ang_hist)
        # -- randomizers (minitaur.py:273-276)
        if self.randomizer:
            from . import randomizer as orand
            u = philox.randomizer_draws(self.seed, self.gid, self.episode) if rand_uniforms is None else np.asarray(rand_uniforms)
            vals = orand.sample(u, params=self.randomizer_ranges)
            a = orand.apply(vals, self.ids, np.asarray(self.tree.mass, dtype=np.float64), None)
            self.link_mass = np.where(mm, a["mass"], self.link_mass)
            self.inertia_scale = np.where(mm, a["inertia_scale"], self.inertia_scale)
            self.motor_strength = np.where(mm, a["motor_strength"], self.motor_strength)
            self.foot_friction = np.where(m, a["foot_friction"], self.foot_friction)
            self.ctrl_latency = np.where(m, a["latency"], self.ctrl_latency)
            jf = np.zeros((N, 12))
            for link_id in range(a["joint_friction"].shape[1]):
                if self.tree.movable[link_id]:
                    jf[:, self.tree.qidx[link_id]] = a["joint_friction"][:, link_id]
            self.joint_friction = np.where(mm, jf, self.joint_friction)
            self.rand_values = vals
        # -- ImitationTask.reset
        if draws is None:
            draws = philox.reset_draws(self.seed, self.gid, self.episode, self.lib.n_clips)
        u_rsi, clip, u_time = draws
        ref_init = u_rsi < self.task["ref_state_init_prob"]                      # :183-184
        warm = (~ref_init) & (self.task["warmup_time"] > 0)                      # :186-188
        self.warmup = np.where(m, warm, self.warmup)
        self.clip = np.where(m, clip, self.clip)                                 # :701
        dur = self.lib.duration[self.clip]
        off = np.where(self.warmup, 0 + (self.task["warmup_time"] - 0) * u_time, 0 + (dur - 0) * u_time)  # :1103-1123
        if time_offset is not None:      # replay of recorded reference draws (tests): exact offset instead of dur * u
            off = np.asarray(time_offset, dtype=np.float64)
        if not self.task["enable_rand_init_time"]:                               # :1104-1105
            off = np.zeros(N)
        self.time_offset = np.where(m, off, self.time_offset)
        self.episode = np.where(m, self.episode + 1, self.episode)
        # _reset_ref_motion :694-732
        self.origin_rot = np.where(mm, np.array([0.0, 0, 0, 1]), self.origin_rot)
        self.origin_pos = np.where(mm, 0.0, self.origin_pos)
        time = self.motion_time()
        ref_pose = self.calc_ref_pose(time)
        op = self.pos - ref_pose[:, 0:3]
        op[:, 2] = 0
        dh = xf.calc_heading(self.quat) - xf.calc_heading(ref_pose[:, 3:7])
        self.origin_pos = np.where(mm, op, self.origin_pos)
        self.origin_rot = np.where(mm, xf.quaternion_about_axis(dh, [0, 0, 1]), self.origin_rot)
        self.ref_pose = np.where(mm, self.calc_ref_pose(time), self.ref_pose)
        self.ref_vel = np.where(mm, self.calc_ref_vel(time), self.ref_vel)
        self.prev_phase = np.where(m, self.lib.calc_phase(self.clip, time), self.prev_phase)
        # _reset_clip_change_time :1057-1067 (the last call of _reset_ref_motion, :730)
        tk = self.task
        switching = np.isfinite(tk["clip_time_min"]) and np.isfinite(tk["clip_time_max"])
        need_opt = switching or tk["perturb_init_state_prob"] > 0.0
        if need_opt:
            u_perturb, u_dur, z, axis = philox.option_reset_draws(self.seed, self.gid, self.episode - 1)
        change_time = np.full(N, np.inf)
        if switching:
            change_time = time + self._draw("u", tk["clip_time_min"] + (tk["clip_time_max"] - tk["clip_time_min"]) * u_dur, m)
        self.clip_change_time = np.where(m, change_time, self.clip_change_time)
        self.clip_changes = np.where(m, 0, self.clip_changes)
        # :192-195
        pose, vel = self.ref_pose, self.ref_vel
        if tk["perturb_init_state_prob"] > 0.0:
            perturb = m & (self._draw("u", 0.0 + (1.0 - 0.0) * u_perturb, m) < tk["perturb_init_state_prob"])
            if np.any(perturb):
                pose, vel = self._apply_state_perturb(pose, vel, perturb, z, axis)
        # _sync_sim_model :778-793 / _set_state :795-829
        self.pos = np.where(mm, pose[:, 0:3], self.pos)
        self.quat = np.where(mm, pose[:, 3:7], self.quat)
        self.q = np.where(mm, pose[:, 7:], self.q)
        self.lin = np.where(mm, vel[:, 0:3], self.lin)
        self.ang = np.where(mm, vel[:, 3:6], self.ang)
        self.qd = np.where(mm, vel[:, 6:], self.qd)
        self.receive_obs(m)
        return self.observation()

    def _apply_state_perturb(self, pose, vel, mask, z, axis):
        """imitation_task.py:1199-1243; z: 32 standard normals, axis: 3 uniforms in (-1, 1) per robot, consumed in the
        reference's order (root xy, axis, angle, joint pose, root velocity xy, angular velocity, joint velocity)."""
        root_pos_std, root_rot_std, joint_pose_std = 0.025, 0.025 * np.pi, 0.05 * np.pi
        root_vel_std, root_ang_vel_std, joint_vel_std = 0.1, 0.05 * np.pi, 0.05 * np.pi
        pose, vel = pose.copy(), vel.copy()
        mm = mask[:, None]
        dx = self._draw("n", root_pos_std * z[:, 0] + 0, mask)
        dy = self._draw("n", root_pos_std * z[:, 1] + 0, mask)
        ax = self._draw("u", axis, mask)
        ax = ax / np.sqrt(np.sum(ax * ax, axis=-1, keepdims=True))
        theta = self._draw("n", root_rot_std * z[:, 2] + 0, mask)
        dj = self._draw("n", joint_pose_std * z[:, 3:15] + 0, mask)
        dvx = self._draw("n", root_vel_std * z[:, 15] + 0, mask)
        dvy = self._draw("n", root_vel_std * z[:, 16] + 0, mask)
        dw = self._draw("n", root_ang_vel_std * z[:, 17:20] + 0, mask)
        djv = self._draw("n", joint_vel_std * z[:, 20:32] + 0, mask)
        p = pose[:, 0:3].copy(); p[:, 0] += dx; p[:, 1] += dy
        rot = xf.quaternion_multiply(xf.quaternion_about_axis(theta, ax), pose[:, 3:7])
        v = vel[:, 0:3].copy(); v[:, 0] += dvx; v[:, 1] += dvy
        out_pose = np.concatenate([p, rot, pose[:, 7:] + dj], axis=-1)
        out_vel = np.concatenate([v, vel[:, 3:6] + dw, vel[:, 6:] + djv], axis=-1)
        return np.where(mm, out_pose, pose), np.where(mm, out_vel, vel)

    # ---- sensors (robot_sensors.py:74-84,153-190; environment_sensors.py:59-66) ------------
    def _sensor_readings(self):
        rpy = self.get_base_rpy()
        drpy = self.ctrl_obs[:, 40:43]
        imu = np.stack([rpy[:, 0], rpy[:, 1], drpy[:, 0], drpy[:, 1]], axis=-1)
        return imu, self.last_action.copy(), self.get_motor_angles()

    def _sensors_on_step(self):
        imu, la, ma = self._sensor_readings()
        self.imu_hist = np.concatenate([imu[:, None], self.imu_hist[:, :-1]], axis=1)
        self.act_hist = np.concatenate([la[:, None], self.act_hist[:, :-1]], axis=1)
        self.ang_hist = np.concatenate([ma[:, None], self.ang_hist[:, :-1]], axis=1)

    def observation(self):
        """minitaur.py:529-541 (alphabetical sensor order) + wrapper_env.py:109-125."""
        N = self.N
        return np.concatenate([self.imu_hist.reshape(N, 12), self.act_hist.reshape(N, 36),
                               self.ang_hist.reshape(N, 36), self.build_target_obs()], axis=-1)

    # ---- target observation (imitation_task.py:254-301) -----------------------------------
    def build_target_obs(self):
        time0 = self.motion_time()
        dt = self.spec.action_repeat * self.spec.time_step          # quadruped_gym_env.py:177
        ref_base_pos = self.ref_pose[:, 0:3]
        heading = xf.calc_heading(self.get_base_orientation())
        if self.task["tar_obs_noise"] is not None:                   # :274-275
            zn = philox.tar_noise_draw(self.seed, self.gid, self.episode - 1, self.env_step_counter)
            heading = heading + self._draw("n", self.task["tar_obs_noise"][0] * zn + 0, np.ones(self.N, dtype=bool))
        inv_rot = xf.quaternion_about_axis(-heading, [0, 0, 1])
        out = []
        for step in self.task["tar_frame_steps"]:
            pose = self.calc_ref_pose(time0 + step * dt)
            p = xf.QuaternionRotatePoint(pose[:, 0:3] - ref_base_pos, inv_rot)
            r = xf.standardize_quaternion(xf.quaternion_multiply(inv_rot, pose[:, 3:7]))
            out.append(np.concatenate([p, r, pose[:, 7:]], axis=-1))
        return np.concatenate(out, axis=-1)

    def target_obs_bounds(self):
        """imitation_task.py:303-335."""
        low, high = self.lib.joint_bounds()
        low = low.copy(); high = high.copy()
        low[0:3] = -2; high[0:3] = 2; low[3:7] = -1; high[3:7] = 1
        n = len(self.task["tar_frame_steps"])
        return np.tile(low, n), np.tile(high, n)

    def observation_bounds(self):
        """robot_sensors.py IMU bounds, LastAction +-1, MotorAngle +-pi, then the target bounds."""
        imu = np.array([2 * np.pi, 2 * np.pi, 2000 * np.pi, 2000 * np.pi] * 3)
        hi = np.concatenate([imu, np.ones(36), np.pi * np.ones(36)])
        tl, th = self.target_obs_bounds()
        return np.concatenate([-hi, tl]), np.concatenate([hi, th])

    # ---- physics substep (replaces pybullet.stepSimulation, quadruped_gym_env.py:223) -----
    def joint_torques(self, tau_motor):
        """Motor torque (motor space) -> URDF-space generalised joint forces incl. limits/damping."""
        t, p = self.tree, self.phys
        tau = np.zeros((self.N, 12))
        tau[:, self.motor_q] = tau_motor * self.spec.joint_dir       # minitaur.py:755-756
        lo, up, damp = t.lower[self.revolute_links], t.upper[self.revolute_links], t.damping[self.revolute_links]
        e = np.maximum(0.0, lo - self.q) - np.maximum(0.0, self.q - up)
        tau = tau + p["k_lim"] * e - p["d_lim"] * np.minimum(1.0, np.abs(e) / p["e0"]) * self.qd - damp * self.qd
        # set_joint_friction (minitaur.py:1063-1070): Bullet's velocity motor with a force limit, here a regularised
        # Coulomb torque  -clip(k qd, -F, F)
        tau = tau - np.clip(p["joint_friction_k"] * self.qd, -self.joint_friction, self.joint_friction)
        return tau

    def contact_forces(self, k, nu):
        """Semi-implicit penalty contact model (DESIGN.md "physics model").

        Per collision point: penetration d = r - P_z; the force at the START-of-substep velocity is
        F0 = (-c_t V_x, -c_t V_y, max(0, kn d - bn s V_z)), s = min(1, d/d0),
        c_t = min(bt, mu F0_z / |V_t|); its velocity dependence is treated implicitly with the frozen
        diagonal D = diag(c_t, c_t, kn dt + bn s):  F_new = F0 - D (V_new - V_old).
        Returns generalised force J^T F0 (N,18), the implicit matrix dt * sum J^T D J (N,18,18) and heights.
        """
        p = self.phys
        dt = self.spec.time_step
        gen = np.zeros((self.N, 18))
        W = np.zeros((self.N, 18, 18))
        heights = np.zeros((self.N, len(self.cp_body)))
        for c in range(len(self.cp_body)):
            b = self.cp_body[c]
            P = k["x"][:, b] + np.einsum("nij,j->ni", k["R"][:, b], self.cp_local[c])
            r = self.cp_r[c]
            heights[:, c] = P[:, 2] - r
            d = r - P[:, 2]
            act = d > 0
            if not np.any(act):
                continue
            Pc = P - np.array([0, 0, r])
            J = self.tree.point_jacobian(k, b, Pc)
            V = np.einsum("nia,na->ni", J, nu)
            sc = np.minimum(1.0, d / p["d0"])
            fn = np.maximum(0.0, p["kn"] * d - p["bn"] * sc * V[:, 2])
            vt = np.sqrt(V[:, 0] ** 2 + V[:, 1] ** 2)
            mu = np.where(self.cp_foot[c] & (self.foot_friction >= 0), self.foot_friction * p["mu_plane"], self.cp_mu[c])   # set_foot_friction
            ct = np.minimum(p["bt"], mu * fn / np.maximum(vt, 1e-9))
            cn = np.where(fn > 0.0, p["kn"] * dt + p["bn"] * sc, 0.0)
            F = np.stack([-ct * V[:, 0], -ct * V[:, 1], fn], axis=-1)
            D = np.stack([ct, ct, cn], axis=-1)
            F = np.where(act[:, None], F, 0.0)
            D = np.where(act[:, None], D, 0.0)
            gen += np.einsum("nia,ni->na", J, F)
            W += dt * np.einsum("nia,ni,nib->nab", J, D, J)
        return gen, W, heights

    def forward_dynamics(self, tau_joint):
        t = self.tree
        nu = np.concatenate([self.ang, self.lin, self.qd], axis=-1)
        k = t.fk(self.pos, self.quat, self.q)
        M = t.mass_matrix(self.pos, self.quat, self.q, mass=self.link_mass, inertia_scale=self.inertia_scale)
        C = t.bias(self.pos, self.quat, self.q, nu, self.phys["gravity"], mass=self.link_mass, inertia_scale=self.inertia_scale, k=k)
        gen_c, W, _ = self.contact_forces(k, nu)
        rhs = gen_c - C
        rhs[:, 6:] += tau_joint
        return np.linalg.solve(M + W, rhs[..., None])[..., 0]

    def integrate(self, nudot):
        dt = self.spec.time_step
        # every generalised velocity coordinate is clamped like btMultiBody::applyDeltaVeeMultiDof
        # (m_maxCoordinateVelocity = 100; Bullet3 knowledge, not in /root/reference): keeps the state finite when
        # the reference's own velocity tables spike (quaternion sign flips in laikago_spin / hopturn / turn give
        # |w| up to 375 rad/s at reference-state initialisation)
        vmax = self.phys["max_coordinate_velocity"]
        self.ang = np.clip(self.ang + dt * nudot[:, 0:3], -vmax, vmax)
        self.lin = np.clip(self.lin + dt * nudot[:, 3:6], -vmax, vmax)
        self.qd = np.clip(self.qd + dt * nudot[:, 6:], -vmax, vmax)
        self.pos = self.pos + dt * self.lin
        wn = np.sqrt(np.sum(self.ang ** 2, axis=-1))
        half = 0.5 * wn * dt
        k = np.where(wn > 1e-12, np.sin(half) / np.where(wn > 1e-12, wn, 1.0), 0.5 * dt)
        dq = np.concatenate([self.ang * k[:, None], np.cos(half)[:, None]], axis=-1)
        qn = xf.quaternion_multiply(dq, self.quat)
        self.quat = qn / np.sqrt(np.sum(qn ** 2, axis=-1, keepdims=True))
        self.q = self.q + dt * self.qd

    # ---- action path (minitaur.py:280-293, 438-460, 706-769; minitaur_motor.py:104-173) ----
    def set_act(self, action):
        s = self.spec
        a = np.array(action, dtype=np.float64) + s.init_motor         # :281
        self.last_action = a
        first = self.substep_counter == 0                              # :1173-1175
        if np.any(first):
            d = self.get_motor_angles()
            f3 = first[:, None, None]
            self.filt_x = np.where(f3, d[:, None], self.filt_x)
            self.filt_y = np.where(f3, d[:, None], self.filt_y)
        # action_filter.py:111-120
        y = a * BUTTER_B[0] + np.sum(self.filt_x * BUTTER_B[1:][None, :, None], axis=1) \
            - np.sum(self.filt_y * BUTTER_A[1:][None, :, None], axis=1)
        self.filt_x = np.concatenate([a[:, None], self.filt_x[:, :1]], axis=1)
        self.filt_y = np.concatenate([y[:, None], self.filt_y[:, :1]], axis=1)
        self.action = y

    def motor_torque(self, substep):
        s = self.spec
        prev = np.where(self.has_prev[:, None], self.prev_filtered, self.get_motor_angles())   # :450-453
        lerp = float(substep + 1) / s.action_repeat
        cmd = prev + lerp * (self.action - prev)
        cur = self.get_motor_angles()
        cmd = np.clip(cmd, cur - s.max_angle_change, cur + s.max_angle_change)                  # :706-723
        q, qdot = self.hist[:, 0, 0:12], self.hist[:, 0, 12:24]                                 # pd latency 0, :359-363
        tau = -1 * (s.kp * (q - cmd)) - s.kd * (qdot - 0.0) + 0.0                               # minitaur_motor.py:163-165
        return self.motor_strength * tau

    def substep(self, i):
        tau_m = self.motor_torque(i)
        self.obs_torque = tau_m
        nudot = self.forward_dynamics(self.joint_torques(tau_m))
        self.substep_counter = self.substep_counter + 1
        if i == self.spec.action_repeat - 1:
            self.prev_filtered = self.action.copy()
            self.has_prev[:] = True
            self.step_counter = self.step_counter + 1
        self.integrate(nudot)
        self.receive_obs()

    # ---- reward (imitation_task.py:341-516) ------------------------------------------------
    def end_effectors(self, pos, quat, q):
        c = self.tree.link_com(pos, quat, q)
        return c[:, self.foot_links]

    def reward_terms(self):
        tk = self.task
        ref = self.ref_pose
        pose_err = np.sum((ref[:, 7:] - self.q) ** 2, axis=-1)
        vel_err = np.sum((self.ref_vel[:, 6:] - self.qd) ** 2, axis=-1)
        inv_def = xf.quaternion_conjugate(self.default_pose[3:7])
        h_ref = xf.calc_heading_rot(xf.quaternion_multiply(ref[:, 3:7], inv_def))
        h_sim = xf.calc_heading_rot(xf.quaternion_multiply(self.quat, inv_def))
        ih_ref, ih_sim = xf.quaternion_conjugate(h_ref), xf.quaternion_conjugate(h_sim)
        ee_ref = self.end_effectors(ref[:, 0:3], ref[:, 3:7], ref[:, 7:])
        ee_sim = self.end_effectors(self.pos, self.quat, self.q)
        ee_err = np.zeros(self.N)
        for j in range(ee_ref.shape[1]):
            rr = xf.QuaternionRotatePoint(ee_ref[:, j] - ref[:, 0:3], ih_ref)
            rs = xf.QuaternionRotatePoint(ee_sim[:, j] - self.pos, ih_sim)
            d = rr - rs
            dh = ee_ref[:, j, 2] - ee_sim[:, j, 2]
            ee_err += d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + tk["end_effector_height_err_scale"] * dh * dh
        dp = ref[:, 0:3] - self.pos
        root_pos_err = np.sum(dp * dp, axis=-1)
        dq = xf.quaternion_multiply(ref[:, 3:7], xf.quaternion_conjugate(self.quat))
        _, ang = xf.QuaternionToAxisAngle(dq)
        ang = xf.normalize_rotation_angle(ang)
        root_pose_err = root_pos_err + 0.5 * ang * ang
        dv = self.ref_vel[:, 0:3] - self.lin
        dw = self.ref_vel[:, 3:6] - self.ang
        root_vel_err = np.sum(dv * dv, axis=-1) + 0.1 * np.sum(dw * dw, axis=-1)
        return np.stack([np.exp(-tk["pose_err_scale"] * pose_err),
                         np.exp(-tk["velocity_err_scale"] * vel_err),
                         np.exp(-tk["end_effector_err_scale"] * ee_err),
                         np.exp(-tk["root_pose_err_scale"] * root_pose_err),
                         np.exp(-tk["root_velocity_err_scale"] * root_vel_err)], axis=-1)

    def reward(self):
        tk = self.task
        w = np.array([tk["pose_weight"], tk["velocity_weight"], tk["end_effector_weight"],
                      tk["root_pose_weight"], tk["root_velocity_weight"]])
        t = self.reward_terms()
        return (t[:, 0] * w[0] + t[:, 1] * w[1] + t[:, 2] * w[2] + t[:, 3] * w[3] + t[:, 4] * w[4]) * tk["weight"]   # :350-356

    # ---- task update (imitation_task.py:201-207, 734-761, 1020-1055) -----------------------
    def update(self):
        tk = self.task
        time = self.motion_time()
        change = (time >= self.clip_change_time) & (self.lib.n_clips > 1)          # _check_change_clip :1096-1101
        if np.any(change):                                                       # :742-746
            clip_new, u_dur, u_off = philox.clip_change_draws(self.seed, self.gid, self.episode - 1, self.clip_changes, self.lib.n_clips)
            self.clip = np.where(change, self._draw("i", clip_new, change), self.clip)
            # _reset_clip_change_time reads the motion clock BEFORE the new time offset is drawn: still `time`
            dur_v = self._draw("u", tk["clip_time_min"] + (tk["clip_time_max"] - tk["clip_time_min"]) * u_dur, change)
            self.clip_change_time = np.where(change, time + dur_v, self.clip_change_time)
            off = self._draw("u", 0 + (self.lib.duration[self.clip] - 0) * u_off, change)      # _sample_time_offset :1112-1123
            self.time_offset = np.where(change, off, self.time_offset)
            self.clip_changes = self.clip_changes + change
        new_phase = self.lib.calc_phase(self.clip, time)            # :748-749: the NEW clip's phase at the OLD clock value
        sync = (tk["enable_cycle_sync"] & (new_phase < self.prev_phase)) | change      # :751-754
        if np.any(sync):
            time_s = self.motion_time()                             # _sync_ref_origin re-reads the clock (:1033)
            ref = self.calc_ref_pose(time_s, apply_origin_offset=False)
            if np.any(change):                                      # sync_root_rotation=change_clip :1037-1045
                dh = self._heading_rel_default(self.quat) - self._heading_rel_default(ref[:, 3:7])
                self.origin_rot = np.where(change[:, None], xf.quaternion_about_axis(dh, [0, 0, 1]), self.origin_rot)
            rp = xf.QuaternionRotatePoint(ref[:, 0:3], self.origin_rot)
            op = self.pos - rp
            op[:, 2] = 0
            self.origin_pos = np.where(sync[:, None], op, self.origin_pos)
        time = self.motion_time()                                   # _update_ref_state :763-768
        self.ref_pose = self.calc_ref_pose(time)
        self.ref_vel = self.calc_ref_vel(time)
        self.prev_phase = new_phase

    # ---- termination (imitation_task.py:518-572, wrapper_env.py:78-80) ---------------------
    def contact_fall(self):
        k = self.tree.fk(self.pos, self.quat, self.q)
        fall = np.zeros(self.N, dtype=bool)
        self._fall_margin = np.full(self.N, np.inf)
        for c in range(len(self.cp_body)):
            if self.cp_foot[c]:
                continue
            b = self.cp_body[c]
            P = k["x"][:, b] + np.einsum("nij,j->ni", k["R"][:, b], self.cp_local[c])
            h = P[:, 2] - self.cp_r[c] - self.cp_margin[c]
            fall |= h < 0
            self._fall_margin = np.minimum(self._fall_margin, np.abs(h))
        return fall & (self.step_counter > 0)

    def terminal_terms(self):
        time = self.motion_time()
        over = self.lib.is_over(self.clip, time)
        fall = self.contact_fall()
        dp = self.ref_pose[:, 0:3] - self.pos
        d2 = np.sum(dp * dp, axis=-1)
        thr = self.task["dist_fail_threshold"]
        pos_fail = d2 > thr * thr
        dq = xf.quaternion_multiply(self.ref_pose[:, 3:7], xf.quaternion_conjugate(self.quat))
        _, ang = xf.QuaternionToAxisAngle(dq)
        ang = xf.normalize_rotation_angle(ang)
        rot_fail = np.abs(ang) > self.task["rot_fail_threshold"]
        # distance of each test from its threshold (tests skip near-threshold samples)
        self._term_margin = np.minimum(np.abs(d2 - thr * thr), np.abs(np.abs(ang) - self.task["rot_fail_threshold"]))
        return over, fall, pos_fail, rot_fail

    # ---- one control step (quadruped_gym_env.py:213-239, wrapper_env.py:58-85) -------------
    def step(self, action):
        self.set_act(action)
        for i in range(self.spec.action_repeat):
            self.substep(i)
        self._sensors_on_step()
        reward = self.reward()
        self.update()
        over, fall, pos_fail, rot_fail = self.terminal_terms()
        done = over | fall | pos_fail | rot_fail
        self.env_step_counter = self.env_step_counter + 1
        obs = self.observation()
        done = done | (self.env_step_counter >= self.max_episode_steps)
        return obs, reward, done

    # ---- state import/export for parity tests ----------------------------------------------
    STATE_FIELDS = ("pos", "quat", "q", "lin", "ang", "qd", "substep_counter", "step_counter",
                    "env_step_counter", "episode", "filt_x", "filt_y", "prev_filtered", "has_prev", "action",
                    "last_action", "hist", "hist_len", "ctrl_obs", "obs_torque", "imu_hist", "act_hist",
                    "ang_hist", "clip", "time_offset", "warmup", "origin_pos", "origin_rot", "prev_phase",
                    "ref_pose", "ref_vel", "motor_strength", "clip_change_time", "clip_changes")

    def get_state(self):
        return {k: np.array(getattr(self, k)) for k in self.STATE_FIELDS}

    def set_state(self, st):
        for k, v in st.items():
            setattr(self, k, np.array(v))
