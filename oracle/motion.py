"""Reference-motion library: float64 restatement of the reference's motion_data.py, batched.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
`OpenRoboRL/envs/quadruped_robot/task/motion_data.py` (line numbers per function).
A `MotionLib` holds several clips in concatenated tables; every query takes a per-robot
clip id array and a per-robot time array (float64) so a whole batch is evaluated at once.
"""
import json
import numpy as np

from . import xform as xf

POS_SIZE, ROT_SIZE, VEL_SIZE, ANG_VEL_SIZE = 3, 4, 3, 3


def parse_motion_json(text_or_dict):
    """motion_data.py:72-100: keys LoopMode, FrameDuration, EnableCycleOffset*, Frames."""
    mj = json.loads(text_or_dict) if isinstance(text_or_dict, (str, bytes)) else text_or_dict
    loop = mj["LoopMode"]
    if loop not in ("Clamp", "Wrap"):
        raise KeyError(loop)
    return dict(
        wrap=(loop == "Wrap"),
        frame_duration=float(mj["FrameDuration"]),
        cycle_pos=bool(mj.get("EnableCycleOffsetPosition", False)),
        cycle_rot=bool(mj.get("EnableCycleOffsetRotation", False)),
        frames=np.array(mj["Frames"], dtype=np.float64))


def postprocess_frames(frames):
    """motion_data.py:527-556: subtract first-frame xy, normalise + standardise quats (in place copy)."""
    frames = np.array(frames, dtype=np.float64)
    if frames.shape[0] > 0:
        start = frames[0, :2].copy()
        frames[:, 0] -= start[0]
        frames[:, 1] -= start[1]
        for f in range(frames.shape[0]):
            # per-row np.linalg.norm of the 1-D quaternion, exactly as pose3d.QuaternionNormalize does
            # (its dot-product summation order differs from a batched axis reduction by 1 ulp)
            q = frames[f, 3:7] / np.linalg.norm(frames[f, 3:7])
            frames[f, 3:7] = xf.standardize_quaternion(q)
    return frames


def calc_frame_vels(frames, dt):
    """motion_data.py:635-680: forward differences; last row copies second-last."""
    F, S = frames.shape
    vels = np.zeros((F, S - POS_SIZE - ROT_SIZE + VEL_SIZE + ANG_VEL_SIZE))
    if F > 1:
        f0, f1 = frames[:-1], frames[1:]
        vels[:-1, 0:3] = (f1[:, 0:3] - f0[:, 0:3]) / dt
        dq = xf.quaternion_multiply(f1[:, 3:7], xf.quaternion_conjugate(f0[:, 3:7]))
        axis, angle = xf.QuaternionToAxisAngle(dq)
        vels[:-1, 3:6] = (angle / dt)[:, None] * axis
        vels[:-1, 6:] = (f1[:, 7:] - f0[:, 7:]) / dt
        vels[-1] = vels[-2]
    return vels


class MotionLib(object):
    """Concatenated clips.  All per-clip scalars are float64 arrays indexed by clip id."""

    def __init__(self, clips):
        """clips: list of dicts from parse_motion_json (raw, un-postprocessed frames)."""
        self.n_clips = len(clips)
        if self.n_clips == 0:
            raise ValueError("No reference motions specified.")  # imitation_task.py:583-584
        self.frames, self.vels = [], []
        self.start = np.zeros(self.n_clips, dtype=np.int64)
        self.num_frames = np.zeros(self.n_clips, dtype=np.int64)
        self.frame_dur = np.zeros(self.n_clips)
        self.wrap = np.zeros(self.n_clips, dtype=bool)
        self.cycle_pos = np.zeros(self.n_clips, dtype=bool)
        self.cycle_rot = np.zeros(self.n_clips, dtype=bool)
        self.delta_pos = np.zeros((self.n_clips, 3))
        self.delta_heading = np.zeros(self.n_clips)
        off = 0
        for c, clip in enumerate(clips):
            fr = postprocess_frames(clip["frames"])
            assert fr.shape[0] > 0 and fr.shape[1] > POS_SIZE + ROT_SIZE and clip["frame_duration"] > 0
            self.frames.append(fr)
            self.vels.append(calc_frame_vels(fr, clip["frame_duration"]))
            self.start[c] = off
            self.num_frames[c] = fr.shape[0]
            self.frame_dur[c] = clip["frame_duration"]
            self.wrap[c] = clip["wrap"]
            self.cycle_pos[c] = clip["cycle_pos"]
            self.cycle_rot[c] = clip["cycle_rot"]
            # motion_data.py:558-589
            d = fr[-1, 0:3] - fr[0, 0:3]
            d[2] = 0
            self.delta_pos[c] = d
            drot = xf.quaternion_multiply(fr[-1, 3:7], xf.quaternion_conjugate(fr[0, 3:7]))
            self.delta_heading[c] = xf.calc_heading(drot)
            off += fr.shape[0]
        self.frames = np.concatenate(self.frames, axis=0)
        self.vels = np.concatenate(self.vels, axis=0)
        # motion_data.py:198-208
        self.duration = self.frame_dur * (self.num_frames - 1)

    # -- motion_data.py:210-232
    def calc_phase(self, clip, time):
        dur = self.duration[clip]
        phase = time / dur
        return np.where(self.wrap[clip], phase - np.floor(phase), np.clip(phase, 0.0, 1.0))

    # -- motion_data.py:234-253
    def calc_cycle_count(self, clip, time):
        count = np.floor(time / self.duration[clip]).astype(np.int64)
        return np.where(self.wrap[clip], count, np.clip(count, 0, 1))

    # -- motion_data.py:265-276
    def is_over(self, clip, time):
        return (~self.wrap[clip]) & (time >= self.duration[clip])

    # -- motion_data.py:682-718
    def calc_blend_idx(self, clip, time):
        clip = np.asarray(clip)
        time = np.asarray(time, dtype=np.float64)
        dur = self.duration[clip]
        nf = self.num_frames[clip]
        fd = self.frame_dur[clip]
        wrap = self.wrap[clip]
        phase = self.calc_phase(clip, time)
        f0 = (phase * (nf - 1)).astype(np.int64)  # int() truncation; phase >= 0
        f1 = np.minimum(f0 + 1, nf - 1)
        norm_time = phase * dur
        time0 = f0 * fd
        time1 = f1 * fd
        with np.errstate(divide="ignore", invalid="ignore"):
            blend = (norm_time - time0) / (time1 - time0)
        lo = (~wrap) & (time <= 0)
        hi = (~wrap) & (time >= dur) & ~lo
        f0 = np.where(lo, 0, np.where(hi, nf - 1, f0))
        f1 = np.where(lo, 0, np.where(hi, nf - 1, f1))
        blend = np.where(lo | hi, 0.0, blend)
        ok = lo | hi | ((norm_time >= time0 - 1e-5) & (norm_time <= time1 + 1e-5))
        assert np.all(ok), "motion_data.py:714 frame-time bracket"
        return f0, f1, blend

    # -- motion_data.py:591-633
    def _cycle_offset_pos(self, clip, num_cycles):
        dp = self.delta_pos[clip]
        plain = num_cycles[..., None] * dp
        out = np.where(self.cycle_rot[clip][..., None], 0.0, plain)
        if np.any(self.cycle_rot[clip]):
            acc = np.zeros_like(plain)
            for i in range(int(max(0, np.max(num_cycles)))):
                rot = xf.quaternion_about_axis(i * self.delta_heading[clip], [0, 0, 1])
                cur = xf.QuaternionRotatePoint(dp, rot)
                acc = acc + np.where((i < num_cycles)[..., None], cur, 0.0)
            out = np.where(self.cycle_rot[clip][..., None], acc, out)
        return np.where(self.cycle_pos[clip][..., None], out, 0.0)

    def _cycle_offset_rot(self, clip, num_cycles):
        heading = num_cycles * self.delta_heading[clip]
        q = xf.quaternion_from_euler(np.zeros_like(heading), np.zeros_like(heading), heading)
        ident = np.array([0.0, 0.0, 0.0, 1.0])
        return np.where(self.cycle_rot[clip][..., None], q, ident)

    # -- motion_data.py:417-449, 478-509
    def calc_frame(self, clip, time):
        clip = np.asarray(clip)
        time = np.asarray(time, dtype=np.float64)
        f0, f1, blend = self.calc_blend_idx(clip, time)
        fr0 = self.frames[self.start[clip] + f0]
        fr1 = self.frames[self.start[clip] + f1]
        b = blend[..., None]
        pos = (1.0 - b) * fr0[..., 0:3] + b * fr1[..., 0:3]
        rot = xf.standardize_quaternion(xf.quaternion_slerp(fr0[..., 3:7], fr1[..., 3:7], blend))
        joints = (1.0 - b) * fr0[..., 7:] + b * fr1[..., 7:]
        cyc = self.calc_cycle_count(clip, time)
        off_pos = self._cycle_offset_pos(clip, cyc)
        off_rot = self._cycle_offset_rot(clip, cyc)
        pos = xf.QuaternionRotatePoint(pos, off_rot) + off_pos
        rot = xf.standardize_quaternion(xf.quaternion_multiply(off_rot, rot))
        return np.concatenate([pos, rot, joints], axis=-1)

    # -- motion_data.py:451-476, 511-525
    def calc_frame_vel(self, clip, time):
        clip = np.asarray(clip)
        time = np.asarray(time, dtype=np.float64)
        f0, f1, blend = self.calc_blend_idx(clip, time)
        v0 = self.vels[self.start[clip] + f0]
        v1 = self.vels[self.start[clip] + f1]
        b = blend[..., None]
        v = (1.0 - b) * v0 + b * v1
        off_rot = self._cycle_offset_rot(clip, self.calc_cycle_count(clip, time))
        lin = xf.QuaternionRotatePoint(v[..., 0:3], off_rot)
        ang = xf.QuaternionRotatePoint(v[..., 3:6], off_rot)
        return np.concatenate([lin, ang, v[..., 6:]], axis=-1)

    def joint_bounds(self):
        """imitation_task.py:315-323: column-wise min/max over all clips' (post-processed) frames."""
        return self.frames.min(axis=0), self.frames.max(axis=0)
