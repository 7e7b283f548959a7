"""Motion-file format reader (host side).

The JSON format is the reference's (OpenRoboRL/envs/quadruped_robot/task/motion_data.py:46-50, 72-100):
keys "LoopMode" ("Wrap"|"Clamp"), "FrameDuration", optional "EnableCycleOffsetPosition" /
"EnableCycleOffsetRotation", "Frames" = rows of [x y z qx qy qz qw joints...].  Only parsing happens
here; the post-processing (first-frame xy subtraction, quaternion normalisation/standardisation, frame
velocities, cycle deltas) is done by `orl_load_motions` in the C library.
"""
import json
import os

import numpy as np

FLAG_WRAP, FLAG_CYCLE_POS, FLAG_CYCLE_ROT = 1, 2, 4
_LOOP_MODES = ("Clamp", "Wrap")


class MotionClip(object):
    def __init__(self, frames, frame_duration, loop_mode="Wrap", cycle_pos=False, cycle_rot=False, name=""):
        if loop_mode not in _LOOP_MODES:
            raise KeyError(loop_mode)                     # LoopMode[...] lookup, motion_data.py:85
        self.frames = np.ascontiguousarray(frames, dtype=np.float64)
        self.frame_duration = float(frame_duration)
        self.loop_mode = loop_mode
        self.cycle_pos = bool(cycle_pos)
        self.cycle_rot = bool(cycle_rot)
        self.name = name
        assert self.frames.ndim == 2 and self.frames.shape[0] > 0, "Must have at least 1 frame."
        assert self.frames.shape[1] > 7, "Frames have too few degrees of freedom."
        assert self.frame_duration > 0, "Frame duration must be positive."

    @property
    def flags(self):
        return ((FLAG_WRAP if self.loop_mode == "Wrap" else 0) | (FLAG_CYCLE_POS if self.cycle_pos else 0)
                | (FLAG_CYCLE_ROT if self.cycle_rot else 0))

    @classmethod
    def from_json(cls, text, name=""):
        mj = json.loads(text)
        return cls(mj["Frames"], mj["FrameDuration"], mj["LoopMode"],
                   mj.get("EnableCycleOffsetPosition", False), mj.get("EnableCycleOffsetRotation", False), name)

    @classmethod
    def load(cls, motion_file):
        with open(motion_file, "r") as f:
            return cls.from_json(f.read(), name=motion_file)

    def to_json(self):
        return json.dumps({"LoopMode": self.loop_mode, "FrameDuration": self.frame_duration,
                           "EnableCycleOffsetPosition": self.cycle_pos, "EnableCycleOffsetRotation": self.cycle_rot,
                           "Frames": self.frames.tolist()})


CLIP_LIBRARY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "motions", "clips.npz")


def bundled_clips(names=None):
    """The motion clips shipped with the package (the reference's clips under envs/quadruped_robot/task/motions/,
    re-serialised by scripts/make_golden.py: raw "Frames" + metadata; post-processing happens in orl_load_motions)."""
    return load_clip_library(CLIP_LIBRARY, names)


def load_clip_library(npz_path, names=None):
    """Clips stored as arrays (assets/motions/clips.npz layout: <name>/frames, <name>/meta)."""
    z = np.load(npz_path)
    all_names = sorted({k.split("/")[0] for k in z.files})
    out = []
    for n in (names or all_names):
        fd, wrap, cp, cr = z[n + "/meta"]
        out.append(MotionClip(z[n + "/frames"], float(fd), "Wrap" if wrap else "Clamp", bool(cp), bool(cr), name=n))
    return out
