"""Generate tests/golden/* from the reference tree (run in the BUILD container only).

    python scripts/make_golden.py [--reference /root/reference]

The GPU box has no /root/reference, so everything the tests and bench need from it is frozen here:

* openroborl_b200/assets/motions/clips.npz : the RAW "Frames" arrays + metadata of every clip under
                         OpenRoboRL/envs/quadruped_robot/task/motions/*.txt (input data in the reference's
                         motion-file format, re-serialised; the JSON loader itself is exercised by re-writing
                         these arrays to JSON at test time).
* policy_<name>.npz    : the actor MLP (pi_fc0, pi_fc1, pi) of the reference's pretrained policy zips, used as a
                         deterministic action source for rollouts, plus the observation-space bounds stored in
                         the zip (pins the observation layout).
* ref_motion_kat.npz   : known answers produced by the reference's OWN, UNMODIFIED modules
                         (motion_data.py, pose3d.py, action_filter.py, minitaur_motor.py) imported from the
                         reference tree with a numpy shim for the absent third-party
                         `pybullet_utils.transformations` (oracle/xform.py restates its algorithms).
* ref_randomizer_kat.npz: sampled parameters and robot-setter arguments of the reference's own
                         ControllableEnvRandomizerFromConfig (seeded RandomState, recording robot stub).
* ref_task_trace.npz   : per-step traces (observation, reward, done, task internals) produced by the
                         reference's own ImitationTask / Minitaur / LocomotionGymEnv / WrapperEnv classes running
                         against a fake BulletClient (scripts/fake_bullet.py) whose physics backend is oracle/rbd.py.
"""
import argparse
import base64
import io
import json
import os
import pickle
import sys
import types
import zipfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def install_shims():
    """Stub the third-party modules the reference imports but this container lacks."""
    from oracle import xform as xf

    tr = types.ModuleType("pybullet_utils.transformations")
    tr.quaternion_multiply = lambda a, b: xf.quaternion_multiply(np.asarray(a, float), np.asarray(b, float))
    tr.quaternion_conjugate = lambda q: xf.quaternion_conjugate(np.asarray(q, float))
    tr.quaternion_inverse = lambda q: xf.quaternion_inverse(np.asarray(q, float))
    tr.quaternion_about_axis = lambda a, ax: xf.quaternion_about_axis(float(a), np.asarray(ax, float))
    tr.quaternion_from_euler = lambda r, p, y: xf.quaternion_from_euler(float(r), float(p), float(y))
    tr.quaternion_slerp = lambda a, b, f: xf.quaternion_slerp(np.asarray(a, float), np.asarray(b, float), float(f))
    pu = types.ModuleType("pybullet_utils")
    pu.transformations = tr
    sys.modules["pybullet_utils"] = pu
    sys.modules["pybullet_utils.transformations"] = tr


def motion_clips(ref):
    mdir = os.path.join(ref, "OpenRoboRL/envs/quadruped_robot/task/motions")
    out = {}
    for fn in sorted(os.listdir(mdir)):
        if not fn.endswith(".txt"):
            continue
        mj = json.load(open(os.path.join(mdir, fn)))
        name = fn[:-4]
        out[name + "/frames"] = np.array(mj["Frames"], dtype=np.float64)
        out[name + "/meta"] = np.array([float(mj["FrameDuration"]), float(mj["LoopMode"] == "Wrap"),
                                        float(bool(mj.get("EnableCycleOffsetPosition", False))),
                                        float(bool(mj.get("EnableCycleOffsetRotation", False)))])
    # a Clamp-mode clip (the reference library only ships Wrap clips): the pace frames replayed once, no cycle offsets
    out["synthetic_pace_clamp/frames"] = out["laikago_pace/frames"].copy()
    out["synthetic_pace_clamp/meta"] = np.array([out["laikago_pace/meta"][0], 0.0, 0.0, 0.0])
    np.savez_compressed(os.path.join(ROOT, "openroborl_b200", "assets", "motions", "clips.npz"), **out)
    print("assets/motions/clips.npz:", sorted({k.split("/")[0] for k in out}))


class _BoxStub(object):
    def __setstate__(self, st):
        self.__dict__.update(st)


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("gym"):
            return _BoxStub
        return super().find_class(module, name)


def policies(ref):
    pdir = os.path.join(ref, "OpenRoboRL/envs/quadruped_robot/task/policies")
    for name in ("laikago_pace", "minicheetah_trot"):
        z = zipfile.ZipFile(os.path.join(pdir, name + ".zip"))
        p = np.load(io.BytesIO(z.read("parameters")))
        data = json.loads(z.read("data"))
        out = {k: p["model/%s:0" % k] for k in ("pi_fc0/w", "pi_fc0/b", "pi_fc1/w", "pi_fc1/b", "pi/w", "pi/b",
                                                 "vf_fc0/w", "vf_fc0/b", "vf_fc1/w", "vf_fc1/b", "vf/w", "vf/b")}
        try:
            box = _Unpickler(io.BytesIO(base64.b64decode(data["observation_space"][":serialized:"]))).load()
            out["obs_low"], out["obs_high"] = np.asarray(box.low), np.asarray(box.high)
        except Exception as exc:   # keep going: the bounds are an extra pin, not required
            print("  (observation_space of %s not decoded: %r)" % (name, exc))
        np.savez_compressed(os.path.join(GOLD, "policy_%s.npz" % name), **{k.replace("/", "_"): v for k, v in out.items()})
        print("policy_%s.npz" % name, sorted(out))


def motion_kat(ref):
    """Known answers from the reference's own motion_data / pose3d / action_filter / minitaur_motor."""
    sys.path.insert(0, os.path.join(ref, "OpenRoboRL"))
    import collections
    import collections.abc
    collections.Sequence = collections.abc.Sequence      # removed in py3.10; minitaur_motor.py:61 uses it
    from envs.quadruped_robot.task import motion_data
    from envs.utilities import pose3d
    from envs.utilities import action_filter
    from envs.quadruped_robot.robots import minitaur_motor
    mdir = os.path.join(ref, "OpenRoboRL/envs/quadruped_robot/task/motions")
    out = {}
    rng = np.random.RandomState(1234)
    import tempfile
    clamp = json.load(open(os.path.join(mdir, "laikago_pace.txt")))
    clamp["LoopMode"] = "Clamp"; clamp["EnableCycleOffsetPosition"] = False; clamp["EnableCycleOffsetRotation"] = False
    tmpd = tempfile.mkdtemp()
    json.dump(clamp, open(os.path.join(tmpd, "synthetic_pace_clamp.txt"), "w"))
    for name in ("laikago_pace", "laikago_spin", "laikago_hopturn", "minicheetah_trot", "synthetic_pace_clamp"):
        m = motion_data.MotionData(os.path.join(tmpd if name.startswith("synthetic") else mdir, name + ".txt"))
        dur = m.get_duration()
        times = np.concatenate([[0.0, 0.033, 0.5, 0.7, 1.2345, -0.1, dur, 2 * dur, dur - 1e-9],
                                rng.uniform(-0.3, 25.0, size=200)])
        rec = dict(times=times, frames=m.get_frames(), vels=m._frame_vels, duration=dur,
                   delta_pos=m._cycle_delta_pos, delta_heading=m._cycle_delta_heading)
        f0s, f1s, bl, cyc, poses, vels, phases = [], [], [], [], [], [], []
        for t in times:
            f0, f1, b = m.calc_blend_idx(t)
            f0s.append(f0); f1s.append(f1); bl.append(b); cyc.append(m.calc_cycle_count(t))
            poses.append(m.calc_frame(t)); vels.append(m.calc_frame_vel(t)); phases.append(m.calc_phase(t))
        rec.update(f0=np.array(f0s), f1=np.array(f1s), blend=np.array(bl), cycle=np.array(cyc), pose=np.array(poses),
                   vel=np.array(vels), phase=np.array(phases), is_over=np.array([m.is_over(t) for t in times]))
        for k, v in rec.items():
            out["%s/%s" % (name, k)] = np.asarray(v)
    # pose3d helpers on random unit quaternions
    q = rng.normal(size=(64, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    pts = rng.normal(size=(64, 3))
    out["pose3d/q"] = q; out["pose3d/pts"] = pts
    out["pose3d/heading"] = np.array([pose3d.calc_heading(x) for x in q])
    out["pose3d/rotate"] = np.array([pose3d.QuaternionRotatePoint(p, x) for p, x in zip(pts, q)])
    aa = [pose3d.QuaternionToAxisAngle(x) for x in q]
    out["pose3d/axis"] = np.array([a[0] for a in aa]); out["pose3d/angle"] = np.array([a[1] for a in aa])
    ang = rng.uniform(-10, 10, size=64)
    out["pose3d/ang_in"] = ang
    out["pose3d/norm_angle"] = np.array([pose3d.normalize_rotation_angle(a) for a in ang])
    out["pose3d/map_pi"] = np.array(pose3d.MapToMinusPiToPi(list(ang)))
    out["pose3d/heading_rot"] = np.array([pose3d.calc_heading_rot(x) for x in q])
    # Butterworth action filter
    f = action_filter.ActionFilterButter(sampling_rate=1 / (0.001 * 33), num_joints=12)
    out["filter/b"] = f.b[0]; out["filter/a"] = f.a[0]
    x0 = np.array([0, 0.67, -1.25] * 4)
    f.init_history(x0)
    xs = x0 + rng.normal(scale=0.2, size=(20, 12))
    out["filter/x0"] = x0; out["filter/x"] = xs
    out["filter/y"] = np.array([f.filter(x) for x in xs])
    # PD motor model
    mm = minitaur_motor.MotorModel(kp=[220.0] * 12, kd=[0.3, 2.0, 2.0] * 4, torque_limits=None)
    qq, qd, cmd = rng.normal(size=(8, 12)), rng.normal(size=(8, 12)), rng.normal(size=(8, 12))
    out["motor/q"], out["motor/qd"], out["motor/cmd"] = qq, qd, cmd
    out["motor/tau"] = np.array([mm.convert_to_torque(c, a, b, b, minitaur_motor.POSITION)[0] for a, b, c in zip(qq, qd, cmd)])
    # the other control modes, strength ratios and torque limits of the same class (minitaur_motor.py:137-171)
    lim = rng.uniform(0.3, 2.0, size=12)
    sr = rng.uniform(0.8, 1.2, size=12)
    mm2 = minitaur_motor.MotorModel(kp=[220.0] * 12, kd=[0.3, 2.0, 2.0] * 4, torque_limits=lim)
    mm2.set_strength_ratios(sr)
    hyb = rng.normal(size=(8, 60))
    hyb[:, 1::5] = rng.uniform(50.0, 250.0, size=(8, 12)); hyb[:, 3::5] = rng.uniform(0.3, 3.0, size=(8, 12))
    out["motor/limits"], out["motor/strength"], out["motor/hybrid_cmd"] = lim, sr, hyb
    out["motor/tau_position_limited"] = np.array([mm2.convert_to_torque(c, a, b, b, minitaur_motor.POSITION)[0] for a, b, c in zip(qq, qd, cmd)])
    out["motor/tau_torque"] = np.array([mm2.convert_to_torque(c, a, b, b, minitaur_motor.TORQUE)[0] for a, b, c in zip(qq, qd, cmd)])
    out["motor/tau_hybrid"] = np.array([mm2.convert_to_torque(c, a, b, b, minitaur_motor.HYBRID)[0] for a, b, c in zip(qq, qd, hyb)])
    np.savez_compressed(os.path.join(GOLD, "ref_motion_kat.npz"), **out)
    print("ref_motion_kat.npz:", len(out), "arrays")


def randomizer_kat(ref):
    """Known answers from the reference's own ControllableEnvRandomizerFromConfig driven with a recording robot."""
    sys.path.insert(0, os.path.join(ref, "OpenRoboRL"))
    tf = types.ModuleType("tensorflow")
    tf.logging = types.SimpleNamespace(info=lambda *a, **k: None)
    sys.modules.setdefault("tensorflow", tf)
    from envs.utilities.randomizer import controllable_env_randomizer_from_config as crc

    class Robot(object):
        num_motors = 12

        def __init__(self):
            rng = np.random.RandomState(7)
            self.base_mass = list(rng.uniform(1, 14, size=5))          # chassis link ids: base + 4 hip links
            self.leg_mass = list(rng.uniform(0.05, 2, size=12))        # 8 leg links + 4 motor links
            self.base_inertia = [tuple(rng.uniform(1e-3, 0.3, size=3)) for _ in range(5)]
            self.leg_inertia = [tuple(rng.uniform(1e-5, 1e-2, size=3)) for _ in range(12)]
            self.calls = {}

        def get_base_mass_from_urdf(self): return self.base_mass
        def get_leg_mass_from_urdf(self): return self.leg_mass
        def get_base_inertia_from_urdf(self): return self.base_inertia
        def get_leg_inertia_from_urdf(self): return self.leg_inertia
        def get_num_knee_joints(self): return 8

        def __getattr__(self, name):
            if name.startswith("set_"):
                return lambda v, _n=name: self.calls.__setitem__(_n, np.array(v, dtype=np.float64))
            raise AttributeError(name)

    out = {}
    for seed in (0, 1, 12345):
        robot = Robot()
        r = crc.ControllableEnvRandomizerFromConfig(verbose=False, randomization_seed=seed)
        r.randomize_env(robot)
        out["seed%d/uniforms" % seed] = np.random.RandomState(seed).random_sample(28)
        for k, v in r.get_randomization_parameters().items():
            out["seed%d/param/%s" % (seed, k)] = np.atleast_1d(np.asarray(v, dtype=np.float64))
        for k, v in robot.calls.items():
            out["seed%d/call/%s" % (seed, k)] = v
        out["robot/base_mass"] = np.array(robot.base_mass); out["robot/leg_mass"] = np.array(robot.leg_mass)
        out["robot/base_inertia"] = np.array(robot.base_inertia); out["robot/leg_inertia"] = np.array(robot.leg_inertia)
    np.savez_compressed(os.path.join(GOLD, "ref_randomizer_kat.npz"), **out)
    print("ref_randomizer_kat.npz:", len(out), "arrays;", sorted(k for k in out if k.startswith("seed0/")))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--skip-trace", action="store_true")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    install_shims()
    motion_clips(args.reference)
    policies(args.reference)
    motion_kat(args.reference)
    randomizer_kat(args.reference)
    if not args.skip_trace:
        try:
            from scripts import make_trace
        except ImportError:
            import make_trace
        make_trace.main(args.reference, GOLD)
