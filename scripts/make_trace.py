"""Golden traces from the reference's OWN env classes (BUILD-CONTAINER TOOLING, see make_golden.py).

The reference's unmodified Minitaur / ImitationTask / LocomotionGymEnv / WrapperEnv are constructed exactly
as run.py:build_env does (run.py itself imports tensorflow/mpi4py and cannot be imported here) on top of
scripts/fake_bullet.py, reset and stepped with actions from the reference's pretrained policy; every
observation, reward, done flag and the task internals are recorded to tests/golden/ref_task_trace.npz.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_reference_env(robot_name, motion_file, num_robot):
    from envs.quadruped_robot.robots import minitaur
    from envs.quadruped_robot.task import imitation_task
    from envs.quadruped_robot import quadruped_gym_env, wrapper_env
    robot = [minitaur.Minitaur(name_robot=robot_name, robot_index=i, enable_randomizer=False) for i in range(num_robot)]
    task = [imitation_task.ImitationTask(ref_motion_filenames=[motion_file], enable_cycle_sync=True, tar_frame_steps=[1, 2, 10, 30],
                                         ref_state_init_prob=0.9, warmup_time=0.25) for _ in range(num_robot)]
    env = quadruped_gym_env.LocomotionGymEnv(robot, task)
    return wrapper_env.WrapperEnv(env, episode_length_start=600, episode_length_end=600, curriculum_steps=30000000,
                                  num_parallel_envs=1)       # mode == "test" (run.py:69-76)


class DrawLog(object):
    """Records np.random.uniform / randint results made by ImitationTask.reset (imitation_task.py:1125-1133)."""

    def __init__(self):
        self.draws = []
        self._u, self._r = np.random.uniform, np.random.randint

    def __enter__(self):
        def uni(a, b, size=None):
            v = self._u(a, b, size=size)
            self.draws.append(("u", a, b, v))
            return v

        def rint(a, b, size=None):
            v = self._r(a, b, size=size)
            self.draws.append(("i", a, b, v))
            return v
        np.random.uniform, np.random.randint = uni, rint
        return self

    def __exit__(self, *a):
        np.random.uniform, np.random.randint = self._u, self._r


class TaskDrawLog(object):
    """Per-robot log of every draw an ImitationTask makes through its own helpers (_rand_uniform / _randint / _randn,
    imitation_task.py:1125-1141), in call order: what oracle/env.py replays for the task options."""

    def __init__(self, tasks):
        self.logs = [[] for _ in tasks]
        for log, t in zip(self.logs, tasks):
            def wrap(kind, fn, log=log):
                def f(*a, **kw):
                    v = fn(*a, **kw)
                    log.append((kind, np.array(v, dtype=np.float64)))
                    return v
                return f
            t._rand_uniform = wrap("u", t._rand_uniform)
            t._randint = wrap("i", t._randint)
            t._randn = wrap("n", t._randn)

    def flat(self):
        """-> kinds (R, D) uint8 codes, values (R, D, 12) padded, counts (R,)"""
        D = max(len(l) for l in self.logs)
        kinds = np.zeros((len(self.logs), D), dtype=np.uint8)
        vals = np.zeros((len(self.logs), D, 12))
        size = np.zeros((len(self.logs), D), dtype=np.int64)
        for r, l in enumerate(self.logs):
            for d, (k, v) in enumerate(l):
                kinds[r, d] = ord(k)
                v = np.atleast_1d(v)
                vals[r, d, :len(v)] = v
                size[r, d] = len(v) if np.ndim(l[d][1]) else 0
        return kinds, vals, size, np.array([len(l) for l in self.logs])


def trace_options(reference, out, prefix, num_robot=4, steps=40):
    """The task options run.py leaves off (clip switching, initial-state perturbation, target-observation noise) on the
    reference's own classes: 3-clip Laikago library, clip timers of 0.2-0.5 s, perturbation probability 0.6."""
    from envs.quadruped_robot.robots import minitaur
    from envs.quadruped_robot.task import imitation_task
    from envs.quadruped_robot import quadruped_gym_env, wrapper_env
    names = ["laikago_pace", "laikago_trot", "laikago_spin"]
    files = [os.path.join(reference, "OpenRoboRL/envs/quadruped_robot/task/motions", n + ".txt") for n in names]
    robot = [minitaur.Minitaur(name_robot="laikago", robot_index=i, enable_randomizer=False) for i in range(num_robot)]
    task = [imitation_task.ImitationTask(ref_motion_filenames=files, enable_cycle_sync=True, tar_frame_steps=[1, 2, 10, 30],
                                         ref_state_init_prob=0.9, warmup_time=0.25, clip_time_min=0.2, clip_time_max=0.5,
                                         perturb_init_state_prob=0.6, tar_obs_noise=[0.1, 0.2], weight=0.7) for _ in range(num_robot)]
    genv = quadruped_gym_env.LocomotionGymEnv(robot, task)
    env = wrapper_env.WrapperEnv(genv, episode_length_start=600, episode_length_end=600, curriculum_steps=30000000, num_parallel_envs=1)
    np.random.seed(11)
    log = TaskDrawLog(env.task)
    obs = env.reset()
    rng = np.random.RandomState(5)
    O, R, D, A, TI = [np.array(obs)], [], [], [], []
    for s in range(steps):
        a = np.clip(rng.normal(scale=0.15, size=(num_robot, 12)), -2 * np.pi, 2 * np.pi)
        A.append(a.copy())
        obs, r, done, info = env.step([x.copy() for x in a])
        O.append(np.array(obs)); R.append(np.array(r)); D.append(np.array(done))
        TI.append(np.array([[t._get_motion_time(), t._prev_motion_phase] + list(t._origin_offset_pos) + list(t._origin_offset_rot)
                            + [t._active_motion_id, t._clip_change_time, t._motion_time_offset] for t in env.task]))
    kinds, vals, size, counts = log.flat()
    out.update({prefix + "/draw_kind": kinds, prefix + "/draw_value": vals, prefix + "/draw_size": size, prefix + "/draw_count": counts,
                prefix + "/obs": np.array(O), prefix + "/reward": np.array(R), prefix + "/done": np.array(D),
                prefix + "/action": np.array(A), prefix + "/task": np.array(TI)})
    changes = int(np.sum(np.diff(np.array(TI)[..., 9], axis=0) != 0))
    print("%s: %d draws per robot %s, %d clip changes, rewards %.3f..%.3f" % (prefix, kinds.shape[1], counts, changes, np.min(R), np.max(R)))


def trace(reference, robot_name, clip_name, policy_name, num_robot, steps, want_warmup, out, prefix):
    from tests import common
    mfile = os.path.join(reference, "OpenRoboRL/envs/quadruped_robot/task/motions", clip_name + ".txt")
    env = build_reference_env(robot_name, mfile, num_robot)
    policy = common.load_policy(policy_name) if policy_name else None
    seed = 0
    while True:
        np.random.seed(seed)
        with DrawLog() as log:
            obs = env.reset()
        warm = [bool(t._curr_episode_warmup) for t in env.task]
        if (not want_warmup) or (any(warm) and not all(warm)):
            break
        seed += 1
    d = log.draws
    assert len(d) == 3 * num_robot, d
    u_rsi = np.array([d[3 * i][3] for i in range(num_robot)])
    clip = np.array([d[3 * i + 1][3] for i in range(num_robot)])
    toff = np.array([d[3 * i + 2][3] for i in range(num_robot)])
    rng = np.random.RandomState(99)
    O, R, D, A, TI = [np.array(obs)], [], [], [], []
    for s in range(steps):
        o = np.array(obs)
        a = policy(o).astype(np.float64) if policy is not None else rng.normal(scale=0.15, size=(num_robot, 12))
        a = a + rng.normal(scale=0.02, size=a.shape)
        a = np.clip(a, -2 * np.pi, 2 * np.pi)
        A.append(a.copy())
        obs, r, done, info = env.step([x.copy() for x in a])
        O.append(np.array(obs)); R.append(np.array(r)); D.append(np.array(done))
        TI.append(np.array([[t._get_motion_time(), t._prev_motion_phase] + list(t._origin_offset_pos) + list(t._origin_offset_rot)
                            for t in env.task]))
    out.update({prefix + "/u_rsi": u_rsi, prefix + "/clip": clip, prefix + "/time_offset": toff, prefix + "/warmup": np.array(warm),
                prefix + "/obs": np.array(O), prefix + "/reward": np.array(R), prefix + "/done": np.array(D),
                prefix + "/action": np.array(A), prefix + "/task": np.array(TI)})
    print("%s: seed %d, warmup %s, rewards %.3f..%.3f, done at end %s" % (prefix, seed, warm, np.min(R), np.max(R), D[-1]))


def main(reference, gold):
    from scripts import fake_bullet
    fake_bullet.install(reference)
    os.chdir(reference)        # the reference opens 'OpenRoboRL/config/pybullet_sim_param.yaml' relative to the cwd
    out = {}
    trace(reference, "laikago", "laikago_pace", "laikago_pace", 3, 22, True, out, "laikago_pace")
    trace(reference, "laikago", "laikago_spin", None, 2, 26, False, out, "laikago_spin")
    trace(reference, "mini_cheetah", "minicheetah_trot", "minicheetah_trot", 2, 18, False, out, "minicheetah_trot")
    np.savez_compressed(os.path.join(gold, "ref_task_trace.npz"), **out)
    print("ref_task_trace.npz:", len(out), "arrays")
    opt = {}
    trace_options(reference, opt, "laikago_options")
    np.savez_compressed(os.path.join(gold, "ref_task_options_trace.npz"), **opt)


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import make_golden
    make_golden.install_shims()
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference", os.path.join(ROOT, "tests", "golden"))
