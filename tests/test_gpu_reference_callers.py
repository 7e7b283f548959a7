"""The reference's OWN callers of the env, unmodified, driving the drop-in (VERDICT r1 "missing" 2):
`traj_segment_generator` (R/agents/imitation_runners.py:22-208) and the `test` loop (R/run.py:151-183) run against
`openroborl_b200.compat.build_env` with stub `gym` / `stable_baselines` / `mpi4py` / `tensorflow` modules and a NumPy
policy.  The sources are read from /root/reference (build container) or from baseline/_ref/ (staged by
scripts/stage_reference_callers.py so that they travel to the GPU box); the test is skipped when neither exists.
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

from tests import common

pytestmark = pytest.mark.gpu


def _reference_dir():
    for d in ("/root/reference/OpenRoboRL", os.path.join(common.ROOT, "baseline", "_ref", "OpenRoboRL")):
        if os.path.exists(os.path.join(d, "agents", "imitation_runners.py")) and os.path.exists(os.path.join(d, "run.py")):
            return d
    return None


REF = _reference_dir()
needs_reference = pytest.mark.skipif(REF is None, reason="reference callers not staged (python scripts/stage_reference_callers.py)")


class _Anything(types.ModuleType):
    """A module whose every attribute exists (the callers import far more than the two functions use)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Anything(self.__name__ + "." + name)
        setattr(self, name, m)
        return m

    def __call__(self, *a, **k):
        return None


class _Comm(object):
    def allreduce(self, x, op=None):
        return x

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1


class _Stubs(object):
    """sys.modules / sys.path patch for the duration of an import of reference code."""
    NAMES = ["agents", "gym", "gym.spaces", "stable_baselines", "stable_baselines.common", "stable_baselines.common.vec_env",
             "stable_baselines.common.callbacks", "mpi4py", "tensorflow", "agents.imitation_policies", "agents.ppo_imitation",
             "envs", "envs.quadruped_robot", "envs.quadruped_robot.robots", "envs.quadruped_robot.task"]

    def __init__(self, box_type):
        self.box_type = box_type

    def __enter__(self):
        self.saved = {k: sys.modules.get(k) for k in self.NAMES}
        for n in self.NAMES:
            sys.modules[n] = _Anything(n)
        sys.modules["gym"].spaces = sys.modules["gym.spaces"]
        sys.modules["gym.spaces"].Box = self.box_type                       # isinstance(env.action_space, gym.spaces.Box)
        sys.modules["stable_baselines.common.vec_env"].VecEnv = type("VecEnv", (), {})
        sys.modules["mpi4py"].MPI = types.SimpleNamespace(COMM_WORLD=_Comm(), SUM="sum")
        return self

    @staticmethod
    def load(rel):
        """Execute one reference source file as a module (by path: an unrelated `agents` package may be installed)."""
        spec = importlib.util.spec_from_file_location("orl_reference_" + os.path.basename(rel)[:-3], os.path.join(REF, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    def __exit__(self, *a):
        for k, v in self.saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


class NumpyPolicy(object):
    """The reference's actor-critic (fixed sigma 0.125) in NumPy with the policy.step / agent.predict signatures."""
    initial_state = None

    def __init__(self, seed):
        z = np.load(os.path.join(common.GOLD, "policy_laikago_pace.npz"))
        self.p = {k: z[k] for k in z.files}
        self.rng = np.random.default_rng(seed)

    def _mlp(self, obs, head):
        h = np.maximum(obs.astype(np.float32) @ self.p[head + "_fc0_w"] + self.p[head + "_fc0_b"], 0)
        h = np.maximum(h @ self.p[head + "_fc1_w"] + self.p[head + "_fc1_b"], 0)
        return h @ self.p[head + "_w"] + self.p[head + "_b"]

    def step(self, obs, state=None, mask=None):
        mean = self._mlp(obs, "pi")
        act = mean + 0.125 * self.rng.standard_normal(mean.shape).astype(np.float32)
        return act, self._mlp(obs, "vf")[:, 0], None, None

    def predict(self, obs, deterministic=True):
        return np.clip(self._mlp(np.asarray(obs), "pi"), -2 * np.pi, 2 * np.pi).astype(np.float64), None


class _Callback(object):
    def on_rollout_start(self): pass
    def on_rollout_end(self): pass
    def on_step(self): return True


@pytest.fixture()
def motion_file(tmp_path):
    clip = common.product_clips(["laikago_pace"])[0]
    p = tmp_path / "laikago_pace.txt"
    p.write_text(clip.to_json())
    return str(p)


def _hand_loop(env, policy, num_robot, steps, steps_per_segment):
    """What traj_segment_generator does with the env, written independently: per-robot policy.step on a [1, 160] row,
    clip, one env.step with the list of actions, reset of everything when any robot is done.  (At a segment boundary the
    generator queries the policy a second time for the bootstrap values, imitation_runners.py:117-118: same draws here.)"""
    rows = []
    obs = env.reset()
    for t in range(steps):
        acts, vps = [], []
        for j in range(num_robot):
            a, v, _, _ = policy.step(np.asarray(obs[j]).reshape(1, -1))
            acts.append(a); vps.append(float(v[0]))
        if t > 0 and t % steps_per_segment == 0:
            for j in range(num_robot):
                policy.step(np.asarray(obs[j]).reshape(1, -1))
        clipped = [np.clip(a, env.action_space.low, env.action_space.high)[0] for a in acts]
        o_before = [np.asarray(o).reshape(-1).copy() for o in obs]
        obs, rew, done, info = env.step(clipped)
        rows.append((o_before, [a[0].copy() for a in acts], vps, list(rew), list(done)))
        if np.sum(done) > 0:
            obs = env.reset()
    return rows


@needs_reference
def test_reference_traj_segment_generator_drives_the_drop_in(motion_file):
    from openroborl_b200 import compat
    R, H = 4, 32                                  # 8 control steps per segment
    mk = lambda: compat.build_env("imitation_learning_laikago", num_robot=R, num_parallel_envs=1, mode="train",
                                  enable_randomizer=False, motion_files=[motion_file], seed=11)
    env = mk()
    with _Stubs(type(env.action_space)):
        runners = _Stubs.load("agents/imitation_runners.py")
        gen = runners.traj_segment_generator(NumpyPolicy(5), env, H, callback=_Callback())
        segs = []
        for _ in range(4):                        # 32 control steps: crosses the 20-step time limit of the curriculum start
            s = next(gen)
            segs.append({k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else (list(v) if isinstance(v, list) else v)) for k, v in s.items()})
    for s in segs:
        assert s["observations"].shape == (H, 160) and s["actions"].shape == (H, 12) and s["total_timestep"] == H
        for k in ("rewards", "vpred", "nextvpreds", "dones", "episode_starts", "true_rewards"):
            assert s[k].shape == (H,)
        assert np.isfinite(s["observations"]).all() and np.isfinite(s["rewards"]).all()
        assert np.array_equal(s["rewards"], s["true_rewards"])
    assert segs[0]["episode_starts"][:R].all() and not segs[0]["episode_starts"][R:2 * R].any()
    # the same env driven by an independent loop with the same policy stream: identical buffers (time-major interleaved)
    rows = _hand_loop(mk(), NumpyPolicy(5), R, 4 * H // R, H // R)
    flat = lambda k: np.concatenate([s[k] for s in segs])
    obs_all, act_all, rew_all, done_all, vp_all = flat("observations"), flat("actions"), flat("rewards"), flat("dones"), flat("vpred")
    for t, (o, a, v, r, d) in enumerate(rows):
        for j in range(R):
            i = t * R + j
            assert np.array_equal(obs_all[i], o[j]) and np.array_equal(act_all[i], a[j])
            assert vp_all[i] == np.float32(v[j]) and rew_all[i] == np.float32(r[j]) and bool(done_all[i]) == bool(d[j])
    # reset-on-any-done: a row with a done flag is followed by the stale default-pose histories for EVERY robot, its
    # bootstrap values are cut, and every done event closes one entry of ep_rets / ep_lens
    done_rows = [t for t, row in enumerate(rows) if any(row[4])]
    assert done_rows and done_rows[0] <= 20                               # the curriculum starts at 20-step episodes
    nv = flat("nextvpreds")
    for t in done_rows:
        d = np.array(rows[t][4], dtype=bool)
        assert np.array_equal(nv[t * R:(t + 1) * R][d], np.zeros(int(d.sum()), np.float32))       # cut where the episode ended
        if t + 1 < len(rows):
            # ... and only there: the robots that were reset along with it bootstrap from the value of the FRESH episode's
            # first observation (imitation_runners.py:134-136 overwrites the zero of :187-189) -- the reference's behaviour
            assert np.array_equal(nv[t * R:(t + 1) * R][~d], vp_all[(t + 1) * R:(t + 2) * R][~d])
            nxt = obs_all[(t + 1) * R:(t + 2) * R]
            assert np.array_equal(nxt[:, 0:48], np.zeros_like(nxt[:, 0:48]))      # IMU and last-action histories after reset
            assert flat("episode_starts")[(t + 1) * R:(t + 2) * R].tolist() == [bool(x) for x in rows[t][4]]
    assert sum(len(s["ep_rets"]) for s in segs) == len([t for t in done_rows if t < len(rows)])
    lens = [l for s in segs for l in s["ep_lens"]]
    assert all(l % R == 0 for l in lens) and lens[0] == (done_rows[0] + 1) * R
    first = sum(sum(rows[t][3]) for t in range(done_rows[0] + 1)) / R
    assert abs([x for s in segs for x in s["ep_rets"]][0] - first) <= 1e-4 * max(1.0, abs(first))


@needs_reference
def test_reference_test_loop_drives_the_drop_in(motion_file, capsys):
    from openroborl_b200 import compat
    R = 3
    robot = [compat.Minitaur(name_robot="laikago", robot_index=i, enable_randomizer=False) for i in range(R)]
    task = [compat.ImitationTask(ref_motion_filenames=[motion_file], enable_cycle_sync=True, tar_frame_steps=[1, 2, 10, 30],
                                 ref_state_init_prob=0.9, warmup_time=0.25) for _ in range(R)]
    mk = lambda: compat.WrapperEnv(compat.LocomotionGymEnv(robot, task, device=0, seed=6), episode_length_start=40, episode_length_end=40,
                                   curriculum_steps=30000000, num_parallel_envs=1)
    env = mk()
    with _Stubs(type(env.action_space)):
        run = _Stubs.load("run.py")
        run.test(NumpyPolicy(0), env, num_procs=1, num_episodes=2)
    out = capsys.readouterr().out
    assert "Episode Count: 2" in out
    mean_return = float(out.split("Mean Return: ")[1].split()[0])
    # the same two episodes by hand (all(done) only at the 40-step limit of the shared env_step_counter)
    env2, pol = mk(), NumpyPolicy(0)
    o, curr, total, eps = env2.reset(), 0.0, 0.0, 0
    while eps < 2:
        a, _ = pol.predict(np.array(o))
        o, r, d, _ = env2.step(a)
        curr += sum(r)
        if all(d):
            o = env2.reset(); eps += 1
            total += curr                         # run.py:170-173 never clears curr_return: episode 2 counts episode 1 again
    assert abs(mean_return - total / 2) <= 1e-9 * max(1.0, abs(total)) and 0.0 < mean_return < 2 * 2 * 40 * R
