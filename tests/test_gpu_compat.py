"""The list-based drop-in classes (openroborl_b200.compat) keep the reference's env contract."""
import numpy as np
import pytest
import torch

from tests import common

pytestmark = pytest.mark.gpu


@pytest.fixture()
def motion_file(tmp_path):
    clip = common.product_clips(["laikago_pace"])[0]
    p = tmp_path / "laikago_pace.txt"
    p.write_text(clip.to_json())
    return str(p)


def test_train_mode_randomizes_the_dynamics_at_every_reset(motion_file):
    """run.py:205-211: enable_randomizer is on in train mode; Minitaur.reset -> randomize_env (minitaur.py:273-276)."""
    from openroborl_b200 import compat
    env = compat.build_env("imitation_learning_laikago", num_robot=16, num_parallel_envs=1, mode="train",
                           enable_randomizer=True, motion_files=[motion_file], seed=4)
    env.reset()
    cu = env._gym_env._cuda
    r1 = cu.get_state()["rand_params"].copy()
    assert (r1[:, 0] >= 0.8).all() and (r1[:, 0] <= 1.2).all() and np.ptp(r1[:, 0]) > 0.05        # mass
    assert (r1[:, 2] >= 0.5).all() and (r1[:, 2] <= 1.5).all()                                    # inertia
    assert (r1[:, 5] >= 0.5).all() and (r1[:, 5] <= 1.25).all()                                   # lateral friction
    assert (r1[:, 6:10] >= 0).all() and (r1[:, 6:10] <= 0.05).all()                               # knee friction
    env.reset()
    r2 = cu.get_state()["rand_params"]
    assert not np.array_equal(r1, r2)                                                             # new draw per episode
    s = cu.get_state()["motor_strength"]
    assert (s >= 0.8).all() and (s <= 1.2).all() and np.ptp(s) > 0.1


def test_build_env_contract(motion_file):
    from openroborl_b200 import compat
    env = compat.build_env("imitation_learning_laikago", num_robot=3, num_parallel_envs=1, mode="test",
                           enable_randomizer=False, motion_files=[motion_file], seed=4)
    assert env.num_robot == 3 and env.observation_space.shape == (160,) and env.action_space.shape == (12,)
    assert np.allclose(env.action_space.high, 2 * np.pi)
    obs = env.reset()
    assert isinstance(obs, list) and len(obs) == 3 and obs[0].shape == (160,) and obs[0].dtype == np.float64
    obs[0] = obs[0].reshape(1, 160)                      # the runner reassigns list items (imitation_runners.py:89)
    acts = [np.zeros(12), np.full(12, 0.01), env.action_space.sample().astype(np.float64)]
    before = [a.copy() for a in acts]
    o, r, d, info = env.step(acts)
    init = np.array([0, 0.67, -1.25] * 4)
    for a, b in zip(acts, before):
        np.testing.assert_allclose(a, b + init)          # set_act mutates the caller's action (minitaur.py:281)
    assert isinstance(o, list) and isinstance(r, list) and isinstance(d, list) and isinstance(r[0], float) and isinstance(d[0], bool)
    assert info[0]["terminated"] is d and info[0].get("episode") is None   # aliasing quirk; runner queries 'episode'
    assert np.sum(d) >= 0 and (all(d) in (True, False))
    # attribute forwarding (wrapper_env.py:55-56) and the task API
    assert env.env_step_counter == 1 and len(env.task) == 3 and len(env.robot) == 3 and env.pybullet_client is None
    t = env.task[1]
    assert t.get_num_motions() == 1 and t.get_num_tar_frames() == 4 and t.is_motion_over() is False
    np.testing.assert_allclose(t.build_target_obs(), o[1][84:])
    low, high = t.get_target_obs_bounds()
    assert low.shape == (76,) and np.allclose(low[:7], [-2, -2, -2, -1, -1, -1, -1])
    assert 0.0 <= t.reward() <= 1.0 and isinstance(t.done(), bool)
    assert env.robot[2].step_counter == 1 and env.robot[0].get_foot_link_ids() == [2, 3, 6, 7, 10, 11, 14, 15]
    with pytest.raises(NotImplementedError):
        t.update()
    env.close()


def test_matches_tensor_env_and_episode_semantics(motion_file):
    from openroborl_b200 import compat
    n = 5
    env = compat.build_env("imitation_learning_laikago", n, 1, "train", False, [motion_file], seed=7)
    assert env._max_episode_steps == 20                                   # curriculum start (run.py:54, wrapper_env.py:151-159)
    ten = common.make_product("laikago", ["laikago_pace"], n, seed=7, max_episode_steps=20)
    o = env.reset()
    ten.reset_host()       # LocomotionGymEnv.__init__ already reset once (quadruped_gym_env.py:211): episode index 1
    np.testing.assert_array_equal(np.array(o, dtype=np.float32), ten.reset_host())
    rng = np.random.default_rng(0)
    count0 = env._total_step_count
    for step in range(22):
        a = rng.normal(scale=0.05, size=(n, 12))
        o, r, d, _ = env.step([x.copy() for x in a])
        ot, rt, dt = ten.step_host(a.astype(np.float32))
        np.testing.assert_array_equal(np.array(o, dtype=np.float32), ot)
        np.testing.assert_array_equal(np.array(r, dtype=np.float32), rt)
        assert d == [bool(x) for x in dt]
        if step == 19:
            assert all(d)                                               # time limit: env_step_counter >= 20
    assert env.env_step_counter == 22                                     # no auto reset: done robots keep stepping
    assert env._total_step_count > count0 and env._total_step_count % n == 0      # += num_robot on steps with a done
    env.reset()
    assert env.env_step_counter == 0
    env.close(); ten.close()


def test_minicheetah_task_name_works(tmp_path):
    """run.py:57 derives 'minicheetah', which the reference's Minitaur rejects; the drop-in accepts it."""
    from openroborl_b200 import compat
    clip = common.product_clips(["minicheetah_trot"])[0]
    p = tmp_path / "mc.txt"; p.write_text(clip.to_json())
    env = compat.build_env("imitation_learning_minicheetah", 2, 1, "test", False, [str(p)])
    o = env.reset()
    o, r, d, _ = env.step([np.zeros(12), np.zeros(12)])
    assert len(o) == 2 and np.isfinite(r).all()
    with pytest.raises(ValueError, match="wrong robot select"):
        compat.Minitaur("hexapod")
    with pytest.raises(ValueError, match="no input robot or task"):
        compat.LocomotionGymEnv(None, None)
    env.close()


def test_vecenv_facade_resets_inside_step_and_reports_terminal_observations():
    """VecEnv contract (individual resets inside step, terminal_observation / episode in infos)."""
    from openroborl_b200.vecenv import VecImitationEnv
    N = 24
    clips = common.product_clips(["laikago_pace"])
    venv = VecImitationEnv("laikago", clips, num_envs=N, seed=3, max_episode_steps=5)
    ref = common.make_product("laikago", ["laikago_pace"], N, seed=3, max_episode_steps=5)       # same seed, no resets inside step
    obs = venv.reset()
    assert obs.shape == (N, 160) and obs.dtype == np.float32 and np.array_equal(obs, ref.reset().cpu().numpy())
    assert venv.observation_space.shape == (160,) and venv.action_space.shape == (12,)
    rng = np.random.default_rng(0)
    ends = 0
    for t in range(12):
        act = rng.normal(scale=0.1, size=(N, 12)).astype(np.float32)
        o, r, d, infos = venv.step(act)
        assert o.shape == (N, 160) and r.shape == (N,) and d.dtype == bool and len(infos) == N
        if t < 5:                                    # until the first time limit both envs see the same episode
            o2, r2, d2 = ref.step(torch.as_tensor(act, device=ref.device))
            assert np.array_equal(r, r2.cpu().numpy()) and np.array_equal(d, d2.cpu().numpy().astype(bool))
            for i in range(N):
                want = infos[i]["terminal_observation"] if d[i] else o[i]
                assert np.array_equal(want, o2.cpu().numpy()[i])
        for i in np.nonzero(d)[0]:
            ends += 1
            assert infos[i]["episode"]["l"] <= 5 and np.isfinite(infos[i]["episode"]["r"])
            assert not np.array_equal(infos[i]["terminal_observation"], o[i])        # the returned row already belongs to the new episode
        assert all("terminal_observation" not in infos[i] for i in np.nonzero(~d)[0])
    assert ends >= 2 * N                              # the 5-step limit fired at steps 5 and 10 for every robot
    venv.close(); ref.close()


def test_sim_param_yaml_is_read_relative_to_the_cwd(motion_file, tmp_path, monkeypatch):
    """quadruped_gym_env.py:158-179: OpenRoboRL/config/pybullet_sim_param.yaml, key quadruped_robot, relative to the cwd."""
    from openroborl_b200 import compat
    cfg = tmp_path / "OpenRoboRL" / "config"
    cfg.mkdir(parents=True)
    monkeypatch.chdir(tmp_path)
    (cfg / "pybullet_sim_param.yaml").write_text(
        "quadruped_robot:\n  enable_rendering: false\n  sim_time_step_s: 0.001\n  enable_rendering_gui: false\n  num_sim_iter_step: 300\n")
    env = compat.build_env("imitation_learning_laikago", num_robot=2, num_parallel_envs=1, mode="test",
                           enable_randomizer=False, motion_files=[motion_file], seed=1)
    assert abs(env._gym_env._env_time_step - 0.033) < 1e-12
    env.close()
    (cfg / "pybullet_sim_param.yaml").write_text("quadruped_robot:\n  enable_rendering: false\n  sim_time_step_s: 0.002\n"
                                                 "  enable_rendering_gui: false\n  num_sim_iter_step: 300\n")
    with pytest.raises(ValueError):
        compat.build_env("imitation_learning_laikago", num_robot=2, num_parallel_envs=1, mode="test",
                         enable_randomizer=False, motion_files=[motion_file], seed=1)
    (cfg / "pybullet_sim_param.yaml").write_text("other_robot:\n  sim_time_step_s: 0.001\n")
    with pytest.raises(ValueError):
        compat.build_env("imitation_learning_laikago", num_robot=2, num_parallel_envs=1, mode="test",
                         enable_randomizer=False, motion_files=[motion_file], seed=1)
