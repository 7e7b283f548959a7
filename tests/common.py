"""Shared helpers of the test-suite: fixtures, oracle construction, state transfer product -> oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
ASSETS = os.path.join(ROOT, "openroborl_b200", "assets")
CLIPS = os.path.join(ASSETS, "motions", "clips.npz")

URDF = {"laikago": os.path.join(ASSETS, "laikago", "laikago_toes_limits.urdf"),
        "mini_cheetah": os.path.join(ASSETS, "mini_cheetah", "mini_cheetah.urdf")}


def golden_clips_raw(names):
    """Oracle-format clip dicts (raw frames) from the package's clip library (assets/motions/clips.npz)."""
    z = np.load(CLIPS)
    out = []
    for n in names:
        fd, wrap, cp, cr = z[n + "/meta"]
        out.append(dict(wrap=bool(wrap), frame_duration=float(fd), cycle_pos=bool(cp), cycle_rot=bool(cr),
                        frames=z[n + "/frames"].copy()))
    return out


def product_clips(names):
    from openroborl_b200.motion import load_clip_library
    return load_clip_library(CLIPS, names)


def load_policy(name):
    z = np.load(os.path.join(GOLD, "policy_%s.npz" % name))
    W = [z["pi_fc0_w"], z["pi_fc1_w"], z["pi_w"]]
    B = [z["pi_fc0_b"], z["pi_fc1_b"], z["pi_b"]]

    def policy(obs):
        h = np.maximum(obs.astype(np.float32) @ W[0] + B[0], 0)
        h = np.maximum(h @ W[1] + B[1], 0)
        return np.clip(h @ W[2] + B[2], -2 * np.pi, 2 * np.pi).astype(np.float32)
    return policy


def make_oracle(robot, clip_names, N, seed=0, gid0=0, max_episode_steps=600, task=None, phys=None, randomizer=False):
    from oracle.env import OracleEnv
    return OracleEnv(robot, URDF[robot], golden_clips_raw(clip_names), N, seed=seed, gid0=gid0,
                     max_episode_steps=max_episode_steps, task=task, phys=phys, randomizer=randomizer)


def make_product(robot, clip_names, N, seed=0, gid0=0, max_episode_steps=600, auto_reset=False, task=None, phys=None,
                 randomizer=False):
    from openroborl_b200.env import BatchedImitationEnv
    return BatchedImitationEnv(robot, product_clips(clip_names), num_robot=N, seed=seed, gid0=gid0,
                               max_episode_steps=max_episode_steps, auto_reset=auto_reset, task=task, phys=phys,
                               enable_randomizer=randomizer)


def oracle_phys_from_product(phys):
    """Map the product's physics parameter names to the oracle's."""
    return dict(gravity=(0.0, 0.0, phys["gravity_z"]), kn=phys["contact_kn"], bn=phys["contact_bn"], d0=phys["contact_d0"],
                bt=phys["contact_bt"], mu_plane=phys["mu_plane"], k_lim=phys["limit_k"], d_lim=phys["limit_d"],
                e0=phys["limit_e0"], margin_factor=phys["margin_factor"],
                max_coordinate_velocity=phys["max_coordinate_velocity"], joint_friction_k=phys["joint_friction_k"])


def load_rand_params(oracle, rp):
    """Product randomizer record (orl_state_layout.rand_params) -> the oracle's per-robot physical parameters."""
    ids, t = oracle.ids, oracle.tree
    N, nb = oracle.N, t.n_links + 1
    rm, ri = np.ones((N, nb)), np.ones((N, nb))
    for i in ids["chassis"]:
        rm[:, i + 1] = rp[:, 0]; ri[:, i + 1] = rp[:, 2]
    for i in list(ids["leg"]) + list(ids["motor"]):
        rm[:, i + 1] = rp[:, 1]; ri[:, i + 1] = rp[:, 3]
    oracle.link_mass = rm * np.asarray(t.mass, dtype=np.float64)[None, :]
    oracle.inertia_scale = ri
    oracle.foot_friction = np.where(rp[:, 5] >= 0, rp[:, 5] / oracle.phys["mu_plane"], -1.0)
    jf = np.zeros((N, 12))
    knees = [i for i in ids["foot"] if t.movable[i]]               # one per leg, URDF leg order
    for leg, link_id in enumerate(knees):
        jf[:, t.qidx[link_id]] = rp[:, 6 + leg]
    oracle.joint_friction = jf


def sync_oracle_from_product(oracle, st, rings=None):
    """Load a product state dict (BatchedImitationEnv.get_state) into an OracleEnv (float32 values widened)."""
    f = lambda a: np.asarray(a, dtype=np.float64)
    oracle.pos, oracle.quat, oracle.q = f(st["pos"]), f(st["quat"]), f(st["q"])
    oracle.lin, oracle.ang, oracle.qd = f(st["lin"]), f(st["ang"]), f(st["qd"])
    oracle.substep_counter = st["substep_counter"].astype(np.int64)
    oracle.step_counter = st["step_counter"].astype(np.int64)
    oracle.env_step_counter = st["env_step_counter"].astype(np.int64)
    oracle.episode = st["episode"].astype(np.int64)
    oracle.filt_x, oracle.filt_y = f(st["filt_x"]), f(st["filt_y"])
    oracle.prev_filtered = f(st["prev_filtered"])
    oracle.has_prev = oracle.step_counter > 0
    oracle.imu_hist, oracle.act_hist, oracle.ang_hist = f(st["imu_hist"]), f(st["act_hist"]), f(st["ang_hist"])
    oracle.last_action = f(st["act_hist"][:, 0])
    oracle.clip = st["clip"].astype(np.int64)
    oracle.time_offset = f(st["time_offset"])
    oracle.warmup = st["warmup"].astype(bool)
    oracle.origin_pos, oracle.origin_rot = f(st["origin_pos"]), f(st["origin_rot"])
    oracle.prev_phase = f(st["prev_phase"])
    oracle.ref_pose, oracle.ref_vel = f(st["ref_pose"]), f(st["ref_vel"])
    oracle.motor_strength = f(st["motor_strength"])
    if "clip_change_time" in st:
        oracle.clip_change_time = f(st["clip_change_time"])
        oracle.clip_changes = st["clip_changes"].astype(np.int64)
    if "rand_params" in st:
        load_rand_params(oracle, f(st["rand_params"]))
    # observation ring: entry 0 = current true state, entries 1/2 carry the delayed motor angles
    N = oracle.N
    from oracle.env import HIST
    oracle.hist = np.zeros((N, HIST, 43))
    if rings is None:
        oracle.hist_len = np.full(N, 4, dtype=np.int64)
        oracle.hist[:, 0] = oracle._true_obs()
        oracle.hist[:, 1, 0:12] = f(st["ring1"])
        oracle.hist[:, 2, 0:12] = f(st["ring2"])
        oracle.hist[:, 1, 36:40] = [0, 0, 0, 1]
        oracle.hist[:, 2, 36:40] = f(st["obs_quat"])
        oracle.hist[:, 3] = oracle.hist[:, 2]
        oracle.ctrl_obs = oracle.hist[:, 2].copy()
    else:
        # randomised control latency: rebuild the observation history from the product's rings
        # (entry E in slot E % 42, newest entry = substep counter + 1)
        ring_q, ring_b = rings
        R = ring_q.shape[1]
        oracle.ctrl_latency = f(st["rand_params"][:, 4])
        el = st["substep_counter"].astype(np.int64) + 1
        oracle.hist_len = np.minimum(el + 1, HIST)
        r = np.arange(N)
        for h in range(HIST):
            e = np.maximum(el - h, 0)
            oracle.hist[:, h, 0:12] = ring_q[r, e % R]
            oracle.hist[:, h, 36:40] = ring_b[r, e % R, 0:4]
            oracle.hist[:, h, 40:43] = ring_b[r, e % R, 4:7]
        oracle.hist[:, 0] = oracle._true_obs()            # the newest entry is the current true state (float64 of the same values)
        oracle.ctrl_obs = oracle._delay_obs(oracle.ctrl_latency)


def rel_err(a, b, floor=1e-12):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def random_states(rng, n, robot="laikago"):
    """Random but plausible poses/velocities: base near standing height, joints inside the limits."""
    quat = rng.normal(size=(n, 4)); quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    pos = np.concatenate([rng.uniform(-3, 3, size=(n, 2)), rng.uniform(0.6, 1.2, size=(n, 1))], axis=1)
    q = rng.uniform(-0.8, 0.8, size=(n, 12))
    pose = np.concatenate([pos, quat, q], axis=1).astype(np.float32)
    vel = np.concatenate([rng.normal(size=(n, 3)), rng.normal(size=(n, 3)) * 2, rng.normal(size=(n, 12)) * 3], axis=1).astype(np.float32)
    return pose, vel


def sync_product_from_oracle_reset(env, orc):
    """Load the state of a freshly reset OracleEnv into the CUDA env (replay of recorded reference draws)."""
    N = orc.N
    f32 = lambda a: np.asarray(a, dtype=np.float32)
    st = dict(pos=f32(orc.pos), quat=f32(orc.quat), q=f32(orc.q), lin=f32(orc.lin), ang=f32(orc.ang), qd=f32(orc.qd),
              filt_x=f32(orc.filt_x), filt_y=f32(orc.filt_y), prev_filtered=f32(orc.prev_filtered),
              ring1=f32(orc.hist[:, 1, 0:12]), ring2=f32(orc.hist[:, 1, 0:12]),
              imu_hist=f32(orc.imu_hist), act_hist=f32(orc.act_hist), ang_hist=f32(orc.ang_hist),
              ref_pose=f32(orc.ref_pose), ref_vel=f32(orc.ref_vel), origin_pos=f32(orc.origin_pos), origin_rot=f32(orc.origin_rot),
              time_offset=orc.time_offset.copy(), prev_phase=orc.prev_phase.copy(), clip=orc.clip.astype(np.int32),
              warmup=orc.warmup.astype(np.int32), substep_counter=np.zeros(N, np.int32), step_counter=np.zeros(N, np.int32),
              env_step_counter=np.zeros(N, np.int32), obs_quat=np.tile(np.array([0, 0, 0, 1], np.float32), (N, 1)))
    env.set_state(st)
