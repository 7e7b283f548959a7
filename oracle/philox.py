"""Philox4x32-10 counter-based generator (Salmon et al., SC'11), NumPy uint32/uint64.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference draws its reset randomness from the
unseeded global `np.random` (imitation_task.py:1125-1141), so RNG-stream parity with it is
impossible by construction.  The B200 env instead keys a counter-based
generator by (seed, global robot id, episode index, draw index) so results do not depend on how
robots are sharded; this file is the CPU statement of that contract.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32(counter, key):
    """counter (..., 4) uint32, key (..., 2) uint32 -> (..., 4) uint32."""
    c = [np.asarray(counter[..., i], dtype=np.uint64) for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint64)
    k1 = np.asarray(key[..., 1], dtype=np.uint64)
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [(hi1 ^ c[1] ^ k0) & MASK, lo1, (hi0 ^ c[3] ^ k1) & MASK, lo0]
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return np.stack([x.astype(np.uint32) for x in c], axis=-1)


def u01_53(hi, lo):
    """53-bit uniform in [0,1): ((hi << 21) | (lo >> 11)) * 2^-53."""
    v = (np.asarray(hi, dtype=np.uint64) << np.uint64(21)) | (np.asarray(lo, dtype=np.uint64) >> np.uint64(11))
    return v.astype(np.float64) * (1.0 / 9007199254740992.0)


def reset_draws(seed, gid, episode, n_clips):
    """The three draws of one episode reset (order of imitation_task.py:183, 701, 1107-1121).

    Returns u_rsi in [0,1), clip id in [0, n_clips), u_time in [0,1).
    """
    gid = np.asarray(gid, dtype=np.uint32)
    episode = np.asarray(episode, dtype=np.uint32)
    key = np.stack([np.full(gid.shape, seed & 0xFFFFFFFF, dtype=np.uint32),
                    np.full(gid.shape, (seed >> 32) & 0xFFFFFFFF, dtype=np.uint32)], axis=-1)
    z = np.zeros_like(gid)
    a = philox4x32(np.stack([gid, episode, z, z], axis=-1), key)
    b = philox4x32(np.stack([gid, episode, z + np.uint32(1), z], axis=-1), key)
    u_rsi = u01_53(a[..., 0], a[..., 1])
    clip = ((a[..., 2].astype(np.uint64) * np.uint64(n_clips)) >> np.uint64(32)).astype(np.int64)
    u_time = u01_53(b[..., 0], b[..., 1])
    return u_rsi, clip, u_time


def randomizer_draws(seed, gid, episode, n=28):
    """Uniforms of the domain randomizer of one reset: Philox blocks 2.. of the (robot, episode) stream, one 32-bit
    word per draw, u = (word + 0.5) / 2^32 (the product's keying; the reference uses a RandomState per robot)."""
    gid = np.asarray(gid, dtype=np.uint32)
    episode = np.asarray(episode, dtype=np.uint32)
    key = np.stack([np.full(gid.shape, seed & 0xFFFFFFFF, dtype=np.uint32),
                    np.full(gid.shape, (seed >> 32) & 0xFFFFFFFF, dtype=np.uint32)], axis=-1)
    z = np.zeros_like(gid)
    words = [philox4x32(np.stack([gid, episode, z + np.uint32(2 + b), z], axis=-1), key) for b in range((n + 3) // 4)]
    w = np.concatenate(words, axis=-1)[..., :n]
    return (w.astype(np.float64) + 0.5) * (1.0 / 4294967296.0)


# ---- task options (imitation_task.py: clip switching :1057-1101, initial-state perturbation :193-195, 1199-1243,
# target-observation noise :274-275): further blocks of the same (robot, episode) stream -----------------------------
OPT_BLOCK = 9            # [0,1] perturb decision, [2,3] clip duration drawn at reset
NORMAL_BLOCK0 = 10       # normal i of the perturbation: block NORMAL_BLOCK0 + i // 2, words 2 (i % 2), 2 (i % 2) + 1
N_PERTURB_NORMALS = 32   # root xy (2), rotation angle (1), joint pose (12), root velocity xy (2), angular velocity (3), joint velocity (12)
AXIS_BLOCK = 26          # words 0..2: rotation axis ~ U(-1, 1)^3
CHANGE_BLOCK0 = 32       # clip change c of the episode: block 32 + 2 c ([0] new clip, [2,3] clip duration), 33 + 2 c ([0,1] time offset)
NOISE_BLOCK0 = 0x40000000  # + env step counter: words 0, 1 -> the heading noise of that observation


def _key(seed, gid):
    return np.stack([np.full(gid.shape, seed & 0xFFFFFFFF, dtype=np.uint32),
                     np.full(gid.shape, (seed >> 32) & 0xFFFFFFFF, dtype=np.uint32)], axis=-1)


def block(seed, gid, episode, ctr):
    gid = np.asarray(gid, dtype=np.uint32)
    episode = np.broadcast_to(np.asarray(episode, dtype=np.uint32), gid.shape)
    c = np.broadcast_to(np.asarray(ctr, dtype=np.uint32), gid.shape)
    return philox4x32(np.stack([gid, episode, c, np.zeros_like(gid)], axis=-1), _key(seed, gid))


def u01_32(w):
    return (np.asarray(w, dtype=np.float64) + 0.5) * (1.0 / 4294967296.0)


def normal(w0, w1):
    """Box-Muller: sqrt(-2 ln u1) cos(2 pi u2) with u = (word + 0.5) / 2^32."""
    return np.sqrt(-2.0 * np.log(u01_32(w0))) * np.cos(2.0 * np.pi * u01_32(w1))


def option_reset_draws(seed, gid, episode):
    """perturb decision u, clip-duration u, 32 standard normals, 3 axis components in (-1, 1)."""
    b = block(seed, gid, episode, OPT_BLOCK)
    u_perturb, u_dur = u01_53(b[..., 0], b[..., 1]), u01_53(b[..., 2], b[..., 3])
    zs = []
    for i in range(N_PERTURB_NORMALS):
        nb = block(seed, gid, episode, NORMAL_BLOCK0 + i // 2)
        zs.append(normal(nb[..., 2 * (i % 2)], nb[..., 2 * (i % 2) + 1]))
    ab = block(seed, gid, episode, AXIS_BLOCK)
    axis = -1.0 + 2.0 * u01_32(ab[..., 0:3])
    return u_perturb, u_dur, np.stack(zs, axis=-1), axis


def clip_change_draws(seed, gid, episode, change, n_clips):
    """new clip id, clip-duration u, time-offset u of clip change number `change` of the episode."""
    change = np.asarray(change, dtype=np.int64)
    a = block(seed, gid, episode, (CHANGE_BLOCK0 + 2 * change).astype(np.uint32))
    b = block(seed, gid, episode, (CHANGE_BLOCK0 + 1 + 2 * change).astype(np.uint32))
    clip = ((a[..., 0].astype(np.uint64) * np.uint64(n_clips)) >> np.uint64(32)).astype(np.int64)
    return clip, u01_53(a[..., 2], a[..., 3]), u01_53(b[..., 0], b[..., 1])


def tar_noise_draw(seed, gid, episode, env_step):
    b = block(seed, gid, episode, (NOISE_BLOCK0 + np.asarray(env_step, dtype=np.int64)).astype(np.uint32))
    return normal(b[..., 0], b[..., 1])
