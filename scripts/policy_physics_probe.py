"""How long does the reference's PRETRAINED policy (trained against Bullet and the real URDF) stay on its clip in this repo's
physics model, and which model parameters does that depend on?  An indirect look at the unpinned part of the path
(DESIGN.md section 5): 1024 robots, test-mode settings, deterministic mean actions, auto-reset, 400 control steps per setting.

    python scripts/policy_physics_probe.py [laikago_pace|laikago_trot|laikago_spin|minicheetah_trot]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import common                                   # noqa: E402


def run(robot, clip, phys, steps=400, N=1024, seed=0, ranges=None):
    from openroborl_b200.env import BatchedImitationEnv
    policy = common.load_policy(clip)
    env = BatchedImitationEnv(robot, common.product_clips([clip]), num_robot=N, seed=seed, auto_reset=True, phys=phys,
                              max_episode_steps=600, enable_randomizer=ranges is not None, randomizer_ranges=ranges)
    o = env.reset().cpu().numpy()
    env.stats(reset=True)
    for t in range(steps):
        a = np.clip(policy(o), -2 * np.pi, 2 * np.pi).astype(np.float32)
        o = env.step(torch.as_tensor(a, device=env.device))[0].cpu().numpy()
    st = env.stats(reset=True)                             # [sum reward, sum finished-episode length, finished episodes, robot steps]
    env.close()
    return st[0] / st[3], (st[1] / st[2] if st[2] else float("nan")), st[2]


if __name__ == "__main__":
    clip = sys.argv[1] if len(sys.argv) > 1 else "laikago_pace"
    robot = "mini_cheetah" if clip.startswith("mini") else "laikago"
    settings = [("default", {})]
    for k, vals in (("mu_plane", (0.5, 2.0)), ("contact_kn", (3e4, 3e5)), ("contact_bn", (100.0, 2000.0)), ("contact_bt", (300.0, 30000.0)),
                    ("contact_d0", (0.0005, 0.008)), ("limit_k", (100.0,)), ("gravity_z", (-9.0,))):
        for v in vals:
            settings.append(("%s=%g" % (k, v), {k: v}))
    if "--model" not in sys.argv:
        for name, phys in settings:
            r, L, n = run(robot, clip, phys)
            print("%-22s mean reward %.3f   mean finished-episode length %6.1f   episodes %d" % (name, r, L, n), flush=True)
    # robot-model parameters through the randomizer's setters with degenerate ranges (lo = hi)
    nominal = {"mass": (1.0, 1.0), "inertia": (1.0, 1.0), "motor strength": (1.0, 1.0), "latency": (0.002, 0.002),
               "lateral friction": (1.0, 1.0), "joint friction": (0.0, 0.0)}
    sweeps = [("nominal (randomizer path)", {})]
    for k, vals in (("mass", (0.5, 0.75, 1.5)), ("inertia", (0.25, 0.5, 2.0, 4.0)), ("motor strength", (0.6, 0.8, 1.25, 1.6)),
                    ("latency", (0.0, 0.01, 0.03)), ("joint friction", (0.05, 0.3))):
        for v in vals:
            sweeps.append(("%s=%g" % (k, v), {k: (v, v)}))
    for name, over in sweeps:
        r, L, n = run(robot, clip, {}, ranges=dict(nominal, **over))
        print("%-28s mean reward %.3f   mean finished-episode length %6.1f   episodes %d" % (name, r, L, n), flush=True)
