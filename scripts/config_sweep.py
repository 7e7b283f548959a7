"""Throughput of the step kernel on the BASELINE.json configurations that fit one GPU (no L2 flush, 100 steps)."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import common
CONFIGS = [("laikago", ["laikago_pace"], 4096), ("mini_cheetah", ["minicheetah_trot"], 16384),
           ("laikago", ["laikago_pace"], 8192),
           ("laikago", ["laikago_pace", "laikago_trot", "laikago_spin", "laikago_hopturn", "laikago_sidesteps"], 8192)]
for robot, clips, n in CONFIGS:
    env = common.make_product(robot, clips, n, seed=0, auto_reset=True)
    env.reset()
    gen = torch.Generator(device=env.device); gen.manual_seed(0)
    acts = [torch.randn((n, 12), device=env.device, generator=gen) * 0.125 for _ in range(8)]
    for i in range(30):
        env.step(acts[i % 8])
    env.stats(reset=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 100
    e0.record()
    for i in range(K):
        env.step(acts[i % 8])
    e1.record()
    torch.cuda.synchronize()
    st = env.stats(reset=True)
    ms = e0.elapsed_time(e1) / K
    print("%-12s clips=%d N=%6d  %.3f ms/step  %.3e env-steps/s  mean episode %.1f steps, mean reward %.3f" %
          (robot, len(clips), n, ms, n / ms * 1e3, st[1] / max(st[2], 1), st[0] / max(st[3], 1)), flush=True)
    env.close()
