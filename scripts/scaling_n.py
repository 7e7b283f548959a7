"""Throughput of the fused step kernel vs batch size on one GPU (no L2 flush; 200 steps)."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import common
for n in [int(x) for x in (sys.argv[1:] or [1024, 2048, 4096, 8192, 16384, 32768, 65536])]:
    env = common.make_product("laikago", ["laikago_pace"], n, seed=0, auto_reset=True)
    env.reset()
    gen = torch.Generator(device=env.device); gen.manual_seed(0)
    acts = [torch.randn((n, 12), device=env.device, generator=gen) * 0.125 for _ in range(8)]
    for i in range(20):
        env.step(acts[i % 8])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 100
    e0.record()
    for i in range(K):
        env.step(acts[i % 8])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print("N=%6d  %.3f ms/step  %.3e env-steps/s  warps/SMSP=%.2f" % (n, ms, n / ms * 1e3, n / 8 / 592.0), flush=True)
    env.close()
