"""Step time with the domain randomizer (incl. randomised control latency: LAT kernel variant) vs without."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import common
for n in (4096, 32768):
    for rand in (False, True):
        env = common.make_product("laikago", ["laikago_pace"], n, seed=0, auto_reset=True, randomizer=rand)
        env.reset()
        gen = torch.Generator(device=env.device); gen.manual_seed(0)
        acts = [torch.randn((n, 12), device=env.device, generator=gen) * 0.125 for _ in range(8)]
        for i in range(30):
            env.step(acts[i % 8])
        env.stats(reset=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 150
        e0.record()
        for i in range(K):
            env.step(acts[i % 8])
        e1.record()
        torch.cuda.synchronize()
        st = env.stats(reset=True)
        ms = e0.elapsed_time(e1) / K
        print("N=%6d randomizer=%-5s %.4f ms/step  %.3e env-steps/s  mean episode %.1f" % (n, rand, ms, n / ms * 1e3, st[1] / max(st[2], 1)), flush=True)
        env.close()
