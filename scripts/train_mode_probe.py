"""One GPU, 4096 robots, train mode (domain randomizer + randomised control latency: the LAT kernel instantiation)."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import common
n = 4096
env = common.make_product("laikago", ["laikago_pace"], n, seed=0, auto_reset=True, randomizer=True)
env.reset()
gen = torch.Generator(device=env.device); gen.manual_seed(0)
acts = [torch.randn((n, 12), device=env.device, generator=gen) * 0.125 for _ in range(8)]
for i in range(60):
    env.step(acts[i % 8])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(100):
    env.step(acts[i % 8])
e1.record()
torch.cuda.synchronize()
print("train mode N=%d  %.4f ms/step" % (n, e0.elapsed_time(e1) / 100))
env.close()
