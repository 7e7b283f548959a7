import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests import common
n = int(sys.argv[1]); steps = int(sys.argv[2])
env = common.make_product("laikago", ["laikago_pace"], n, seed=0, auto_reset=True)
env.reset()
gen = torch.Generator(device=env.device); gen.manual_seed(0)
acts = [torch.randn((n, 12), device=env.device, generator=gen) * 0.125 for _ in range(8)]
for i in range(steps):
    env.step(acts[i % 8])
torch.cuda.synchronize()
st = env.get_state()
import numpy as np
print("z mean %.3f min %.3f  |qd| mean %.2f  step_counter mean %.1f" % (st["pos"][:,2].mean(), st["pos"][:,2].min(), np.abs(st["qd"]).mean(), st["env_step_counter"].mean()), env.stats())
