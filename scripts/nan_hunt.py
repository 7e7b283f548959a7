import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import common
clips = ["laikago_pace", "laikago_trot", "laikago_spin", "laikago_hopturn", "laikago_sidesteps"]
n = 8192
env = common.make_product("laikago", clips, n, seed=0, auto_reset=True)
obs = env.reset()
print("nan in reset obs:", torch.isnan(obs).any().item())
gen = torch.Generator(device=env.device); gen.manual_seed(0)
acts = [torch.randn((n, 12), device=env.device, generator=gen) * 0.125 for _ in range(8)]
for i in range(200):
    prev = env.get_state()
    o, r, d = env.step(acts[i % 8])
    bad = torch.isnan(r) | torch.isnan(o).any(dim=1)
    if bad.any():
        idx = torch.nonzero(bad)[:, 0].cpu().numpy()
        print("step", i, "bad robots", idx[:10], "count", len(idx))
        j = int(idx[0])
        st = env.get_state()
        for k in ("clip", "warmup", "substep_counter", "step_counter", "env_step_counter", "episode", "time_offset"):
            if k in prev: print(" prev", k, prev[k][j])
        for k in ("pos", "quat", "q", "lin", "ang", "qd", "ref_pose", "ref_vel", "origin_pos", "origin_rot"):
            if k in prev: print(" prev", k, np.round(prev[k][j], 4))
        print(" reward", r[j].item(), "done", d[j].item())
        print(" obs nan idx", torch.nonzero(torch.isnan(o[j]))[:, 0].cpu().numpy()[:20])
        for k in ("pos", "quat", "q", "ref_pose", "origin_pos", "origin_rot"):
            if k in st: print(" now", k, np.round(st[k][j], 4))
        break
else:
    print("no NaN in 200 steps")
print(sorted(prev.keys()))
