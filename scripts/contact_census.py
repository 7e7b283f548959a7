"""Which collision points are below the ground (a) right after reset (RSI poses) and (b) in steady state."""
import sys, os, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests import common
n = 4096
env = common.make_product("laikago", ["laikago_pace"], n, seed=0, auto_reset=True)
orc = common.make_oracle("laikago", ["laikago_pace"], n)
names = ["base"] + orc.tree.link_names
def census(tag):
    st = env.get_state()
    k = orc.tree.fk(st["pos"].astype(np.float64), st["quat"].astype(np.float64), st["q"].astype(np.float64))
    act = np.zeros((n, len(orc.cp_body)), bool); depth = np.zeros((n, len(orc.cp_body)))
    for c in range(len(orc.cp_body)):
        b = orc.cp_body[c]
        P = k["x"][:, b] + np.einsum("nij,j->ni", k["R"][:, b], orc.cp_local[c])
        depth[:, c] = -(P[:, 2] - orc.cp_r[c]); act[:, c] = depth[:, c] > 0
    print("== %s: points per robot active: mean %.2f" % (tag, act.sum(1).mean()))
    for c in range(len(orc.cp_body)):
        if act[:, c].any():
            print("  cp %2d %-16s r=%.3f local=%s  frac %.4f  max depth %.4f" % (c, names[orc.cp_body[c]], orc.cp_r[c], np.round(orc.cp_local[c], 3), act[:, c].mean(), depth[:, c].max()))
    return st
env.reset()
census("after reset")
gen = torch.Generator(device=env.device); gen.manual_seed(0)
acts = [torch.randn((n, 12), device=env.device, generator=gen) * 0.125 for _ in range(8)]
o, r, d = env.step(acts[0])
print("done after first step: %.4f" % d.float().mean().item())
for i in range(120):
    env.step(acts[i % 8])
st = census("steady state")
print("env_step_counter histogram (first 6):", np.bincount(st["env_step_counter"].astype(int))[:6])
