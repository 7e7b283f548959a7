import sys, os, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests import common
n = 4096
env = common.make_product("laikago", ["laikago_pace"], n, seed=0, auto_reset=True)
orc = common.make_oracle("laikago", ["laikago_pace"], n)
env.reset()
gen = torch.Generator(device=env.device); gen.manual_seed(0)
acts = [torch.randn((n, 12), device=env.device, generator=gen) * 0.125 for _ in range(8)]
for i in range(120):
    env.step(acts[i % 8])
st = env.get_state()
k = orc.tree.fk(st["pos"].astype(np.float64), st["quat"].astype(np.float64), st["q"].astype(np.float64))
act = np.zeros((n, len(orc.cp_body)), bool)
for c in range(len(orc.cp_body)):
    b = orc.cp_body[c]
    P = k["x"][:, b] + np.einsum("nij,j->ni", k["R"][:, b], orc.cp_local[c])
    act[:, c] = P[:, 2] - orc.cp_r[c] < 0
names = ["base"] + orc.tree.link_names
print("points per robot active: mean %.2f" % act.sum(1).mean())
for b in sorted(set(orc.cp_body)):
    sel = orc.cp_body == b
    print("%-16s frac robots with an active point %.3f  (points %d, radius>0: %d)" % (names[b], act[:, sel].any(1).mean(), sel.sum(), (orc.cp_r[sel] > 0).sum()))
