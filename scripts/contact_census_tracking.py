import sys, os, torch, numpy as np
ROOT = "/root/repo"; sys.path.insert(0, ROOT)
from tests import common
n = 4096
env = common.make_product("laikago", ["laikago_pace"], n, seed=0, auto_reset=True)
orc = common.make_oracle("laikago", ["laikago_pace"], n)
names = ["base"] + orc.tree.link_names
dev = env.device
spec, mdl = env.model.spec, env.model
perm = torch.as_tensor([mdl.q_index[j] for j in mdl.motor_id_list], device=dev)
joff = torch.as_tensor(np.asarray(spec.JOINT_OFFSETS, np.float32), device=dev)
jdir = torch.as_tensor(np.asarray(spec.JOINT_DIRECTIONS, np.float32), device=dev)
init = torch.as_tensor(np.asarray(spec.INIT_MOTOR_ANGLES, np.float32), device=dev)
gen = torch.Generator(device=dev); gen.manual_seed(1234)
noise = [(torch.randn((n, 12), device=dev, generator=gen) * 0.125) for _ in range(16)]
def census(tag):
    st = env.get_state()
    k = orc.tree.fk(st["pos"].astype(np.float64), st["quat"].astype(np.float64), st["q"].astype(np.float64))
    act = np.zeros((n, len(orc.cp_body)), bool)
    for c in range(len(orc.cp_body)):
        b = orc.cp_body[c]
        P = k["x"][:, b] + np.einsum("nij,j->ni", k["R"][:, b], orc.cp_local[c])
        act[:, c] = -(P[:, 2] - orc.cp_r[c]) > 0
    print("== %s" % tag)
    for c in range(len(orc.cp_body)):
        if act[:, c].any():
            print("  cp %2d %-22s foot=%d r=%.3f local=%s  robots %5d" % (c, names[orc.cp_body[c]], int(orc.cp_foot[c]) if hasattr(orc, "cp_foot") else -1, orc.cp_r[c], np.round(orc.cp_local[c], 3), act[:, c].sum()))
    # legs whose lower-leg (knee-end) point touches: is the toe of the same leg down as well?
    for knee, toe in ((11, 13), (17, 19), (23, 25), (29, 31)):
        if act[:, knee].any():
            print("  leg %s: knee-end point down on %d robots, of which toe down too: %d" % (names[orc.cp_body[knee]], act[:, knee].sum(), (act[:, knee] & act[:, toe]).sum()))
    return act
obs = env.reset()
acc = None
for i in range(80):
    a = (obs[:, 84 + 7:84 + 19][:, perm] - joff) * jdir - init + noise[i % 16]
    obs, r, d = env.step(torch.clamp(a, -6.28, 6.28).contiguous())
    if i in (20, 40, 60, 79):
        act = census("tracking workload, after step %d (done this step: %d)" % (i + 1, int(d.sum())))
