"""Run the same batch twice (and optionally with a second build via ORL_LIB2 in a subprocess) and compare states."""
import sys, os, subprocess
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import common
def run(n, steps):
    env = common.make_product("laikago", ["laikago_pace"], n, seed=0, auto_reset=True)
    env.reset()
    gen = torch.Generator(device=env.device); gen.manual_seed(0)
    acts = [torch.randn((n, 12), device=env.device, generator=gen) * 0.125 for _ in range(8)]
    outs = []
    for i in range(steps):
        o, r, d = env.step(acts[i % 8])
        if i in (0, 4, 19, steps - 1):
            outs.append((o.cpu().numpy().copy(), r.cpu().numpy().copy(), d.cpu().numpy().copy()))
    env.close()
    return outs
n, steps = int(sys.argv[1]), int(sys.argv[2])
if len(sys.argv) > 3:
    np.savez(sys.argv[3], *[x for t in run(n, steps) for x in t]); sys.exit(0)
a, b = run(n, steps), run(n, steps)
print("same build twice identical:", all(np.array_equal(x, y) for s, t in zip(a, b) for x, y in zip(s, t)))
if os.environ.get("ORL_LIB2"):
    env2 = dict(os.environ, ORL_LIB=os.environ["ORL_LIB2"])
    subprocess.check_call([sys.executable, __file__, str(n), str(steps), "/tmp/other.npz"], env=env2)
    z = np.load("/tmp/other.npz"); flat = [x for t in a for x in t]
    for i, x in enumerate(flat):
        y = z["arr_%d" % i]
        print(i, "max abs diff", float(np.max(np.abs(x.astype(np.float64) - y.astype(np.float64)))), "equal", np.array_equal(x, y))
