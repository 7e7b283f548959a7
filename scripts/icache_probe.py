"""Step-kernel time on a workload without falls (zero residual actions, episodes cut at 15 steps): separates the
cost of the common path from the slowest-warp tail of fallen robots.  ORL_LIB selects the build."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import common
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for mode in ("nofall", "bench"):
    env = common.make_product("laikago", ["laikago_pace"], n, seed=0, auto_reset=True)
    if mode == "nofall":
        env.set_max_episode_steps(15)
    env.reset()
    gen = torch.Generator(device=env.device); gen.manual_seed(0)
    sc = 0.0 if mode == "nofall" else 0.125
    acts = [torch.randn((n, 12), device=env.device, generator=gen) * sc for _ in range(8)]
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
    print("%s %-7s N=%d  %.4f ms/step  episodes=%d mean_len=%.1f" % (os.environ.get("ORL_LIB", "default").split("/")[-1], mode, n,
          e0.elapsed_time(e1) / K, st[2], st[1] / max(st[2], 1)), flush=True)
    env.close()
