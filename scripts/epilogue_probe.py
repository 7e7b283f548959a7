"""Kernel time as a function of the number of physics substeps per control step: t = t0 + AR * t_sub separates
the once-per-step part (staging, reward, reference update, termination, observation) from the substep loop."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openroborl_b200.env import BatchedImitationEnv
from tests import common
n = 4096
res = []
for ar in (3, 18, 33, 66):
    env = BatchedImitationEnv("laikago", common.product_clips(["laikago_pace"]), num_robot=n, seed=0, auto_reset=True, action_repeat=ar)
    env.reset()
    gen = torch.Generator(device=env.device); gen.manual_seed(0)
    acts = [torch.randn((n, 12), device=env.device, generator=gen) * 0.125 for _ in range(8)]
    for i in range(30):
        env.step(acts[i % 8])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 150
    e0.record()
    for i in range(K):
        env.step(acts[i % 8])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    res.append((ar, ms))
    print("AR=%2d  %.4f ms/step" % (ar, ms), flush=True)
    env.close()
(a0, t0), (a1, t1) = res[1], res[3]
b = (t1 - t0) / (a1 - a0)
print("per substep %.2f us, once-per-step part %.1f us" % (b * 1e3, (t0 - b * a0) * 1e3))
