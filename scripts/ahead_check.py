"""Reference packets computed ahead of the workers (DevEnv::ahead) against the same build with ORL_AHEAD=0 (every packet
computed inline by the worker): observations, rewards, done flags and the final state records must be bit-identical.

    python scripts/ahead_check.py            (parent: runs both children and compares their dumps)
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [("laikago", "laikago_pace", 4096, False, 60), ("laikago", "laikago_pace,laikago_trot,laikago_spin", 1000, False, 60),
         ("minicheetah", "minicheetah_trot", 33, True, 40), ("laikago", "laikago_pace", 16384, False, 12)]


def child(out):
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    from tests import common
    res = {}
    for ci, (robot, clips, N, rand, K) in enumerate(CASES):
        K = K * int(os.environ.get("AHEAD_SOAK", "1"))             # AHEAD_SOAK=25: 1500-step runs (rare interleavings)
        env = common.make_product(robot, clips.split(","), N, seed=5, auto_reset=True, randomizer=rand)
        obs = env.reset()
        gen = torch.Generator(device=env.device); gen.manual_seed(77)
        h = torch.zeros((), dtype=torch.float64, device=env.device)
        dones = 0
        for i in range(K):
            a = (torch.randn((N, 12), device=env.device, generator=gen) * 0.2).contiguous()
            obs, r, d = env.step(a)
            h = h * 1.000001 + obs.double().sum() + 3.0 * r.double().sum() + 7.0 * d.double().sum()
            dones += int(d.sum().item())
        torch.cuda.synchronize()
        res["obs%d" % ci] = obs.cpu().numpy(); res["state%d" % ci] = env.state.cpu().numpy().copy(); res["h%d" % ci] = np.float64(h.item())
        res["dones%d" % ci] = np.int64(dones)
        env.stats(reset=True)
        env.close()
    np.savez(out, **res)


if __name__ == "__main__":
    if os.environ.get("AHEAD_CHILD"):
        child(os.environ["AHEAD_CHILD"])
        sys.exit(0)
    import numpy as np
    tmp = tempfile.mkdtemp()
    outs = []
    for name, val in (("on", "1"), ("off", "0"), ("on2", "1")):
        out = os.path.join(tmp, name + ".npz")
        subprocess.run([sys.executable, os.path.abspath(__file__)], env=dict(os.environ, AHEAD_CHILD=out, ORL_AHEAD=val), check=True)
        outs.append(np.load(out))
    ok = True
    for k in outs[0].files:
        a, b = outs[0][k], outs[1][k]
        same = np.array_equal(a.view(np.uint8) if a.dtype.kind == "f" and a.ndim else a, b.view(np.uint8) if b.dtype.kind == "f" and b.ndim else b)
        ok &= bool(same)
        print("%-8s %s %s" % (k, "identical" if same else "DIFFERENT (max |d| %.3e)" % float(np.nanmax(np.abs(a.astype(np.float64) - b.astype(np.float64)))),
                              a if a.ndim == 0 else a.shape))
    for k in outs[0].files:          # which record fields differ (on / off, and between two runs of the same setting)
        if k.startswith("state"):
            for other, label in ((1, "on vs off"), (2, "on vs on")):
                a, b = outs[0][k].view(np.uint32), outs[other][k].view(np.uint32)
                cols = np.nonzero((a != b).any(axis=0))[0]
                print("%s %s: differing columns %s in %d rows" % (k, label, cols.tolist(), int((a != b).any(axis=1).sum())))
    print("AHEAD CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)
