"""Known answers of the reference's learner-side NumPy code (BUILD-CONTAINER TOOLING, like make_golden.py).

Executes, unmodified, from /root/reference:
* `add_vtarg_and_adv`  (OpenRoboRL/agents/ppo_imitation.py:68-93; pure NumPy inside a module that imports tensorflow, mpi4py
  and stable_baselines -- stubbed, none of them is touched by the function), and
* `MpiAdam.update / sync` (OpenRoboRL/stable_baselines/common/mpi_adam.py:8-70) on NumPy-backed "variables", with a two-rank
  communicator stub whose Allreduce adds a recorded second rank's gradient,
and writes tests/golden/ref_ppo_kat.npz.  oracle/ppo.py, the CUDA GAE kernel (orl_gae) and learner.MpiAdam are tested
against these vectors.

    python scripts/make_ppo_golden.py [--reference /root/reference]
"""
import argparse
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = Anything(self.__name__ + "." + name)
        setattr(self, name, m)
        return m

    def __call__(self, *a, **k):
        return None


class Var(object):
    """A NumPy-backed stand-in for a TF variable."""

    def __init__(self, value):
        self.value = np.array(value, dtype=np.float32)


class Comm(object):
    """Two MPI ranks seen from rank 0: Allreduce(SUM) adds the other rank's recorded gradient."""

    def __init__(self, other_grads):
        self.other, self.i = other_grads, 0

    def Allreduce(self, local, out, op=None):
        out[...] = local + (self.other[self.i] if self.other is not None else 0)
        self.i += 1

    def Get_size(self):
        return 2 if self.other is not None else 1

    def Get_rank(self):
        return 0

    def Bcast(self, buf, root=0):
        return None


def load(path, name, stubs):
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    R = os.path.join(args.reference, "OpenRoboRL")
    names = ["tensorflow", "mpi4py", "stable_baselines", "stable_baselines.common", "stable_baselines.common.tf_util",
             "stable_baselines.common.policies", "stable_baselines.common.mpi_adam", "stable_baselines.common.mpi_moments",
             "stable_baselines.common.misc_util", "stable_baselines.common.runners", "stable_baselines.ppo1", "agents",
             "agents.imitation_runners", "stable_baselines.logger"]
    stubs = {n: Anything(n) for n in names}
    for n, m in stubs.items():                      # `import a.b.c as x` walks attributes: link children into their parents
        if "." in n and n.rsplit(".", 1)[0] in stubs:
            setattr(stubs[n.rsplit(".", 1)[0]], n.rsplit(".", 1)[1], m)
    stubs["stable_baselines.ppo1"].pposgd_simple = types.SimpleNamespace(PPO1=object)
    tfu = stubs["stable_baselines.common.tf_util"]
    tfu.numel = lambda v: int(v.value.size)
    tfu.in_session = lambda f: f

    def get_flat(var_list, sess=None):
        return lambda: np.concatenate([v.value.reshape(-1) for v in var_list])

    def set_from_flat(var_list, sess=None):
        def assign(theta):
            off = 0
            for v in var_list:
                n = v.value.size
                v.value = np.asarray(theta[off:off + n], dtype=np.float32).reshape(v.value.shape)
                off += n
        return assign
    tfu.GetFlat, tfu.SetFromFlat = get_flat, set_from_flat
    ppo = load(os.path.join(R, "agents", "ppo_imitation.py"), "ref_ppo_imitation", stubs)
    adam_mod = load(os.path.join(R, "stable_baselines", "common", "mpi_adam.py"), "ref_mpi_adam", stubs)

    out = {}
    rng = np.random.default_rng(2024)
    cases = [(16, 4), (32, 13), (8, 1), (5, 64), (24, 2)]
    out["gae/cases"] = np.array(cases)
    for c, (T, N) in enumerate(cases):
        seg = {"rewards": rng.uniform(0, 1, T * N).astype(np.float32), "vpred": rng.normal(size=T * N).astype(np.float32),
               "nextvpreds": rng.normal(size=T * N).astype(np.float32), "episode_starts": rng.random(T * N) < 0.15}
        for k, v in seg.items():
            out["gae/%d/%s" % (c, k)] = v.copy()
        ppo.add_vtarg_and_adv(seg, N, 0.95, 0.95)
        out["gae/%d/adv" % c] = seg["adv"].copy()
        out["gae/%d/tdlamret" % c] = seg["tdlamret"].copy()
    # MpiAdam: 12 updates (crosses nothing special; check_synced runs at step 0), PPO's epsilon 1e-5, one and two ranks
    for tag, two in (("adam1", False), ("adam2", True)):
        shapes = [(3,), (2, 5), (7, 4)]
        var_list = [Var(rng.normal(size=s)) for s in shapes]
        n = sum(v.value.size for v in var_list)
        grads = rng.normal(size=(12, n)).astype(np.float32)
        other = rng.normal(size=(12, n)).astype(np.float32) if two else None
        lrs = np.array([1e-2 * (1.0 - 0.05 * i) for i in range(12)])
        opt = adam_mod.MpiAdam(var_list, epsilon=1e-5, comm=Comm(other))
        out[tag + "/theta0"] = np.concatenate([v.value.reshape(-1) for v in var_list])
        opt.sync()
        thetas = []
        for g, lr in zip(grads, lrs):
            opt.update(g, lr)
            thetas.append(np.concatenate([v.value.reshape(-1) for v in var_list]))
        out[tag + "/grads"] = grads
        if two:
            out[tag + "/other_grads"] = other
        out[tag + "/lrs"] = lrs
        out[tag + "/thetas"] = np.array(thetas)
        out[tag + "/exp_avg"] = opt.exp_avg.copy()
        out[tag + "/exp_avg_sq"] = opt.exp_avg_sq.copy()
        out[tag + "/shapes"] = np.array([int(np.prod(s)) for s in shapes])
    path = os.path.join(ROOT, "tests", "golden", "ref_ppo_kat.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "with", len(out), "arrays")


if __name__ == "__main__":
    main()
