"""A/B timing of the host-buffer path (orl_step_host) of step-kernel builds on one GPU.

    python scripts/e2e_probe.py build/exp/lib_a.so build/exp/lib_b.so ...      (or "default" for the in-tree build)

Per build (fresh process, ORL_LIB): the bench.py C2 workload stepped through `step_host` with page-locked buffers,
host-side tracking policy between the calls (untimed); prints mean / median / p90 of the per-call wall time over
E2E_STEPS calls after 20 warm-up calls, three repeats.
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    import numpy as np
    sys.path.insert(0, ROOT)
    from tests import common
    N = int(os.environ.get("AB_N", "4096"))
    K = int(os.environ.get("E2E_STEPS", "300"))
    for rep in range(3):
        env = common.make_product("laikago", ["laikago_pace"], N, seed=rep, auto_reset=True)
        spec, mdl = env.model.spec, env.model
        perm = np.asarray([mdl.q_index[j] for j in mdl.motor_id_list])
        joff, jdir = np.asarray(spec.JOINT_OFFSETS, np.float32), np.asarray(spec.JOINT_DIRECTIONS, np.float32)
        init = np.asarray(spec.INIT_MOTOR_ANGLES, np.float32)
        rng = np.random.default_rng(7 + rep)
        noise = [(rng.normal(size=(N, 12)) * 0.125).astype(np.float32) for _ in range(16)]
        a_host, out = env.alloc_host_buffers()
        obs = env.reset().cpu().numpy()
        ts = []
        for i in range(K + 20):
            a = (obs[:, 84 + 7:84 + 19][:, perm] - joff) * jdir - init + noise[i % 16]
            a_host[...] = np.clip(a, -2 * np.pi, 2 * np.pi)
            t0 = time.perf_counter()
            obs = env.step_host(a_host, out)[0]
            ts.append(time.perf_counter() - t0)
        ts = np.asarray(ts[20:]) * 1e3
        print("  rep %d: mean %.4f median %.4f p90 %.4f ms  -> %.3e env-steps/s (mean)" %
              (rep, ts.mean(), np.median(ts), np.percentile(ts, 90), N / (ts.mean() * 1e-3)), flush=True)
        env.close()


if __name__ == "__main__":
    if os.environ.get("E2E_CHILD"):
        child()
    else:
        for lib in sys.argv[1:]:
            envv = dict(os.environ, E2E_CHILD="1")
            if lib != "default":
                envv["ORL_LIB"] = os.path.abspath(lib)
            print(lib, flush=True)
            subprocess.call([sys.executable, os.path.abspath(__file__)], env=envv)
