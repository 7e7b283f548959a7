"""A/B timing of step-kernel builds on one GPU.

    python scripts/ab_probe.py build/exp/lib_a.so build/exp/lib_b.so ...      (or "default" for the in-tree build)

For every build (loaded through ORL_LIB in a fresh process): the bench.py C2 workload measured the way the driver runs
it (5 warm-up + 20 timed steps from a fresh reset, L2 flushed between steps, CUDA events per step; median over repeats),
the same for 100 steps after 40 warm-up steps, and the common-path time (no falls, no resets in the timed window).
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    from tests import common
    robot = os.environ.get("AB_ROBOT", "laikago")
    clips = os.environ.get("AB_CLIPS", "laikago_pace").split(",")
    N = int(os.environ.get("AB_N", "4096"))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def run(W, K, mode, seed):
        env = common.make_product(robot, clips, N, seed=seed, auto_reset=True)
        if mode == "nofall":
            env.set_max_episode_steps(1000000)
        obs = env.reset()
        dev = env.device
        gen = torch.Generator(device=dev); gen.manual_seed(1234 + seed)
        noise = [(torch.randn((N, 12), device=dev, generator=gen) * (0.0 if mode == "nofall" else 0.125)).contiguous() for _ in range(16)]
        spec, mdl = env.model.spec, env.model
        perm = torch.as_tensor([mdl.q_index[j] for j in mdl.motor_id_list], device=dev)
        joff = torch.as_tensor(np.asarray(spec.JOINT_OFFSETS, np.float32), device=dev)
        jdir = torch.as_tensor(np.asarray(spec.JOINT_DIRECTIONS, np.float32), device=dev)
        init = torch.as_tensor(np.asarray(spec.INIT_MOTOR_ANGLES, np.float32), device=dev)

        def policy(o, i):
            if mode == "noise":
                return noise[i % 16]
            a = (o[:, 84 + 7:84 + 19][:, perm] - joff) * jdir - init + noise[i % 16]
            return torch.clamp(a, -2 * np.pi, 2 * np.pi).contiguous()
        buf = [torch.empty((N, 12), device=dev) for _ in range(2)]
        for i in range(W):
            buf[i % 2].copy_(policy(obs, i)); obs, _, _ = env.step(buf[i % 2])
        torch.cuda.synchronize()
        env.stats(reset=True)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for i in range(K):
            buf[i % 2].copy_(policy(obs, i))
            flush.zero_()
            ev[i][0].record(); obs, _, _ = env.step(buf[i % 2]); ev[i][1].record()
        torch.cuda.synchronize()
        ms = [a.elapsed_time(b) for a, b in ev]
        st = env.stats(reset=True)
        env.close()
        return float(np.mean(ms)), float(np.median(ms)), st[2]

    name = os.environ.get("ORL_LIB", "default").split("/")[-1]
    R = int(os.environ.get("AB_REPEATS", "5"))
    a = [run(5, 20, "tracking", s) for s in range(R)]
    b = [run(40, 100, "tracking", 7)]
    c = [run(3, 12, "nofall", 3) for _ in range(3)]       # RSI start, zero residual actions: nobody falls within 15 steps
    print("%-22s N=%d  driver-window(5+20) mean %.4f ms [%s]  episodes %.0f | steady(40+100) mean %.4f median %.4f | nofall mean %.4f" % (
        name, N, np.median([x[0] for x in a]), " ".join("%.4f" % x[0] for x in a), np.mean([x[2] for x in a]), b[0][0], b[0][1],
        np.median([x[0] for x in c])), flush=True)


if __name__ == "__main__":
    if os.environ.get("AB_CHILD"):
        child()
    else:
        for lib in sys.argv[1:]:
            env = dict(os.environ, AB_CHILD="1")
            if lib != "default":
                env["ORL_LIB"] = os.path.join(ROOT, lib) if not os.path.isabs(lib) else lib
            subprocess.run([sys.executable, os.path.abspath(__file__)], env=env)
