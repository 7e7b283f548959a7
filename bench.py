"""Headline benchmark: env-steps/s of the batched Laikago motion-imitation env (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one control step (33 physics substeps + task/observation update) of every robot of the
batch.  Workload (config.workload): BASELINE.json configs[1], "Laikago dog_pace imitation, 4096 batched
robots on 1xB200"; with N GPUs every rank steps its own 4096 robots (weak scaling, no data-path collective).

`value`  : robots x steps / device time, actions already resident in HBM, CUDA events on the launching stream,
           L2 flushed between timed steps.
`e2e`    : the same through the C-ABI host-buffer call (orl_step_host): pinned H2D of the actions, step, D2H of
           observations/rewards/dones, host synchronisation, wall clock.
`roofline`: FP32-pipe roofline of the fused step kernel (2.3e5 algorithmic FLOP per robot-step)
           against an FFMA peak measured in this run; the HBM side (2.7 KB per robot-step) is reported under "hbm".
`cpu_baseline` / `--impl reference`: the float64 CPU oracle (oracle/env.py) timed on this box's host cores --
           a restatement stand-in, NOT PyBullet (the reference's PyBullet path cannot be installed here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROBOTS_PER_GPU = 4096
FLOP_PER_ROBOT_STEP = 2.3e5
BYTES_PER_ROBOT_STEP = 2.7e3
WORKLOADS = {
    # a = (q_ref_motor(t + dt) - INIT) + N(0, 0.125^2), clipped to +-2 pi, seed 1234
    "tracking": "laikago_pace imitation, %d robots per B200, RSI p=0.9, auto-reset, actions = next reference motor angles - INIT + N(0, 0.125^2) (SURVEY 8d C2)" % ROBOTS_PER_GPU,
    "noise": "laikago_pace imitation, %d robots per B200, RSI p=0.9, auto-reset, actions = N(0, 0.125^2) around the default pose (robots fall within ~40 steps)" % ROBOTS_PER_GPU,
}
WORKLOAD = WORKLOADS["tracking"]


def dist_info():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def _nvml_sample(self):
        """Fast path: NVML through nvidia_ml_py (no process spawn per sample)."""
        import pynvml as nv
        if not getattr(self, "_h", None):
            nv.nvmlInit()
            self._h = nv.nvmlDeviceGetHandleByIndex(self.index)
        sm = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
        mx = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        flag = lambda bit: "Active" if (r & bit) else "Not Active"
        return [str(sm), str(mx), flag(nv.nvmlClocksThrottleReasonHwSlowdown), flag(nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                flag(nv.nvmlClocksThrottleReasonSwThermalSlowdown), flag(nv.nvmlClocksThrottleReasonSwPowerCap)]

    def run(self):
        use_nvml = True
        while not self.stop_flag:
            try:
                if use_nvml:
                    try:
                        self.samples.append(self._nvml_sample())
                        time.sleep(0.01)
                        continue
                    except Exception:
                        use_nvml = False
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


def oracle_chunk(args):
    """Worker of the CPU arm: steps `n` oracle robots `steps` times on the bench workload (SURVEY 8d C2 action source);
    returns (robot-steps, seconds)."""
    n, steps, seed = args
    from tests import common
    orc = common.make_oracle("laikago", ["laikago_pace"], n, seed=seed)
    obs = orc.reset()
    rng = np.random.default_rng(seed)
    sp = orc.spec

    def policy(o):
        q_ref = np.asarray(o)[:, 84 + 7:84 + 19]                   # joints of the next target frame, URDF order
        a = (q_ref[:, orc.motor_q] - sp.joint_offset) * sp.joint_dir - sp.init_motor + rng.normal(scale=0.125, size=(n, 12))
        return np.clip(a, -2 * np.pi, 2 * np.pi)

    obs = orc.step(policy(obs))[0]                                 # warm-up step (imports, caches)
    t0 = time.perf_counter()
    for _ in range(steps):
        obs = orc.step(policy(obs))[0]
    return n * steps, time.perf_counter() - t0


def cpu_oracle_throughput(cores, robots_per_core, steps):
    import multiprocessing as mp
    if cores == 1:
        n, dt = oracle_chunk((robots_per_core, steps, 0))
        return n / dt, dt
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(oracle_chunk, [(robots_per_core, steps, s) for s in range(cores)])
    wall = max(r[1] for r in res)
    return sum(r[0] for r in res) / wall, time.perf_counter() - t0


def run_reference(args):
    rank, _, world = dist_info()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    robots_per_core, steps = 32, max(1, min(args.steps, 4))
    # keep BLAS single-threaded inside each worker: one process per core
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    value, wall = cpu_oracle_throughput(cores, robots_per_core, steps)
    sample = "%d processes x %d robots x %d control steps of the float64 oracle (laikago_pace)" % (cores, robots_per_core, steps)
    line = {
        "impl": "reference", "metric": "env-steps/sec (Laikago imitation, whole box)", "value": value, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU restatement stand-in (oracle/), not PyBullet: pybullet/gym/tf/mpi4py are not installable here"},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from tests import common
    from openroborl_b200 import _abi
    import ctypes as C

    rank, local_rank, world = dist_info()
    os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    N = ROBOTS_PER_GPU
    from openroborl_b200.env import BatchedImitationEnv
    env = BatchedImitationEnv("laikago", common.product_clips(["laikago_pace"]), num_robot=N, device=local_rank, seed=0,
                              gid0=rank * N, auto_reset=True, max_episode_steps=600)
    env.reset()
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    noise = [(torch.randn((N, 12), device=dev, generator=gen) * 0.125).contiguous() for _ in range(16)]
    workload = WORKLOADS[args.actions]
    # motor m reads URDF joint coordinate perm[m]; motor angle = (joint angle - offset) * direction (minitaur.py:549-553)
    spec, mdl = env.model.spec, env.model
    perm = torch.as_tensor([mdl.q_index[j] for j in mdl.motor_id_list], device=dev)
    joff = torch.as_tensor(np.asarray(spec.JOINT_OFFSETS, np.float32), device=dev)
    jdir = torch.as_tensor(np.asarray(spec.JOINT_DIRECTIONS, np.float32), device=dev)
    init = torch.as_tensor(np.asarray(spec.INIT_MOTOR_ANGLES, np.float32), device=dev)

    def policy(obs, i):
        """The action source of the workload, computed on the device between the timed launches."""
        if args.actions == "noise":
            a = noise[i % 16]
        else:
            q_ref = obs[:, 84 + 7:84 + 19]                                  # joints of target frame t + 1 step (URDF order)
            a = (q_ref[:, perm] - joff) * jdir - init + noise[i % 16]
        return torch.clamp(a, -2 * np.pi, 2 * np.pi).contiguous()

    act_buf = [torch.empty((N, 12), device=dev) for _ in range(2)]
    obs_now = env.obs
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)    # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        act_buf[i % 2].copy_(policy(obs_now, i))
        obs_now, _, _ = env.step(act_buf[i % 2])
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    K = args.steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches0 = env.launch_count
    env.stats(reset=True)
    for i in range(K):
        act_buf[i % 2].copy_(policy(obs_now, i))     # untimed: the policy is not part of the env step
        flush.zero_()                       # L2 flush between timed iterations (outside the event pair)
        ev[i][0].record()
        obs_now, _, _ = env.step(act_buf[i % 2])
        ev[i][1].record()
    barrier()
    launches = env.launch_count - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    st = env.stats(reset=True)
    # ---- end to end through the host-buffer C-ABI call -------------------------------------------------
    # every step's actions start in page-locked host memory, observation / reward / done end there; the host-side
    # policy between two calls is not timed (the timed region is the sum of the orl_step_host calls)
    a_host, out = env.alloc_host_buffers()
    noise_h = [n.cpu().numpy() for n in noise]
    perm_h, joff_h, jdir_h, init_h = perm.cpu().numpy(), joff.cpu().numpy(), jdir.cpu().numpy(), init.cpu().numpy()

    def policy_host(obs, i):
        if args.actions == "noise":
            a = noise_h[i % 16]
        else:
            a = (obs[:, 84 + 7:84 + 19][:, perm_h] - joff_h) * jdir_h - init_h + noise_h[i % 16]
        return np.clip(a, -2 * np.pi, 2 * np.pi)

    obs_h = obs_now.cpu().numpy()
    for i in range(3):
        a_host[...] = policy_host(obs_h, i)
        obs_h = env.step_host(a_host, out)[0]
    barrier()
    Ke = min(K, 200)
    e2e_s = 0.0
    for i in range(Ke):
        a_host[...] = policy_host(obs_h, i)
        t0 = time.perf_counter()
        obs_h = env.step_host(a_host, out)[0]
        e2e_s += time.perf_counter() - t0
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = float(t[0]), float(t[1])
    if rank == 0:
        value = world * N * K / (total_ms * 1e-3)
        e2e = world * N * Ke / e2e_s
        lib = _abi.load_library()
        tf = C.c_double(0.0)
        lib.orl_fp32_peak(local_rank, 5, C.byref(tf))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, hbm_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        kern_s = np.median(step_ms) * 1e-3            # one launch = one control step of N robots
        traffic = None                                # dram bytes per launch from the committed `ncu --set full` capture
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["orl_step_kernel"]
            if int(tj["robots"]) == N:
                traffic = int(tj["dram_bytes_read"]) + int(tj["dram_bytes_write"])
        except Exception:
            pass
        ach_tf = FLOP_PER_ROBOT_STEP * N / kern_s / 1e12
        ach_gb = BYTES_PER_ROBOT_STEP * N / kern_s / 1e9
        line = {
            "metric": "env-steps/sec (Laikago imitation, whole box)", "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "robots_per_gpu": N, "substeps_per_step": 33, "l2": "flushed between timed steps (256 MiB memset)",
                       "actions": "computed on the device from the previous observation between the timed launches", "episodes_finished_in_timed_region": st[2],
                       "mean_reward": st[0] / max(st[3], 1.0)},
            "e2e": {"value": e2e, "unit": "env-steps/s", "h2d_bytes_per_step": N * 12 * 4, "d2h_bytes_per_step": N * (160 * 4 + 4 + 1),
                    "steps": Ke, "api": "orl_step_host (C ABI, page-locked host buffers)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp32", "achieved": ach_tf, "peak": tf.value, "unit": "TFLOP/s", "frac": ach_tf / tf.value if tf.value else None,
                         "traffic": traffic, "kernel": "orl_step_kernel", "flop_per_robot_step": FLOP_PER_ROBOT_STEP,
                         "peak_source": "FFMA micro-benchmark measured in this run (orl_fp32_peak)",
                         "hbm": {"achieved": ach_gb, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gb / hbm_peak, "peak_source": hbm_src,
                                 "bytes_per_robot_step": BYTES_PER_ROBOT_STEP}},
            "clocks": sampler.summary(),
        }
        if not args.no_cpu_baseline and world == 1:
            v, wall = cpu_oracle_throughput(1, 64, 3)
            line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": 1, "kind": "port",
                                    "sample": "64 robots x 3 control steps of the float64 oracle (laikago_pace), one process"}
        print(json.dumps(line), flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--actions", default="tracking", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
