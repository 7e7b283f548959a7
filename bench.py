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
`configs` : the other single-GPU configurations of BASELINE.json measured the same way (sub-results, per rank):
           C3 = 16384 mini-cheetahs per GPU, C5 = the 5-clip Laikago motion library with random reference start.
`comm`    : (N > 1) one iteration of config C4 -- 8192 Laikagos per GPU, horizon 32, fp32 policy inference in the
           loop, one flat policy broadcast and one 32-byte statistics all-reduce over NCCL -- with the collectives'
           share of the iteration.
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
LIBRARY = ["laikago_pace", "laikago_trot", "laikago_spin", "laikago_hopturn", "laikago_sidesteps"]   # BASELINE.json configs[4]
# sub-results: name -> (robot, clips, robots per GPU, BASELINE.json configuration)
SUB_CONFIGS = {
    "C3": ("mini_cheetah", ["minicheetah_trot"], 16384, "mini-cheetah imitation_learning_minicheetah, 16384 robots on 1xB200"),
    "C5": ("laikago", LIBRARY, 8192, "Laikago motion library (pace/trot/spin/hopturn/sidesteps), random reference start, 8192 robots per B200"),
}
URDF = {"laikago": os.path.join(ROOT, "openroborl_b200", "assets", "laikago", "laikago_toes_limits.urdf"),
        "mini_cheetah": os.path.join(ROOT, "openroborl_b200", "assets", "mini_cheetah", "mini_cheetah.urdf")}


def dist_info():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def pin_to_gpu_numa(index):
    """Run this rank on the CPUs NVML reports as local to its GPU: the end-to-end path moves 2.8 MB per step through
    host memory and every rank on NUMA node 0 was what capped the 8-GPU e2e line in round 1."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        words = nv.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


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
    from oracle.env import OracleEnv                        # the CPU arm is the one place the bench runs the oracle
    z = np.load(os.path.join(ROOT, "openroborl_b200", "assets", "motions", "clips.npz"))     # read directly: this arm loads none of the product
    fd, wrap, cp, cr = z["laikago_pace/meta"]
    clips = [dict(wrap=bool(wrap), frame_duration=float(fd), cycle_pos=bool(cp), cycle_rot=bool(cr), frames=z["laikago_pace/frames"].copy())]
    orc = OracleEnv("laikago", URDF["laikago"], clips, n, seed=seed)
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


def base_config():
    """The `config` object both arms print (the driver compares them)."""
    return {"workload": WORKLOAD, "robots_per_gpu": ROBOTS_PER_GPU, "substeps_per_step": 33}


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
        "config": base_config(),
        "run_detail": {"note": "CPU restatement stand-in (oracle/), not PyBullet: pybullet/gym/tf/mpi4py are not installable here; "
                               "bounded sample of the workload: " + sample},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line), flush=True)


class Workload(object):
    """One env of a configuration plus the action source of SURVEY.md 8d (computed on the device between launches)."""

    def __init__(self, torch, robot, clip_names, N, rank, local_rank, actions="tracking"):
        from openroborl_b200.env import BatchedImitationEnv
        from openroborl_b200.motion import bundled_clips
        self.torch, self.N, self.mode = torch, N, actions
        dev = torch.device("cuda", local_rank)
        self.env = BatchedImitationEnv(robot, bundled_clips(clip_names), num_robot=N, device=local_rank, seed=0, gid0=rank * N,
                                       auto_reset=True, max_episode_steps=600)
        self.obs = self.env.reset()
        gen = torch.Generator(device=dev)
        gen.manual_seed(1234 + rank)
        self.noise = [(torch.randn((N, 12), device=dev, generator=gen) * 0.125).contiguous() for _ in range(16)]
        # motor m reads URDF joint coordinate perm[m]; motor angle = (joint angle - offset) * direction (minitaur.py:549-553)
        spec, mdl = self.env.model.spec, self.env.model
        self.perm = torch.as_tensor([mdl.q_index[j] for j in mdl.motor_id_list], device=dev)
        self.joff = torch.as_tensor(np.asarray(spec.JOINT_OFFSETS, np.float32), device=dev)
        self.jdir = torch.as_tensor(np.asarray(spec.JOINT_DIRECTIONS, np.float32), device=dev)
        self.init = torch.as_tensor(np.asarray(spec.INIT_MOTOR_ANGLES, np.float32), device=dev)
        self.act_buf = [torch.empty((N, 12), device=dev) for _ in range(2)]

    def policy(self, obs, i):
        torch = self.torch
        if self.mode == "noise":
            a = self.noise[i % 16]
        else:
            q_ref = obs[:, 84 + 7:84 + 19]                                  # joints of target frame t + 1 step (URDF order)
            a = (q_ref[:, self.perm] - self.joff) * self.jdir - self.init + self.noise[i % 16]
        return torch.clamp(a, -2 * np.pi, 2 * np.pi).contiguous()

    def run(self, W, K, flush):
        """W untimed + K timed steps (one CUDA-event pair per step on the launching stream, L2 flushed in between)."""
        torch, env = self.torch, self.env
        for i in range(W):
            self.act_buf[i % 2].copy_(self.policy(self.obs, i))
            self.obs, _, _ = env.step(self.act_buf[i % 2])
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        launches0 = env.launch_count
        env.stats(reset=True)
        for i in range(K):
            self.act_buf[i % 2].copy_(self.policy(self.obs, i))     # untimed: the policy is not part of the env step
            flush.zero_()                                           # L2 flush between timed iterations (outside the event pair)
            ev[i][0].record()
            self.obs, _, _ = env.step(self.act_buf[i % 2])
            ev[i][1].record()
        torch.cuda.synchronize()
        step_ms = [a.elapsed_time(b) for a, b in ev]
        return step_ms, env.launch_count - launches0, env.stats(reset=True)


def comm_iteration(torch, dist, rank, local_rank, world, iters=4):
    """BASELINE.json config C4 on this box: 8192 Laikagos per GPU, horizon 32, fp32 policy inference in the loop (random
    weights of the reference architecture 160-512-256-{12,1}), per iteration ONE flat policy broadcast (MpiAdam.sync) and
    ONE 32-byte all-reduce of the rollout statistics; the segment is one CUDA graph.  Returns the timing of one
    iteration and the collectives' share (CUDA events, max over ranks)."""
    from openroborl_b200.env import BatchedImitationEnv
    from openroborl_b200.motion import bundled_clips
    from openroborl_b200 import rollout
    N, T = 8192, 32
    env = BatchedImitationEnv("laikago", bundled_clips(["laikago_pace"]), num_robot=N, device=local_rank, seed=0, gid0=rank * N,
                              auto_reset=True, max_episode_steps=600)
    rng = np.random.default_rng(0)
    shapes = {"fc0": (160, 512), "fc1": (512, 256)}
    params = {}
    for head, out in (("pi", 12), ("vf", 1)):
        for l, (a, b) in shapes.items():
            params["%s_%s/w" % (head, l)] = rng.normal(scale=(2.0 / a) ** 0.5, size=(a, b)).astype(np.float32)
            params["%s_%s/b" % (head, l)] = np.zeros(b, np.float32)
        params[head + "/w"] = rng.normal(scale=0.01, size=(256, out)).astype(np.float32)
        params[head + "/b"] = np.zeros(out, np.float32)
    pol = rollout.MlpGaussianPolicy(params, env.device, tf32=False)
    col = rollout.RolloutCollector(env, pol, T, reset_all_on_any_done=False, seed=rank, use_cuda_graph=True)
    E = lambda: torch.cuda.Event(enable_timing=True)

    def iteration(ev=None):
        if ev: ev[0].record()
        pol.broadcast(src=0)
        if ev: ev[1].record()
        col.collect()
        if ev: ev[2].record()
        st = col.reduced_stats()
        if ev: ev[3].record()
        return st

    for _ in range(2):
        iteration()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    evs = [[E() for _ in range(4)] for _ in range(iters)]
    for e in evs:
        iteration(e)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    tot = sum(e[0].elapsed_time(e[3]) for e in evs) / iters
    coll = sum(e[0].elapsed_time(e[1]) + e[2].elapsed_time(e[3]) for e in evs) / iters
    t = torch.tensor([tot, coll], dtype=torch.float64, device=env.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    env.close()
    tot, coll = float(t[0]), float(t[1])
    n_param = sum(int(np.prod(v.shape)) for v in params.values())
    return {"config": "C4: Laikago imitation, %d robots sharded across %dxB200, PPO rollout collection end to end" % (world * N, world),
            "robots_per_gpu": N, "horizon": T, "policy_dtype": "f32 (TF32 off)", "cuda_graph": True,
            "value": world * N * T / (tot * 1e-3), "unit": "env-steps/s", "ms_per_iteration": tot,
            "collectives": "1 flat broadcast of %d f32 + 1 all-reduce of 4 f64 per iteration (NCCL)" % n_param,
            "collective_ms_per_iteration": coll, "collective_share": coll / tot if tot else None}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from openroborl_b200 import _abi
    import ctypes as C

    rank, local_rank, world = dist_info()
    # NCCL's INFO banner goes to stdout; WARN keeps rank 0's single JSON line clean -- unless the caller asked for more
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    pinned_cpus = pin_to_gpu_numa(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    N = ROBOTS_PER_GPU
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)    # > 126 MB L2
    # Device warm-up on a throw-away env (clocks, lazy module loading, instruction caches): the measured env below still
    # gets exactly W warm-up + K timed steps from its own fresh reset, i.e. the same window of the workload as before.
    scratch = Workload(torch, "laikago", ["laikago_pace"], N, rank, local_rank, actions=args.actions)
    scratch.run(40, 0, flush)
    scratch.env.close()
    wl = Workload(torch, "laikago", ["laikago_pace"], N, rank, local_rank, actions=args.actions)
    env = wl.env
    workload = WORKLOADS[args.actions]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W, K = max(args.warmup, 3), args.steps
    wl.run(W, 0, flush)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    step_ms, launches, st = wl.run(0, K, flush)
    barrier()
    total_ms = float(sum(step_ms))
    # ---- end to end through the host-buffer C-ABI call -------------------------------------------------
    # every step's actions start in page-locked host memory, observation / reward / done end there; the host-side
    # policy between two calls is not timed (the timed region is the sum of the orl_step_host calls)
    a_host, out = env.alloc_host_buffers()
    noise_h = [n.cpu().numpy() for n in wl.noise]
    perm_h, joff_h, jdir_h, init_h = wl.perm.cpu().numpy(), wl.joff.cpu().numpy(), wl.jdir.cpu().numpy(), wl.init.cpu().numpy()

    def policy_host(obs, i):
        if args.actions == "noise":
            a = noise_h[i % 16]
        else:
            a = (obs[:, 84 + 7:84 + 19][:, perm_h] - joff_h) * jdir_h - init_h + noise_h[i % 16]
        return np.clip(a, -2 * np.pi, 2 * np.pi)

    obs_h = wl.obs.cpu().numpy()
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
    # ---- the other single-GPU configurations (sub-results; bounded: 5 + min(K, 20) steps each) ---------------
    subs = {}
    if not args.no_sub_configs:
        for name, (robot, clips, n_sub, label) in SUB_CONFIGS.items():
            w = Workload(torch, robot, clips, n_sub, rank, local_rank, actions="tracking")
            w.run(5, 0, flush)
            barrier()
            ms, _, st_sub = w.run(0, min(K, 20), flush)
            barrier()
            subs[name] = (label, n_sub, float(sum(ms)), len(ms), float(np.median(ms)), st_sub)
            w.env.close()
    comm = None
    if world > 1 and not args.no_comm:
        comm = comm_iteration(torch, dist, rank, local_rank, world)
    vals = [total_ms, e2e_s] + [v[2] for v in subs.values()]
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
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
            "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(base_config(), workload=workload),
            "run_detail": {"l2": "flushed between timed steps (256 MiB memset)", "reference_dtype": "f64 (the kernel computes in f32, FP64 only for the motion clock)",
                           "actions": "computed on the device from the previous observation between the timed launches",
                           "episodes_finished_in_timed_region": st[2], "mean_reward": st[0] / max(st[3], 1.0),
                           "cpus_pinned_to_gpu_numa_node": pinned_cpus},
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
        if subs:
            line["configs"] = {}
            for j, (name, (label, n_sub, _, k_sub, med_ms, st_sub)) in enumerate(subs.items()):
                ms_total = float(t[2 + j])
                v = world * n_sub * k_sub / (ms_total * 1e-3)
                a_tf = FLOP_PER_ROBOT_STEP * n_sub / (med_ms * 1e-3) / 1e12
                line["configs"][name] = {"config": label, "robots_per_gpu": n_sub, "value": v, "unit": "env-steps/s", "steps": k_sub, "warmup": 5,
                                         "ms_per_step": ms_total / k_sub, "roofline_frac": a_tf / tf.value if tf.value else None,
                                         "l2": "flushed between timed steps", "mean_reward": st_sub[0] / max(st_sub[3], 1.0)}
        if comm:
            line["comm"] = comm
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
    ap.add_argument("--no-sub-configs", action="store_true", help="skip the C3 / C5 sub-results")
    ap.add_argument("--no-comm", action="store_true", help="skip the C4 iteration (N > 1)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
