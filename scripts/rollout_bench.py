"""BASELINE.json config C4: PPO rollout collection end to end -- 8192 Laikagos per GPU, horizon T = 32
control steps per iteration with batched policy inference in the loop, per iteration one flat policy broadcast
(MpiAdam.sync) and one all-reduce of the rollout statistics over NCCL, GAE on the device.

    python scripts/rollout_bench.py [--iters 10]                         # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/rollout_bench.py

Prints one JSON line on rank 0: env-steps/s = robots x T x iterations / device time (max over ranks)."""
import argparse, json, os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import common
from openroborl_b200.env import BatchedImitationEnv
from openroborl_b200 import rollout, dist as odist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--robots-per-gpu", type=int, default=8192)
    ap.add_argument("--horizon", type=int, default=32)
    ap.add_argument("--randomizer", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--tf32", action="store_true", help="TF32 policy GEMMs (default: fp32 like the reference's TF graph)")
    args = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    N = args.robots_per_gpu
    env = BatchedImitationEnv("laikago", common.product_clips(["laikago_pace"]), num_robot=N, device=local, seed=0,
                              gid0=rank * N, auto_reset=True, max_episode_steps=600, enable_randomizer=args.randomizer)
    z = np.load(os.path.join(ROOT, "tests", "golden", "policy_laikago_pace.npz"))
    params = {k.rsplit("_", 1)[0] + "/" + k.rsplit("_", 1)[1]: z[k] for k in z.files if k.rsplit("_", 1)[-1] in ("w", "b")}
    pol = rollout.MlpGaussianPolicy(params, env.device, tf32=args.tf32)
    col = rollout.RolloutCollector(env, pol, args.horizon, reset_all_on_any_done=False, seed=rank, use_cuda_graph=not args.no_graph)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def iteration():
        pol.broadcast(src=0)                       # MpiAdam.sync
        seg = col.collect()                        # T steps: policy forward, env step, buffers; GAE at the end
        return col.reduced_stats(), seg

    for _ in range(2):
        iteration()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        stats, seg = iteration()
    e1.record()
    sync()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=env.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t[0])
        steps = world * N * args.horizon * args.iters
        print(json.dumps({"metric": "env-steps/sec, PPO rollout collection end to end (config C4)", "value": steps / (ms * 1e-3),
                          "unit": "env-steps/s", "n_gpus": world, "robots_per_gpu": N, "horizon": args.horizon, "iterations": args.iters,
                          "ms_per_iteration": ms / args.iters, "randomizer": bool(args.randomizer), "cuda_graph": not args.no_graph,
                          "policy": "160-512-256-12 actor + critic (reference zip laikago_pace), TF32 matmuls, sigma 0.125",
                          "collectives_per_iteration": "1 flat policy broadcast (434k f32 x2 nets) + 1 all-reduce of 4 doubles",
                          "last_iteration_stats": [float(x) for x in stats]}), flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
