"""End-to-end PPO training on the CUDA environment (reference: `python3 OpenRoboRL/run.py --mode train`, run.py:186-232).

    python scripts/train.py --iters 40                                                      # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/train.py

Each rank steps its shard of the robots (train mode: domain randomizer on), collects `horizon` steps with batched
policy inference (CUDA-graph replay), runs the PPO1 update with gradient averaging over NCCL (MpiAdam) and rank 0
prints one JSON record per iteration and finally saves a stable-baselines zip."""
import argparse, json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import common
from openroborl_b200.env import BatchedImitationEnv
from openroborl_b200 import rollout, learner


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--robots-per-gpu", type=int, default=4096)
    ap.add_argument("--horizon", type=int, default=32)
    ap.add_argument("--batch", type=int, default=16384)
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--init", default="random", choices=["random", "pretrained"])
    ap.add_argument("--no-randomizer", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    os.environ["NCCL_DEBUG"] = "WARN"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    N = args.robots_per_gpu
    env = BatchedImitationEnv("laikago", common.product_clips(["laikago_pace"]), num_robot=N, device=local, seed=0, gid0=rank * N,
                              auto_reset=True, max_episode_steps=600, enable_randomizer=not args.no_randomizer)
    if args.init == "pretrained":
        z = np.load(os.path.join(ROOT, "tests", "golden", "policy_laikago_pace.npz"))
        params = {k.rsplit("_", 1)[0] + "/" + k.rsplit("_", 1)[1]: z[k] for k in z.files if k.rsplit("_", 1)[-1] in ("w", "b")}
    else:                                        # small random actor (outputs ~ 0 -> the default pose), zero critic head
        rng = np.random.default_rng(0)
        shp = {"pi_fc0": (160, 512), "pi_fc1": (512, 256), "pi": (256, 12), "vf_fc0": (160, 512), "vf_fc1": (512, 256), "vf": (256, 1)}
        params = {}
        for k, (a, b) in shp.items():
            s = 0.01 if k == "pi" else (1.0 if k == "vf" else np.sqrt(2.0))
            params[k + "/w"] = (rng.normal(size=(a, b)) * s / np.sqrt(a)).astype(np.float32)
            params[k + "/b"] = np.zeros(b, np.float32)
    pol = rollout.MlpGaussianPolicy(params, env.device, tf32=True)
    lrn = learner.PPOLearner(pol, optim_epochs=args.epochs, optim_stepsize=args.lr, optim_batchsize=args.batch, seed=rank)
    t0 = time.time()

    def log(rec):
        if rank == 0:
            rec = dict(rec, wall_s=round(time.time() - t0, 2))
            print(json.dumps(rec), flush=True)

    total = args.iters * world * N * args.horizon
    hist = lrn.learn(env, total_timesteps=total, horizon=args.horizon, use_cuda_graph=True, seed=rank, log=log)
    torch.cuda.synchronize()
    if rank == 0:
        wall = time.time() - t0
        print(json.dumps({"summary": True, "n_gpus": world, "iterations": len(hist), "timesteps": lrn.num_timesteps, "wall_s": wall,
                          "env_steps_per_s_incl_learning": lrn.num_timesteps / wall,
                          "mean_reward_first": hist[0]["mean_reward"], "mean_reward_last": hist[-1]["mean_reward"],
                          "ep_len_first": hist[0]["ep_len_mean"], "ep_len_last": hist[-1]["ep_len_mean"]}), flush=True)
        if args.out:
            lrn.save(args.out)
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
