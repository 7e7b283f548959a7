"""Multi-GPU plumbing: robots shard across ranks, collectives only for statistics and policy weights.

The reference is data-parallel over MPI ranks, each rank owning its own world:
`MpiAdam.sync` broadcasts the flat policy parameters from rank 0 (stable_baselines/common/mpi_adam.py:69)
and episode statistics are gathered/summed (agents/ppo_imitation.py:408-421, run.py:174-175).  Here rank r
owns robots [r*N/G, (r+1)*N/G) with RNG streams keyed by the GLOBAL robot id, there is no exchange inside
reset()/step(), and per iteration there are exactly two small collectives.  Works with the `nccl` backend
on GPUs (NVLink/NVSwitch) and with `gloo` on CPU (tests).
"""
import torch
import torch.distributed as dist


def shard_range(num_robot_total, rank, world_size):
    """Contiguous shard of global robot ids owned by `rank` (remainder spread over the first ranks)."""
    if num_robot_total < 0 or world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad shard arguments")
    base, rem = divmod(num_robot_total, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_stats(stats, device=None):
    """Sum [sum reward, sum episode length, episodes, robot steps] over ranks (one 32-byte all-reduce)."""
    t = torch.as_tensor(stats, dtype=torch.float64, device=device).clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def broadcast_policy(params, src=0):
    """Broadcast a list of parameter tensors as ONE flat buffer (MpiAdam.sync semantics)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return params
    flat = torch.cat([p.reshape(-1) for p in params])
    dist.broadcast(flat, src=src)
    off = 0
    for p in params:
        n = p.numel()
        p.copy_(flat[off:off + n].view_as(p))
        off += n
    return params


def check_synced(params):
    """MpiAdam.check_synced (mpi_adam.py:72-83): every rank must hold rank 0's parameters."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return True
    flat = torch.cat([p.reshape(-1) for p in params])
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    ok = torch.tensor([1.0 if torch.equal(ref, flat) else 0.0], device=flat.device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return bool(ok.item() > 0.5)
