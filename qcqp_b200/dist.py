"""Multi-GPU plumbing: one process per GPU (torchrun), restarts sharded contiguously, no collective on the data path;
one tiny reduction at the end picks the best point in QCQPForm.better order (utilities.py:135-146).
Works over NCCL (GPU tensors) and gloo (CPU tensors, used by the CPU test-suite)."""
import numpy as np


def shard_range(total, rank, world):
    """Contiguous slice [lo, hi) of `total` restarts owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def local_best(f0, maxviol, tol=1e-4):
    """(bucket, f0, local index) of the best local point: lexicographic min on (int(maxviol / tol), f0), later index wins
    exact ties -- the fold `best = better(best, x_r)` over r (better returns its second argument on a tie)."""
    f0 = np.asarray(f0, dtype=np.float64); mv = np.asarray(maxviol, dtype=np.float64)
    if f0.size == 0:
        return np.iinfo(np.int64).max, np.inf, -1
    bucket = (mv / tol).astype(np.int64)
    ok = ~np.isnan(f0)
    if not ok.any():
        return np.iinfo(np.int64).max, np.inf, -1
    b = bucket[ok].min()
    cand = ok & (bucket == b)
    f = f0[cand].min()
    idx = int(np.flatnonzero(cand & (f0 == f)).max())
    return int(b), float(f), idx


def global_best(bucket, f0, global_index, device=None):
    """All ranks call this with their local best; returns (bucket, f0, global index) of the overall best on every rank.
    Three scalar all-reduces (MIN bucket, MIN f0 among the bucket's holders, MAX index among exact ties)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return bucket, f0, global_index
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    tb = torch.tensor([bucket], dtype=torch.int64, device=dev)
    dist.all_reduce(tb, op=dist.ReduceOp.MIN)
    bmin = int(tb.item())
    tf = torch.tensor([f0 if (bucket == bmin and global_index >= 0) else float("inf")], dtype=torch.float64, device=dev)
    dist.all_reduce(tf, op=dist.ReduceOp.MIN)
    fmin = float(tf.item())
    ti = torch.tensor([global_index if (bucket == bmin and f0 == fmin) else -1], dtype=torch.int64, device=dev)
    dist.all_reduce(ti, op=dist.ReduceOp.MAX)
    return bmin, fmin, int(ti.item())


def owner_rank(mine, device=None):
    """The rank that holds the winner: every rank says whether the global best is its own (one MAX all-reduce)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([dist.get_rank() if mine else -1], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())


def broadcast_point(x, owner_rank, n, device=None):
    """The winner's x[n] from its owner to every rank (8 n bytes)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(x, dtype=np.float64)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.zeros(n, dtype=torch.float64, device=dev)
    if dist.get_rank() == owner_rank:
        t.copy_(torch.as_tensor(np.asarray(x, dtype=np.float64)))
    dist.broadcast(t, src=owner_rank)
    return t.cpu().numpy()
