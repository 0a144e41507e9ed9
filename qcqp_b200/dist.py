"""Multi-GPU plumbing: one process per GPU (torchrun), restarts sharded contiguously, no collective on the data path;
one tiny all-gather at the end picks the best point in QCQPForm.better order (utilities.py:135-146).
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


def global_best(bucket, f0, global_index, device=None, x=None):
    """All ranks call this with their local best; returns (bucket, f0, global index) of the overall best on every rank -- and,
    when every rank passes its best point `x` (length n), that point as a fourth value.  ONE all-gather of 24 (+ 8 n) bytes per
    rank and one synchronisation; the pick is the same fold as local_best (later global index wins exact ties), done on the host
    by every rank on identical data."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return (bucket, f0, global_index) if x is None else (bucket, f0, global_index, np.asarray(x, dtype=np.float64))
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    world = dist.get_world_size()
    head = np.array([float(bucket) if global_index >= 0 else np.inf, float(f0) if global_index >= 0 else np.inf, float(global_index)])
    mine = head if x is None else np.concatenate([head, np.asarray(x, dtype=np.float64).ravel()])
    send = torch.from_numpy(mine).to(dev)
    recv = torch.empty(world * mine.size, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv, send)
    tab = recv.cpu().numpy().reshape(world, mine.size)        # the one synchronisation
    global _LAST_TABLE
    _LAST_TABLE = tab
    best = -1
    for r in range(world):                                    # buckets are exact in a double up to 2^53
        if tab[r, 2] < 0 or np.isnan(tab[r, 1]):
            continue
        if best < 0 or (tab[r, 0], tab[r, 1]) < (tab[best, 0], tab[best, 1]) or \
                ((tab[r, 0], tab[r, 1]) == (tab[best, 0], tab[best, 1]) and tab[r, 2] > tab[best, 2]):
            best = r
    if best < 0:
        out = (np.iinfo(np.int64).max, np.inf, -1)
        return out if x is None else out + (None,)
    b = np.iinfo(np.int64).max if np.isinf(tab[best, 0]) else int(tab[best, 0])
    out = (b, float(tab[best, 1]), int(tab[best, 2]))
    return out if x is None else out + (tab[best, 3:].copy(),)


_LAST_TABLE = None


def last_gather_columns(first, count):
    """Columns [first, first + count) of the payload every rank sent in the last global_best(x=...) call (a copy of data that is
    identical on all ranks): side information that travels with the best-pick, e.g. per-rank counts of failed restarts."""
    return _LAST_TABLE[:, 3 + first:3 + first + count].copy()


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
