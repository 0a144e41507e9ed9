"""Two (or more) ranks, one per GPU: the in-library best pick across GPUs (qcqp_best_multi, NCCL owned by the library) against the
host-side fold over all restarts.  Launch:  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/multi_probe.py
(torch.distributed / gloo only carries the 128-byte NCCL id to the other ranks.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from qcqp_b200 import engine
from qcqp_b200.dist import local_best, shard_range

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("gloo")
box = [engine.Comm.unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
comm = engine.Comm(rank, world, box[0])
dev = torch.device("cuda", lr)
ok = True
for trial, (R, n) in enumerate(((37, 5), (1024, 1000), (3, 2), (world - 1, 9))):
    rs = np.random.RandomState(trial)
    f0 = np.round(rs.randn(R), 1); mv = np.abs(rs.randn(R)) * 3e-4; X = rs.randn(R, n)
    lo, hi = shard_range(R, rank, world)
    df, dv, dX = (torch.from_numpy(np.ascontiguousarray(a[lo:hi])).to(dev) for a in (f0, mv, X))
    dx = torch.zeros(n, dtype=torch.float64, device=dev)
    gi, rk, bf, bv = comm.best(df.data_ptr() if hi > lo else 0, dv.data_ptr() if hi > lo else 0, dX.data_ptr() if hi > lo else 0, hi - lo, n,
                               index_offset=lo, d_xbest=dx.data_ptr())
    b = local_best(f0, mv)[2]
    owner = [r for r in range(world) if shard_range(R, r, world)[0] <= b < shard_range(R, r, world)[1]][0]
    good = (gi, rk, bf, bv) == (b, owner, f0[b], mv[b]) and np.array_equal(dx.cpu().numpy(), X[b])
    ok = ok and good
    if rank == 0:
        print("R=%d n=%d: global best %d on rank %d, f0 %.3f -- %s" % (R, n, gi, rk, bf, "matches the host fold" if good else "MISMATCH"))
t = torch.tensor([1 if ok else 0]); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("qcqp_best_multi over %d GPUs: %s" % (world, "OK" if int(t) == 1 else "FAILED"))
comm.close()
dist.destroy_process_group()
