#!/usr/bin/env python
"""The facade's batch mode on several GPUs of one box: one process per GPU, the S draws / restarts sharded over the ranks, one
best-pick reduction at the end.  The printed result does not depend on the number of GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        examples/torchrun_batch.py --n 1000 --m 1500 --samples 8192
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import qcqp_b200.model as cvx                                  # noqa: E402
from qcqp_b200 import QCQP                                     # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200)
    ap.add_argument("--m", type=int, default=300)
    ap.add_argument("--samples", type=int, default=1024)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)                          # the pack lives on the device that is current at QCQP(prob)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank = dist.get_rank() if dist.is_initialized() else 0

    np.random.seed(args.seed)                                  # every rank builds the same problem and shares the host stream
    A = np.random.randn(args.m, args.n)
    b = np.random.randn(args.m, 1)
    x = cvx.Variable(args.n)
    qcqp = QCQP(cvx.Problem(cvx.Minimize(cvx.sum_squares(A*x - b)), [cvx.square(x) == 1]))

    t0 = time.perf_counter()
    f, v = qcqp.suggest_improve(samples=args.samples, seed=1000)       # host SDP once, then one engine call per rank
    dt = time.perf_counter() - t0
    if rank == 0:
        print("best of %d draws on %d GPU(s): objective %.6f, violation %.3g, SDR bound %.6f, restart %d (rank %d), %.2f s incl. the host SDP"
              % (args.samples, dist.get_world_size() if dist.is_initialized() else 1, f, v, qcqp.sdr_bound, qcqp.best_index,
                 getattr(qcqp, "best_rank", 0), dt))
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
