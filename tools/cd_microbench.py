"""Phase-by-phase timing of the CD kernel on the C2 workload (device-resident inputs, CUDA events)."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from qcqp_b200 import engine, problems as pb, _lib

n = int(os.environ.get("N", 1000))
forms, _ = pb.boolean_least_squares(n, int(1.5 * n))
pack = engine.Pack(forms)
L = _lib.load()
dev = torch.device("cuda:0")


def run(R, x0_kind, **kw):
    rs = np.random.RandomState(3)
    if x0_kind == "randn":
        X0 = rs.randn(R, n)
    else:  # feasible-ish start for phase-2-only timing
        X0 = np.sign(rs.randn(R, n)) * np.sqrt(1 + 5e-3 * rs.rand(R, n))
    rng = engine.rng_states(seeds=np.arange(R))
    dX0 = torch.from_numpy(X0).to(dev)
    drng = torch.from_numpy(engine.rng_states_as_tensor_bytes(rng)).to(dev)
    dX = torch.empty_like(dX0); df = torch.empty(R, dtype=torch.float64, device=dev); dm = torch.empty_like(df)
    dst = torch.zeros(R * C.sizeof(_lib.CdStats), dtype=torch.uint8, device=dev)
    prm = _lib.CdParams(kw.get("num_iters", 1000), 1e-2, 1e-4, int(kw.get("phase1", True)), int(os.environ.get("MODE", 0)), 0)
    stream = torch.cuda.current_stream().cuda_stream
    times = []
    for it in range(3):
        drng2 = drng.clone()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.qcqp_cd_improve_device(pack.handle, C.byref(prm), dX0.data_ptr(), R, drng2.data_ptr(), dX.data_ptr(), df.data_ptr(),
                                            dm.data_ptr(), dst.data_ptr(), stream))
        e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    st = np.frombuffer(dst.cpu().numpy().tobytes(), dtype=np.dtype([("s1", "<i8"), ("s2", "<i8"), ("u1", "<i8"), ("u2", "<i8"),
                       ("w1", "<i4"), ("w2", "<i4"), ("status", "<i4"), ("ran2", "<i4"), ("skip", "<i8")]))
    steps = int(st["s1"].sum() + st["s2"].sum())
    ms = min(times)
    print("R=%4d %-8s %-28s: %8.2f ms | steps p1 %d p2 %d | max sweeps p1 %d p2 %d mean p2 %.1f | %.0f restart-sweeps/s | per-step (max chain) %.2f us"
          % (R, x0_kind, str(kw), ms, st["s1"].sum(), st["s2"].sum(), st["w1"].max(), st["w2"].max(), st["w2"].mean(),
             steps / n / (ms * 1e-3), ms * 1e3 / max(1, (st["s1"] + st["s2"]).max())))


for R in (1024,):
    run(R, "randn", num_iters=1000)
    run(R, "randn", num_iters=1)          # one phase-1 sweep + at most one phase-2 sweep
    run(R, "feasible", phase1=False, num_iters=3)
    run(R, "feasible", phase1=False, num_iters=1000)
