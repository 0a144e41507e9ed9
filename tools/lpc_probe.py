"""Marginal cost per phase-2 sweep of the separable CD kernel on the C2 workload, and move statistics."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from qcqp_b200 import engine, problems as pb, _lib
n = 1000; R = 1024
forms, _ = pb.boolean_least_squares(n, 1500)
pack = engine.Pack(forms); L = _lib.load(); dev = torch.device("cuda:0")
rs = np.random.RandomState(3)
X0 = rs.randn(R, n)
rng = engine.rng_states(seeds=np.arange(R))
dX0 = torch.from_numpy(X0).to(dev); drng = torch.from_numpy(engine.rng_states_as_tensor_bytes(rng)).to(dev)
dX = torch.empty_like(dX0); df = torch.empty(R, dtype=torch.float64, device=dev); dm = torch.empty_like(df)
dst = torch.zeros(R * C.sizeof(_lib.CdStats), dtype=torch.uint8, device=dev)
DT = np.dtype([("s1", "<i8"), ("s2", "<i8"), ("u1", "<i8"), ("u2", "<i8"), ("w1", "<i4"), ("w2", "<i4"), ("status", "<i4"), ("ran2", "<i4"), ("skip", "<i8")])
def run(iters, phase1=1, x0=None):
    prm = _lib.CdParams(iters, 1e-2, 1e-4, phase1, 0, 0)
    src = dX0 if x0 is None else x0
    best = 1e9
    for _ in range(3):
        d2 = drng.clone(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.qcqp_cd_improve_device(pack.handle, C.byref(prm), src.data_ptr(), R, d2.data_ptr(), dX.data_ptr(), df.data_ptr(), dm.data_ptr(), dst.data_ptr(), torch.cuda.current_stream().cuda_stream))
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    st = np.frombuffer(dst.cpu().numpy().tobytes(), dtype=DT)
    return best, st
prev = None
for it in (1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 64, 1000):
    ms, st = run(it)
    print("num_iters=%4d: %7.3f ms | p2 sweeps mean %.2f max %d | updates p2 mean %.1f max %d | still running after cap: %d"
          % (it, ms, st["w2"].mean(), st["w2"].max(), st["u2"].mean(), st["u2"].max(), int((st["w2"] >= it).sum())))
