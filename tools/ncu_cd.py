"""One short CD launch for ncu: C2 pack, R restarts, phase-2-only or full."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from qcqp_b200 import engine, problems as pb, _lib

n = int(os.environ.get("N", 1000)); R = int(os.environ.get("R", 1024))
mode = os.environ.get("MODE", "p2")
forms, _ = pb.boolean_least_squares(n, int(1.5 * n))
pack = engine.Pack(forms)
L = _lib.load()
dev = torch.device("cuda:0")
rs = np.random.RandomState(3)
X0 = rs.randn(R, n) if mode != "p2" else np.sign(rs.randn(R, n)) * np.sqrt(1 + 5e-3 * rs.rand(R, n))
rng = engine.rng_states(seeds=np.arange(R))
dX0 = torch.from_numpy(X0).to(dev)
drng = torch.from_numpy(engine.rng_states_as_tensor_bytes(rng)).to(dev)
dX = torch.empty_like(dX0); df = torch.empty(R, dtype=torch.float64, device=dev); dm = torch.empty_like(df)
dst = torch.zeros(R * C.sizeof(_lib.CdStats), dtype=torch.uint8, device=dev)
prm = _lib.CdParams(int(os.environ.get("ITERS", 3)), 1e-2, 1e-4, int(mode != "p2"), 0, 0)
for _ in range(int(os.environ.get("REPS", 2))):
    d2 = drng.clone()
    _lib.check(L.qcqp_cd_improve_device(pack.handle, C.byref(prm), dX0.data_ptr(), R, d2.data_ptr(), dX.data_ptr(), df.data_ptr(),
                                        dm.data_ptr(), dst.data_ptr(), torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("done")
