"""Mints tests/golden/c4_admm_oracle.json: the CPU oracle (oracle/qcqp_oracle.c) on BASELINE configuration C4 -- beamforming
n=64 antennas (N=128 real variables), 32 constraints, improve_admm over the 16-value rho sweep plus rho = 2 exactly.
Run here (CPU, ~10 s); the GPU tests compare both ADMM kernels with it."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
from oracle import oracle as orc
from qcqp_b200 import problems as pb

forms, _ = pb.beamforming(n=64, m=24, l=8, seed=1)
P = orc.Problem(forms)
rhos = np.concatenate([np.sqrt(32) * 2.0 ** (np.arange(-8, 8) / 2.0), [2.0]])
np.random.seed(4)
X0 = 2 * np.random.randn(1, 128)
Xo, fo, vo, so = P.improve_admm_batch(X0, rhos)
out = dict(rhos=rhos.tolist(), x0=X0[0].tolist(), f0=np.asarray(fo).ravel().tolist(), maxviol=np.asarray(vo).ravel().tolist(),
           iters_p1=[int(s.iters_p1) for s in so], iters_p2=[int(s.iters_p2) for s in so])
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "c4_admm_oracle.json"), "w"))
print(out["iters_p1"], out["iters_p2"])
