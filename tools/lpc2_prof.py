"""Per-restart cycle breakdown of cd_lpc2_kernel (QCQP_LPC2_PROF): where the resolver and helper 0 of the slowest restarts spend time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
os.environ["QCQP_LPC2_PROF"] = "/tmp/lpc2_prof.bin"
from qcqp_b200 import engine, problems as pb
n, R = 1000, 1024
forms, _ = pb.boolean_least_squares(n, 1500, seed=1)
pack = engine.Pack(forms)
mu, _S, F = engine.sdr_factor(pb.synthetic_sdr_solution(n, rank=16, seed=5))
Z = np.random.RandomState(2).standard_normal((R, n))
X0, _f, _v = pack.sdr_sample_eval(mu, F, Z=Z)
for rep in range(2):
    rng = engine.rng_states(seeds=1000 + np.arange(R))
    X, f0, mv, st = pack.cd_improve(X0, rng)
a = np.fromfile("/tmp/lpc2_prof.bin", dtype=np.uint64).reshape(R, 8).astype(np.float64)
passes = np.array([np.ceil(s.steps_p2 / 32.0) for s in st]); moves = np.array([s.updates_p2 for s in st], dtype=float)
names = ["total", "prologue", "resolve", "mbar", "drain", "classify", "push", "rows"]
print("clock cycles per restart; NH=%s" % os.environ.get("QCQP_LPC2_NH", "2"))
order = np.argsort(-a[:, 0])
for tag, idx in (("mean of all", slice(None)), ("slowest 8", order[:8]), ("median 8", order[508:516])):
    m = a[idx].mean(axis=0); ps = passes[idx].mean(); mvs = moves[idx].mean()
    print("%-12s" % tag, "  ".join("%s %.0f" % (nm, v) for nm, v in zip(names, m)), " passes %.0f moves %.0f" % (ps, mvs))
    print("             per pass: prologue %.0f resolve %.0f mbar %.0f drain %.0f | per move: resolve %.0f = classify+ballot %.0f + push %.0f + rest %.0f | total/1.9GHz %.2f ms"
          % (m[1] / ps, m[2] / ps, m[3] / ps, m[4] / ps, m[2] / mvs, m[5] / mvs, m[6] / mvs, (m[2] - m[5] - m[6]) / mvs, m[0] / 1.9e6))
