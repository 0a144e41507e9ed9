"""Timings of the non-headline BASELINE.json configurations (C3 MAXCUT, C4 beamforming ADMM, C5 circle packing) on one GPU,
through the host-buffer C ABI.  Parity for these shapes is covered by tests/; this script only reports throughput."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qcqp_b200 import engine, problems as pb

out = {}
# C3: MAXCUT G(2000, 0.1), 256 restarts (32 per GPU on 8 GPUs; here all 256 on one)
forms, info = pb.maxcut(2000, 0.1, seed=1)
pack = engine.Pack(forms)
X0 = np.random.RandomState(3).randn(256, 2000)
for R in (32, 256):
    best = 1e9
    for _ in range(3):
        rng = engine.rng_states(seeds=1000 + np.arange(R))
        t0 = time.perf_counter(); X, f0, mv, st = pack.cd_improve(X0[:R], rng, num_iters=100); best = min(best, time.perf_counter() - t0)
    sweeps = sum(s.steps_p1 + s.steps_p2 for s in st) / 2000.0
    out["C3_maxcut_n2000_R%d" % R] = dict(seconds=best, restart_sweeps=sweeps, restart_sweeps_per_s=sweeps / best,
                                          mean_cut=float(np.mean(-f0)), bytes_per_sweep=pack.info.bytes_per_sweep_phase2,
                                          note="num_iters=100 cap (balanced vertices make phase 2 random-walk forever, SURVEY H5)")
pack.close()
# C4: beamforming N=128, m=32, ADMM rho sweep of 16 values
forms, _ = pb.beamforming(n=64, m=24, l=8, seed=1)
pack = engine.Pack(forms)
rhos = np.sqrt(32) * 2.0 ** (np.arange(-8, 8) / 2.0)
np.random.seed(4); X0 = 2 * np.random.randn(1, 128)
pack.compute_eig()
best = 1e9
for _ in range(2):
    t0 = time.perf_counter(); X, f0, mv, st = pack.admm_improve(X0, rhos); best = min(best, time.perf_counter() - t0)
iters = sum(s.iters_p1 + s.iters_p2 for s in st)
out["C4_beamforming_N128_admm_16rho"] = dict(seconds=best, admm_iterations=int(iters), iterations_per_s=iters / best,
                                             onecons_calls=int(sum(s.onecons_calls for s in st)), feasible_runs=int((mv < 1e-2).sum()),
                                             best_f0=float(f0[mv < 1e-2].min()) if (mv < 1e-2).any() else None)
pack.close()
# C5: circle packing 200 circles, N=401, m=20701; 10 sweeps (mostly phase 1) and a phase-2-heavy run
forms, _ = pb.circle_packing(200)
pack = engine.Pack(forms)
for R, iters in ((64, 10), (512, 10), (512, 40)):
    X0 = np.random.RandomState(5).randn(R, 401)
    rng = engine.rng_states(seeds=np.arange(R))
    dt = 1e9
    for _rep in range(2):       # the first call at a new size grows the pack's workspace (cudaMalloc): time the second
        rng = engine.rng_states(seeds=np.arange(R))
        t0 = time.perf_counter(); X, f0, mv, st = pack.cd_improve(X0, rng, num_iters=iters); dt = min(dt, time.perf_counter() - t0)
    sw1 = sum(s.steps_p1 for s in st) / 401.0; sw2 = sum(s.steps_p2 for s in st) / 401.0
    out["C5_circle_200_R%d_%dsweeps" % (R, iters)] = dict(seconds=dt, restart_sweeps=sw1 + sw2, restart_sweeps_p1=sw1, restart_sweeps_p2=sw2,
                                                          restart_sweeps_per_s=(sw1 + sw2) / dt, max_violation=float(mv.max()),
                                                          best_r=float(-f0[mv < 1e-2].min()) if (mv < 1e-2).any() else None,
                                                          bytes_per_sweep=pack.info.bytes_per_sweep_phase2)
pack.close()
print(json.dumps(out, indent=1))
