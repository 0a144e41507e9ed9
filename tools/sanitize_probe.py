"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qcqp_b200 import engine, problems as pb
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "admm"):
    forms, _ = pb.beamforming(n=12, m=6, l=3, seed=1)
    pack = engine.Pack(forms)
    X0 = 2 * np.random.RandomState(4).randn(3, pack.n)
    rhos = np.sqrt(9) * 2.0 ** (np.arange(-2, 3) / 2.0)
    X, f0, mv, st = pack.admm_improve(X0, rhos, num_iters=25)
    print("admm", [(s.iters_p1, s.iters_p2) for s in st][:4], float(f0.min()))
    pack.close()
if which in ("all", "pipe"):
    n, S = 48, 40
    forms, _ = pb.boolean_least_squares(n, 70, seed=3)
    pack = engine.Pack(forms)
    mu, _Sg, F = engine.sdr_factor(pb.synthetic_sdr_solution(n, rank=5, seed=2))
    Z = np.random.RandomState(8).standard_normal((S, n))
    res = pack.sdr_cd_pipeline(77 + np.arange(S), mu=mu, F=F, Z=Z, want_draws=True, want_rng=True)
    print("pipeline best", res["best"], float(res["f0"].min()))
    pack.close()
if which in ("all", "lpc2"):
    # separable dense objective, R >= 64, n > 64: phase 1 in cd_lpc_kernel (its lanes cross several 624-word MT19937 refills),
    # dgemm_mma_kernel for G = X P0, phase 2 in cd_lpc2_kernel (TMA blocks / rows, mbarrier command queue), batched eval
    n, S = 100, 64
    forms, _ = pb.boolean_least_squares(n, 150, seed=3)
    pack = engine.Pack(forms)
    mu, _Sg, F = engine.sdr_factor(pb.synthetic_sdr_solution(n, rank=5, seed=2))
    Z = np.random.RandomState(8).standard_normal((S, n))
    res = pack.sdr_cd_pipeline(77 + np.arange(S), mu=mu, F=F, Z=Z)
    print("lpc2 pipeline best", res["best"], float(res["f0"].min()), sum(s.updates_p2 for s in res["stats"]))
    pack.close()
if which in ("all", "cd"):
    # general kernel with dense constraints: producer warp + cp.async.bulk ring (cd_kernel)
    forms, _ = pb.beamforming(n=10, m=4, l=2, seed=1)
    pack = engine.Pack(forms)
    X0 = np.random.RandomState(2).randn(9, pack.n)
    for mode in (0, 2):
        X, f0, mv, st = pack.cd_improve(X0, engine.rng_states(seeds=np.arange(9)), num_iters=3, strict=mode)
        print("cd_kernel mode", mode, float(f0.min()), st[0].steps_p1, st[0].steps_p2)
    pack.close()
if which in ("all", "blk"):
    forms, _ = pb.circle_packing(40)
    pack = engine.Pack(forms)
    X0 = np.abs(np.random.RandomState(1).randn(2, 81)) * 3 + 0.5
    for T in ("128", "256", "512"):
        os.environ["QCQP_BLK_THREADS"] = T
        X, f0, mv, st = pack.cd_improve(X0, engine.rng_states(seeds=[1, 2]), num_iters=2, strict=4)
        print("blk", T, float(f0[0]), st[0].steps_p1, st[0].steps_p2)
    pack.close()
