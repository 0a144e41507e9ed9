"""GPU parity tests of batched eval, the SDR sampler, ADMM / one-constraint projection and best-pick, through the C ABI."""
import numpy as np
import pytest

from helpers import forms_of, rel_close, GEN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("gen,gargs", [
    ("bls", dict(n=10, m=15, seed=1)), ("bls", dict(n=130, m=200, seed=2)), ("maxcut", dict(n=60, p=0.1, seed=1)),
    ("circle", dict(ncirc=6)), ("beam", dict(n=20, m=5, l=2, seed=1)), ("random", dict(n=9, m=12, seed=5, density=0.6)),
])
def test_eval_matches_oracle(gen, gargs):
    from oracle import oracle as orc
    from qcqp_b200 import engine
    forms, _ = GEN[gen](**gargs)
    P = orc.Problem(forms); pack = engine.Pack(forms)
    rs = np.random.RandomState(0)
    X = rs.randn(37, P.n)
    fo, vo, violo = P.eval_batch(X, want_viol=True)
    fg, vg, violg = pack.eval(X, want_viol=True)
    assert rel_close(fg, fo, rtol=1e-12, atol=1e-10)
    assert rel_close(vg, vo, rtol=1e-12, atol=1e-10)
    assert rel_close(violg, violo, rtol=1e-12, atol=1e-10)
    # empty batch and a single point
    f1, v1 = pack.eval(X[:1])
    assert rel_close(f1, fo[:1], rtol=1e-12, atol=1e-10)
    pack.close()


def test_eval_golden_onevar_func_problems(golden):
    """f_j(x) against values the reference itself produced (QuadraticFunction.eval)."""
    from qcqp_b200 import engine
    for c in golden["onevar_func"]:
        forms, _ = forms_of(c)
        pack = engine.Pack(forms)
        x = np.array(c["x"])
        f0, mv, viol = pack.eval(x[None, :], want_viol=True)
        assert rel_close(f0[0], c["evals"][0], rtol=1e-12, atol=1e-12)
        for j in range(1, len(c["evals"])):
            want = abs(c["evals"][j]) if forms[j][3] == "==" else max(0.0, c["evals"][j])
            assert rel_close(viol[0, j - 1], want, rtol=1e-12, atol=1e-12)
        pack.close()


def test_sdr_sampler_golden_and_oracle(golden):
    """Parity mode: the caller's standard normals -> the reference's np.random.multivariate_normal draws and their (f0, v)."""
    from oracle import oracle as orc
    from qcqp_b200 import engine, problems as pb
    for c in golden["sdr"]:
        forms, _ = pb.boolean_least_squares(**c["gargs"])
        pack = engine.Pack(forms)
        Xs = pb.synthetic_sdr_solution(c["n"], rank=c["rank"], seed=c["xs_seed"])
        mu, Sigma, F = engine.sdr_factor(Xs)
        rs = np.random.RandomState(c["seed"])
        Z = np.stack([rs.standard_normal(c["n"]) for _ in c["draws"]])
        X, f0, mv = pack.sdr_sample_eval(mu, F, Z=Z)
        for i, d in enumerate(c["draws"]):
            assert rel_close(X[i], d["x"], rtol=1e-9, atol=1e-10)
            assert rel_close(f0[i], d["f0"], rtol=1e-9) and rel_close(mv[i], d["maxviol"], rtol=1e-9)
        pack.close()
    # larger, against the oracle
    forms, _ = pb.boolean_least_squares(150, 220, seed=3)
    P = orc.Problem(forms); pack = engine.Pack(forms)
    mu, Sigma, F = engine.sdr_factor(pb.synthetic_sdr_solution(150, rank=8, seed=1))
    Z = np.random.RandomState(2).standard_normal((64, 150))
    Xo, fo, vo = P.sdr_sample_eval(mu, F, Z)
    Xg, fg, vg = pack.sdr_sample_eval(mu, F, Z=Z)
    assert rel_close(Xg, Xo, rtol=1e-11, atol=1e-12) and rel_close(fg, fo, rtol=1e-10) and rel_close(vg, vo, rtol=1e-9, atol=1e-12)
    pack.close()


def test_sdr_device_rng_statistics():
    """Throughput mode (device Philox + Box-Muller): draws are N(mu, F'F) -- checked on the first two moments."""
    from qcqp_b200 import engine, problems as pb
    n = 24
    forms, _ = pb.boolean_least_squares(n, 30, seed=1)
    pack = engine.Pack(forms)
    mu, Sigma, F = engine.sdr_factor(pb.synthetic_sdr_solution(n, rank=6, seed=2), corrected=True)
    X, f0, mv = pack.sdr_sample_eval(mu, F, Z=None, S=20000, seed=7)
    assert np.abs(X.mean(0) - mu).max() < 0.03
    C = np.cov(X.T)
    assert np.abs(C - F.T.dot(F)).max() < 0.05
    X2, _, _ = pack.sdr_sample_eval(mu, F, Z=None, S=64, seed=7)
    assert np.array_equal(X2, X[:64])      # counter-based: a draw depends only on (seed, index)
    X3, _, _ = pack.sdr_sample_eval(mu, F, Z=None, S=16, seed=7)      # small batches take the warp-per-draw kernel
    assert np.allclose(X3, X[:16], rtol=1e-12, atol=1e-13)
    f_chk, v_chk = pack.eval(X[:50])
    assert rel_close(f0[:50], f_chk, rtol=1e-12) and rel_close(mv[:50], v_chk, rtol=1e-12, atol=1e-12)
    pack.close()


@pytest.mark.parametrize("kernel", ["resident", "run"])
def test_admm_goldens(golden, kernel, monkeypatch):
    """G4 and the other improve_admm goldens of the reference: 1e-6 on (objective, max violation), same number of
    one-constraint projections.  Both kernels: admm_res.cu (eigenbases resident in shared memory, CTA per constraint) and
    admm.cu (CTA per run)."""
    from qcqp_b200 import engine
    monkeypatch.setenv("QCQP_ADMM_KERNEL", kernel)
    for c in golden["admm"]:
        forms, _ = forms_of(c)
        pack = engine.Pack(forms)
        kw = dict(c["kwargs"])
        rho = kw.pop("rho", None)
        if rho is None:
            lmb = np.linalg.eigvalsh(np.asarray(forms[0][0].todense()))
            rho = 50. * (2. * (1. - lmb.min()) / pack.m if lmb.min() < 0 else 1. / pack.m)
        X, f0, mv, st = pack.admm_improve(np.array(c["x0"])[None, :], [rho], **kw)
        assert st[0].onecons_calls == c["onecons_calls"], (c["name"], st[0].onecons_calls, c["onecons_calls"])
        assert rel_close(f0[0, 0], c["f0"], rtol=1e-6, atol=1e-9), (c["name"], f0[0, 0], c["f0"])
        assert rel_close(mv[0, 0], c["maxviol"], rtol=1e-6, atol=1e-8), c["name"]
        pack.close()


@pytest.mark.parametrize("kernel", ["resident", "run"])
def test_admm_rho_sweep_matches_oracle(kernel, monkeypatch):
    """K rho values x R starts in one launch (the C4 shape, reduced): every run against the oracle."""
    from oracle import oracle as orc
    from qcqp_b200 import engine, problems as pb
    monkeypatch.setenv("QCQP_ADMM_KERNEL", kernel)
    forms, _ = pb.beamforming(n=12, m=6, l=3, seed=1)
    P = orc.Problem(forms); pack = engine.Pack(forms)
    rs = np.random.RandomState(4)
    X0 = 2 * rs.randn(3, P.n)
    rhos = np.sqrt(9) * 2.0 ** (np.arange(-2, 3) / 2.0)
    Xo, fo, vo, so = P.improve_admm_batch(X0, rhos, num_iters=300)
    Xg, fg, vg, sg = pack.admm_improve(X0, rhos, num_iters=300)
    for i in range(len(rhos) * 3):
        assert sg[i].iters_p1 == so[i].iters_p1 and sg[i].iters_p2 == so[i].iters_p2, i
    assert rel_close(fg, fo, rtol=1e-6, atol=1e-9) and rel_close(vg, vo, rtol=1e-6, atol=1e-8)
    pack.close()


def test_best_pick_order():
    """QCQPForm.better folded over a list: bucketised violation first, then objective, later index on exact ties."""
    from qcqp_b200 import engine
    from qcqp_b200.dist import local_best
    rs = np.random.RandomState(1)
    for t in range(50):
        R = int(rs.randint(1, 3000))
        f0 = np.round(rs.randn(R), 1 if t % 2 else 6)
        mv = np.abs(rs.randn(R)) * 10.0 ** rs.randint(-6, 0, size=R)
        if t % 5 == 0:
            f0[rs.randint(0, R)] = np.nan
        want = local_best(f0, mv)[2]
        assert engine.best(f0, mv) == want


def test_best_multi_single_rank_equals_best():
    """qcqp_best_multi on a one-rank communicator (the GPU box of the test tier has one device; tools/multi_probe.py is the
    two-rank run): same pick as qcqp_best, the winner's point and values returned, an empty shard and NaN objectives handled."""
    import torch
    from qcqp_b200 import engine
    from qcqp_b200.dist import local_best
    dev = torch.device("cuda:0")
    comm = engine.Comm(0, 1, engine.Comm.unique_id())
    rs = np.random.RandomState(3)
    for R, n in ((1, 4), (50, 7), (1000, 33)):
        f0 = np.round(rs.randn(R), 1); mv = np.abs(rs.randn(R)) * 2e-4; X = rs.randn(R, n)
        if R > 10:
            f0[3] = np.nan
        df, dv, dX = (torch.from_numpy(a).to(dev) for a in (f0, mv, X))
        dx = torch.zeros(n, dtype=torch.float64, device=dev)
        gi, rk, bf, bv = comm.best(df.data_ptr(), dv.data_ptr(), dX.data_ptr(), R, n, index_offset=100, d_xbest=dx.data_ptr())
        b = local_best(f0, mv)[2]
        assert (gi, rk) == (100 + b, 0) and bf == f0[b] and bv == mv[b] and np.array_equal(dx.cpu().numpy(), X[b])
    gi, rk, bf, bv = comm.best(0, 0, 0, 0, 5)              # nothing to offer
    assert gi == -1 and rk == -1 and np.isinf(bf)
    comm.close()


def test_admm_device_side_setup_matches_host_setup():
    """eigendecompositions of the constraint matrices as one batched GPU `eigh` and the per-rho inverses
    as one batched Cholesky, against the host (LAPACK) setup the reference uses.  NOT held to the 1e-6 parity bar of the default path:
    cuSOLVER and LAPACK return different bases of the degenerate eigenspaces (C4: 126 noise eigenvalues per constraint), the ADMM
    iterates differ at 1e-9 and the runs amplify that differently.  Measured: small instance, all 10 runs within 2.7e-5 absolute
    (values 7..76); C4 size (N = 128, 32 rank-2 constraints), 9 of 10 runs within 6e-7 relative with equal iteration counts, one
    run ends 0.3 % away in f0 (maxviol 0 vs 4e-5).  The test states exactly that: every run within 2 % and feasible to the same
    1e-3, at least 80 % of the runs within 1e-5 relative.  Hence opt-in; the default stays the host LAPACK setup."""
    import time
    from qcqp_b200 import engine, problems as pb
    for gargs, rhos in ((dict(n=12, m=6, l=3, seed=1), np.sqrt(9) * 2.0 ** (np.arange(-2, 3) / 2.0)),
                        (dict(n=64, m=24, l=8, seed=1), np.sqrt(32) * 2.0 ** (np.arange(-4, 5, 2) / 2.0))):
        forms, _ = pb.beamforming(**gargs)
        X0 = 2 * np.random.RandomState(4).randn(2, 2 * gargs["n"])
        a = engine.Pack(forms)
        t0 = time.time(); Xa, fa, va, sa = a.admm_improve(X0, rhos, setup="host"); th = time.time() - t0
        b = engine.Pack(forms)
        t0 = time.time(); Xb, fb, vb, sb = b.admm_improve(X0, rhos, setup="device"); td = time.time() - t0
        fa, fb, va, vb = (np.asarray(v, dtype=np.float64).ravel() for v in (fa, fb, va, vb))
        rel = np.abs(fa - fb) / np.maximum(np.abs(fa), 1e-9)
        assert rel.max() < 2e-2 and np.abs(va - vb).max() < 1e-3, (rel, va, vb)
        assert np.mean(rel < 1e-5) >= 0.8, rel
        same_iters = sum(int(sa[i].iters_p1 == sb[i].iters_p1 and sa[i].iters_p2 == sb[i].iters_p2) for i in range(len(sa)))
        print("N=%d m=%d: host setup %.2f s, device setup %.2f s; runs within 1e-5 relative: %d of %d (max %.2e); identical iteration counts in %d"
              % (2 * gargs["n"], a.m, th, td, int(np.sum(rel < 1e-5)), rel.size, rel.max(), same_iters))
        a.close(); b.close()
