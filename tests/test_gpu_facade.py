"""The Suggest-and-Improve facade on the GPU: the reference's example flows (examples/*.py), parity of the single-point
drop-in mode with the process-global np.random stream, batch mode, error behaviour."""
import numpy as np
import pytest

from helpers import rel_close

pytestmark = pytest.mark.gpu


def test_boolean_least_squares_example_flow():
    """examples/boolean_least_squares.py:6-38 with a supplied X* (the SDP solve stays on the host, out of scope)."""
    from oracle import oracle as orc
    import qcqp_b200 as Q
    from qcqp_b200 import problems as pb
    forms, info = pb.boolean_least_squares(10, 15, seed=1)
    qc = Q.QCQP(forms)
    qc.set_sdr_solution(pb.synthetic_sdr_solution(10, rank=3, seed=5))
    P = orc.Problem(forms)
    # reference-side replay of the same calls on the oracle, sharing the global stream semantics
    np.random.seed(3)
    f_s, v_s = qc.suggest(Q.SDR)
    x_s = qc.x.copy()
    f_cd, v_cd = qc.improve(Q.COORD_DESCENT)
    state_after = np.random.get_state()
    np.random.seed(3)
    mu, Sigma, F = orc.sdr_factor(pb.synthetic_sdr_solution(10, rank=3, seed=5))
    st = orc.RngState.from_numpy(np.random.mtrand._rand)
    xo, _z = orc.sdr_sample(mu, F, st)
    assert rel_close(x_s, xo, rtol=1e-10, atol=1e-12)
    assert rel_close(f_s, P.eval(0, xo), rtol=1e-10) and rel_close(v_s, P.max_violation(xo), rtol=1e-10)
    xc, so = P.improve_cd(xo, st, fast=True)
    assert rel_close(f_cd, P.eval(0, xc), rtol=1e-6) and rel_close(v_cd, P.max_violation(xc), rtol=1e-6, atol=1e-10)
    assert state_after[2] == st.pos and np.array_equal(state_after[1], np.frombuffer(st.key, dtype=np.uint32))
    # chained improves, as the example does
    f2, v2 = qc.improve([Q.COORD_DESCENT, Q.ADMM], phase1=False, rho=2.0, num_iters=50)
    assert v2 < 1e-2 + 1e-9 or f2 <= f_cd + 1e-6


def test_maxcut_maximize_and_batch():
    import qcqp_b200 as Q
    from qcqp_b200 import problems as pb
    forms, info = pb.maxcut(25, 0.2, seed=1)
    qc = Q.QCQP(forms, maximize=True)
    qc.set_sdr_solution(pb.synthetic_sdr_solution(25, rank=4, seed=2))
    np.random.seed(11)
    f, v = qc.suggest(Q.SDR, samples=64)
    assert qc.X.shape == (64, 25)
    fb, vb = qc.improve(Q.COORD_DESCENT, seed=100, num_iters=50)
    assert vb < 1e-2 and fb > 0           # a cut value, sign flipped back (qcqp.py:400,416)
    assert fb == qc.batch_f0[qc.best_index] and fb >= np.max(qc.batch_f0[qc.batch_maxviol < 1e-4 * (int(vb / 1e-4) + 1)]) - 1e-12
    W = info["W"]
    cut = 0.25 * (W.sum() - qc.x.dot(W).dot(qc.x))
    assert rel_close(cut, fb, rtol=1e-9)


def test_beamforming_admm_flow():
    import qcqp_b200 as Q
    from qcqp_b200 import problems as pb
    forms, _ = pb.beamforming(n=20, m=5, l=2, seed=1)
    qc = Q.QCQP(forms)
    np.random.seed(4)
    qc.suggest(Q.RANDOM)
    f, v = qc.improve(Q.ADMM, rho=np.sqrt(7))
    assert v <= 1e-2 and 5 < f < 40
    with pytest.raises(Exception, match="rho parameter is too small"):
        Q.QCQP([(forms[0][0] * -1.0, forms[0][1], 0.0, None)] + forms[1:]).improve(Q.ADMM, rho=1e-3)


def test_facade_errors():
    import qcqp_b200 as Q
    from qcqp_b200 import problems as pb
    forms, _ = pb.boolean_least_squares(6, 9, seed=1)
    qc = Q.QCQP(forms)
    with pytest.raises(Exception, match="Unknown suggest method"):
        qc.suggest("nope")
    with pytest.raises(Exception, match="Unknown improve method"):
        qc.improve("nope")
    with pytest.raises(Exception, match="DCCP package is not installed"):
        qc.suggest(Q.RANDOM); qc.improve(Q.DCCP)
    with pytest.raises(Exception, match="PyIpopt package is not installed"):
        qc.improve(Q.IPOPT)


def test_suggest_sdr_and_spectral_with_the_host_relaxation():
    """examples/boolean_least_squares.py / maxcut.py end to end with the bundled host SDP (qcqp_b200/relax.py):
    the relaxation value bounds every point the improve step returns."""
    import itertools
    import qcqp_b200 as Q
    from qcqp_b200 import problems as pb
    forms, _ = pb.boolean_least_squares(10, 15, seed=1)
    qc = Q.QCQP(forms)
    np.random.seed(1)
    qc.suggest(Q.SDR)
    assert qc.sdr_sol.shape == (11, 11) and abs(qc.sdr_sol[-1, -1] - 1) < 1e-9
    f_cd, v_cd = qc.improve(Q.COORD_DESCENT)
    P0 = np.asarray(forms[0][0].todense()); q0 = forms[0][1]; r0 = forms[0][2]
    best = min(np.array(s) @ P0 @ np.array(s) + q0 @ np.array(s) + r0 for s in itertools.product([-1.0, 1.0], repeat=10))
    assert qc.sdr_bound <= best + 1e-6 and qc.sdr_bound <= f_cd + 1e-3 and v_cd < 1e-2
    f_sp, v_sp = qc.suggest(Q.SPECTRAL)
    assert qc.spectral_bound <= qc.sdr_bound + 1e-6
    forms, info = pb.maxcut(25, 0.2, seed=1)
    qm = Q.QCQP(forms, maximize=True)
    np.random.seed(1)
    qm.suggest(Q.SDR, samples=32)
    f, v = qm.improve(Q.COORD_DESCENT, seed=3, num_iters=50)
    assert f <= qm.sdr_bound + 1e-3 and v < 1e-2      # SDR-based upper bound on the cut


def test_sdr_cd_pipeline_equals_separate_calls():
    """qcqp_sdr_cd_pipeline (draws stay on the device, MT19937 streams seeded on the device, cached SDR factor) returns
    bit for bit what qcqp_sdr_sample_eval -> qcqp_cd_improve -> qcqp_best return through host buffers."""
    from qcqp_b200 import engine, problems as pb
    n, S = 48, 40
    forms, _ = pb.boolean_least_squares(n, 70, seed=3)
    pack = engine.Pack(forms)
    mu, _Sg, F = engine.sdr_factor(pb.synthetic_sdr_solution(n, rank=5, seed=2))
    Z = np.random.RandomState(8).standard_normal((S, n))
    seeds = 77 + 3 * np.arange(S)
    X0, fs, vs = pack.sdr_sample_eval(mu, F, Z=Z)
    rng = engine.rng_states(seeds=seeds)
    X, f0, mv, st = pack.cd_improve(X0, rng)
    res = pack.sdr_cd_pipeline(seeds, mu=mu, F=F, Z=Z, want_draws=True, want_rng=True)
    assert np.array_equal(res["X0"], X0) and np.array_equal(res["f0_draw"], fs) and np.array_equal(res["maxviol_draw"], vs)
    assert np.array_equal(res["X"], X) and np.array_equal(res["f0"], f0) and np.array_equal(res["maxviol"], mv)
    assert res["best"] == engine.best(f0, mv)
    for s in range(S):
        assert bytes(res["rng"][s]) == bytes(rng[s])                         # device init_genrand == np.random.seed
        assert (res["stats"][s].steps_p1, res["stats"][s].steps_p2) == (st[s].steps_p1, st[s].steps_p2)
    res2 = pack.sdr_cd_pipeline(seeds, Z=Z)                                  # cached factor
    assert np.array_equal(res2["X"], X)
    pack2 = engine.Pack(forms)
    with pytest.raises(Exception):
        pack2.sdr_cd_pipeline(seeds, Z=Z)                                    # nothing cached yet
    pack.close(); pack2.close()


def test_pinned_result_array_is_filled_by_the_kernel():
    """Host-buffer calls whose result array X is PINNED take the in-kernel delivery path of cd_lpc2_kernel (every restart writes its
    finished point through the mapped alias of X; no trailing device-to-host copy of X): same bytes as the pageable path, for the
    pipeline and for qcqp_cd_improve, including restarts that never reach phase 2."""
    import torch
    from qcqp_b200 import engine, problems as pb
    n, S = 100, 96                                                           # R >= 64, n > 64: phase 2 runs in cd_lpc2_kernel
    forms, _ = pb.boolean_least_squares(n, 150, seed=3)
    pack = engine.Pack(forms)
    mu, _Sg, F = engine.sdr_factor(pb.synthetic_sdr_solution(n, rank=5, seed=2))
    Z = np.random.RandomState(8).standard_normal((S, n))
    seeds = 500 + np.arange(S)
    ref = pack.sdr_cd_pipeline(seeds, mu=mu, F=F, Z=Z, want_draws=True)
    pin = lambda *shape: torch.empty(shape, dtype=torch.float64).pin_memory().numpy()
    out = (pin(S, n), pin(S), pin(S))
    out[0][:] = np.nan
    res = pack.sdr_cd_pipeline(seeds, Z=Z, out=out)
    assert res["X"] is out[0] and np.array_equal(out[0], ref["X"]) and np.array_equal(out[1], ref["f0"]) and np.array_equal(out[2], ref["maxviol"])
    assert res["best"] == ref["best"]
    # qcqp_cd_improve with a pinned X: monkey-level check through ctypes (engine.cd_improve allocates pageable arrays)
    import ctypes as C
    from qcqp_b200 import _lib
    L = _lib.load()
    X0 = np.ascontiguousarray(ref["X0"]); Xp = pin(S, n); Xp[:] = np.nan
    f0 = np.empty(S); mv = np.empty(S)
    rng = engine.rng_states(seeds=seeds)
    prm = _lib.CdParams(1000, 1e-2, 1e-4, 1, 0, 0)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.check(L.qcqp_cd_improve(pack.handle, C.byref(prm), ptr(X0), S, C.cast(rng, C.c_void_p), ptr(Xp), ptr(f0), ptr(mv), None))
    assert np.array_equal(Xp, ref["X"]) and np.array_equal(f0, ref["f0"])
    pack.close()


def test_prefetched_standard_normals_give_the_same_run():
    """qcqp_sdr_prefetch: the standard normals of a later pipeline call uploaded ahead on a private stream.  The call that gets the very
    same (pinned) array consumes the prefetched copy -- oldest first when the same array was prefetched twice -- and returns the bytes
    of the plain call; a call with another array, or another S, ignores the prefetch and uploads as usual."""
    import torch
    from qcqp_b200 import engine, problems as pb
    n, S = 100, 96
    forms, _ = pb.boolean_least_squares(n, 150, seed=3)
    pack = engine.Pack(forms)
    mu, _Sg, F = engine.sdr_factor(pb.synthetic_sdr_solution(n, rank=5, seed=2))
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    Za, Zb = pin(np.random.RandomState(8).standard_normal((S, n))), pin(np.random.RandomState(9).standard_normal((S, n)))
    seeds = 500 + np.arange(S)
    ref_a = pack.sdr_cd_pipeline(seeds, mu=mu, F=F, Z=Za)
    ref_b = pack.sdr_cd_pipeline(seeds, Z=Zb)
    assert not np.array_equal(ref_a["X"], ref_b["X"])
    same = lambda r, q: np.array_equal(r["X"], q["X"]) and np.array_equal(r["f0"], q["f0"]) and np.array_equal(r["maxviol"], q["maxviol"]) and r["best"] == q["best"]
    # batch after batch: prefetch B, run A (plain upload), run B (prefetched), prefetch A twice, run A twice
    pack.sdr_prefetch(Zb)
    assert same(pack.sdr_cd_pipeline(seeds, Z=Za), ref_a)
    assert same(pack.sdr_cd_pipeline(seeds, Z=Zb), ref_b)
    pack.sdr_prefetch(Za); pack.sdr_prefetch(Za)
    assert same(pack.sdr_cd_pipeline(seeds, Z=Za), ref_a)
    assert same(pack.sdr_cd_pipeline(seeds, Z=Za), ref_a)
    # a prefetch nobody consumes, then a call with fewer draws of the same array: not a match, plain upload of its own rows
    pack.sdr_prefetch(Zb)
    half = pack.sdr_cd_pipeline(seeds[:64], Z=Zb[:64])
    assert np.array_equal(half["X"], ref_b["X"][:64])
    with pytest.raises(Exception):
        pack.sdr_prefetch(Zb[:, ::2])                 # not the contiguous array a pipeline call would receive
    pack.close()


def test_reserved_pack_runs_under_cuda_graph_capture():
    """After qcqp_pack_reserve the `_device` entry points only enqueue work: the whole step (SDR draws -> coordinate descent ->
    best pick) is captured into a CUDA graph on a side stream and replayed; same bytes as the eager calls."""
    import ctypes as C
    import torch
    from qcqp_b200 import _lib, engine, problems as pb
    L = _lib.load()
    dev = torch.device("cuda:0")
    n, R = 100, 128
    forms, _ = pb.boolean_least_squares(n, 150, seed=3)
    pack = engine.Pack(forms)
    _lib.check(L.qcqp_pack_reserve(pack.handle, R, 1))
    mu, _Sg, F = engine.sdr_factor(pb.synthetic_sdr_solution(n, rank=5, seed=2))
    Z = np.random.RandomState(8).standard_normal((R, n))
    rng_host = engine.rng_states(seeds=300 + np.arange(R))
    d_mu, d_F, d_Z = (torch.from_numpy(a).to(dev) for a in (mu, F, Z))
    d_rng0 = torch.from_numpy(engine.rng_states_as_tensor_bytes(rng_host)).to(dev); d_rng = torch.empty_like(d_rng0)
    d_X0 = torch.empty((R, n), dtype=torch.float64, device=dev); d_X = torch.empty_like(d_X0)
    d_f = torch.empty(R, dtype=torch.float64, device=dev); d_v = torch.empty_like(d_f); d_fs = torch.empty_like(d_f); d_vs = torch.empty_like(d_f)
    d_st = torch.zeros(R * C.sizeof(_lib.CdStats), dtype=torch.uint8, device=dev)
    d_best = torch.zeros(1, dtype=torch.int32, device=dev)
    prm = _lib.CdParams(1000, 1e-2, 1e-4, 1, 0, 0)

    def step(stream):
        d_rng.copy_(d_rng0)
        _lib.check(L.qcqp_sdr_sample_eval_device(pack.handle, d_mu.data_ptr(), d_F.data_ptr(), d_Z.data_ptr(), 0, R, d_X0.data_ptr(), d_fs.data_ptr(),
                                                  d_vs.data_ptr(), stream))
        _lib.check(L.qcqp_cd_improve_device(pack.handle, C.byref(prm), d_X0.data_ptr(), R, d_rng.data_ptr(), d_X.data_ptr(), d_f.data_ptr(), d_v.data_ptr(),
                                             d_st.data_ptr(), stream))
        _lib.check(L.qcqp_best_device(d_f.data_ptr(), d_v.data_ptr(), R, 1e-4, d_best.data_ptr(), None, None, stream))

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        step(side.cuda_stream)                                   # eager, warms every lazily created object (events, tensor map)
    side.synchronize()
    want = (d_X.clone(), d_f.clone(), d_v.clone(), int(d_best.item()))
    d_X.zero_(); d_f.zero_()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        step(side.cuda_stream)
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(d_X, want[0]) and torch.equal(d_f, want[1]) and torch.equal(d_v, want[2]) and int(d_best.item()) == want[3]
    pack.close()


def test_facade_suggest_improve_batch_equals_two_step_flow():
    """QCQP.suggest_improve(samples=S, seed=s) == suggest(SDR, samples=S) followed by improve(COORD_DESCENT, seed=s)."""
    from qcqp_b200 import QCQP, COORD_DESCENT, SDR, problems as pb
    n, S = 30, 12
    forms, _ = pb.boolean_least_squares(n, 45, seed=4)
    Xs = pb.synthetic_sdr_solution(n, rank=4, seed=1)
    a = QCQP(forms); a.set_sdr_solution(Xs)
    b = QCQP(forms); b.set_sdr_solution(Xs)
    np.random.seed(5)
    a.suggest(SDR, samples=S)
    fa, va = a.improve(COORD_DESCENT, seed=900)
    np.random.seed(5)
    fb, vb = b.suggest_improve(samples=S, seed=900)
    assert fa == fb and va == vb and np.array_equal(a.X, b.X) and a.best_index == b.best_index
    fb2, vb2 = b.suggest_improve(samples=S, seed=901)          # second call: factor already on the device
    assert np.isfinite(fb2) and vb2 < 1e-2
