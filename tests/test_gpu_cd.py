"""GPU parity tests of the coordinate-descent engine against the oracle, through the C ABI (qcqp_cd_improve)."""
import numpy as np
import pytest

from helpers import forms_of, rel_close, GEN

pytestmark = pytest.mark.gpu


def _oracle_and_pack(forms):
    from oracle import oracle as orc
    from qcqp_b200 import engine
    return orc.Problem(forms), engine.Pack(forms)


def _run_both(forms, X0, seeds, strict, **kw):
    from oracle import oracle as orc
    from qcqp_b200 import engine
    P, pack = _oracle_and_pack(forms)
    R = X0.shape[0]
    rng_g = engine.rng_states(seeds=seeds)
    Xg, fg, vg, sg = pack.cd_improve(X0, rng_g, strict=strict, **kw)
    out = []
    for r in range(R):
        st = orc.RngState.from_seed(int(seeds[r]))
        xo, so = P.improve_cd(X0[r], st, fast=True, **kw)
        out.append((xo, P.eval(0, xo), P.max_violation(xo), so, st))
    pack.close()
    return (Xg, fg, vg, sg, rng_g), out


def test_golden_cd_cases_strict_bit_exact(golden):
    """Strict mode against the reference's own goldens: (f0, maxviol) to 1e-9 and the same MT19937 position."""
    from qcqp_b200 import engine
    for c in golden["cd"]:
        if c["only_phase"] is not None or c["name"] == "beam_cd":
            continue
        forms, _ = forms_of(c)
        pack = engine.Pack(forms)
        x0 = np.array(c["x0"])
        rs = np.random.RandomState(c["seed"]); rs.standard_normal(len(x0))
        rng = engine.rng_states(states=[rs.get_state()])
        X, f0, mv, st = pack.cd_improve(x0[None, :], rng, strict=True, **c["kwargs"])
        assert st[0].status == 0, c["name"]
        assert rel_close(f0[0], c["f0"], rtol=1e-9, atol=1e-9), (c["name"], f0[0], c["f0"])
        assert rel_close(mv[0], c["maxviol"], rtol=1e-6, atol=1e-9), (c["name"], mv[0], c["maxviol"])
        assert rel_close(X[0], c["x"], rtol=1e-9, atol=1e-9), c["name"]
        assert rng[0].pos == c["rng"]["pos"], (c["name"], rng[0].pos)
        pack.close()


@pytest.mark.parametrize("gen,gargs,R,kw", [
    ("bls", dict(n=10, m=15, seed=1), 8, {}),
    ("bls", dict(n=48, m=70, seed=3), 16, {}),
    ("bls", dict(n=130, m=180, seed=2), 6, {}),
    ("maxcut", dict(n=25, p=0.2, seed=1), 8, dict(num_iters=40)),
    ("maxcut", dict(n=60, p=0.1, seed=1), 8, dict(num_iters=30)),
    ("circle", dict(ncirc=3), 6, dict(num_iters=10)),
    ("circle", dict(ncirc=6), 4, dict(num_iters=5)),
    ("random", dict(n=6, m=4, seed=0), 6, dict(num_iters=20)),
    ("random", dict(n=9, m=12, seed=5, density=0.6), 6, dict(num_iters=15)),
])
def test_cd_strict_matches_oracle(gen, gargs, R, kw):
    """Same x0, same per-restart seed: strict mode reproduces the oracle's (cached-f) run: x to 1e-9, identical
    stream positions and step counts."""
    forms, _ = GEN[gen](**gargs)
    n = forms[0][1].size
    rs = np.random.RandomState(123)
    X0 = rs.randn(R, n) if gen != "circle" else np.abs(rs.randn(R, n)) * 3 + 0.5
    seeds = 1000 + np.arange(R)
    (Xg, fg, vg, sg, rng_g), out = _run_both(forms, X0, seeds, True, **kw)
    for r in range(R):
        xo, fo, vo, so, st = out[r]
        assert sg[r].status == so.status
        assert (sg[r].steps_p1, sg[r].steps_p2, sg[r].steps_skipped) == (so.steps_p1, so.steps_p2, so.steps_skipped), (r, sg[r].steps_p1, so.steps_p1, sg[r].steps_p2, so.steps_p2)
        assert rng_g[r].pos == st.pos, r
        assert rel_close(Xg[r], xo, rtol=1e-9, atol=1e-9), r
        assert rel_close(fg[r], fo, rtol=1e-9, atol=1e-9)
        assert rel_close(vg[r], vo, rtol=1e-6, atol=1e-10)


@pytest.mark.parametrize("mode", [0, 2, 3])
@pytest.mark.parametrize("gen,gargs,R,kw", [
    ("bls", dict(n=10, m=15, seed=1), 16, {}),
    ("bls", dict(n=33, m=50, seed=2), 16, {}),
    ("bls", dict(n=64, m=96, seed=1), 32, {}),
    ("bls", dict(n=200, m=300, seed=1), 16, {}),
    ("bls", dict(n=333, m=500, seed=5), 16, {}),          # odd n: padded rows
    ("beam", dict(n=8, m=4, l=2, seed=1), 8, dict(num_iters=2)),   # dense constraints (n=16); short: the instance is chaotic
    ("circle", dict(ncirc=8), 8, dict(num_iters=6)),
])
def test_cd_fast_matches_oracle_1e6(gen, gargs, R, kw, mode):
    """The production modes -- 0 (default): separable problems (Boolean LS here) run the lane-per-coordinate kernel
    cd_lpc.cu, others the general kernel with cached dense row dots; 3: the general kernel with cached row dots even for
    separable problems; 2: warp-parallel fma dot of the staged row at every step.
    North-star bar: 1e-6 relative on (objective, max violation); stream positions and step counts must also agree."""
    forms, _ = GEN[gen](**gargs)
    n = forms[0][1].size
    rs = np.random.RandomState(7)
    X0 = rs.randn(R, n) if gen != "circle" else np.abs(rs.randn(R, n)) * 3 + 0.5
    seeds = 500 + np.arange(R)
    (Xg, fg, vg, sg, rng_g), out = _run_both(forms, X0, seeds, mode, **kw)
    if gen == "beam":
        # chaotic instance (DESIGN.md "conditioning"): the reference's own result moves by 1e-6 under a 1e-15 perturbation
        # of x0, so only feasibility and descent are comparable across summation orders
        from oracle import oracle as orc
        P = orc.Problem(forms)
        for r in range(R):
            assert sg[r].status == 0 and (vg[r] < 1e-2 or vg[r] <= P.max_violation(X0[r]) + 1e-9)
        return
    bad = 0
    for r in range(R):
        xo, fo, vo, so, st = out[r]
        ok = rel_close(fg[r], fo, rtol=1e-6, atol=1e-9) and rel_close(vg[r], vo, rtol=1e-6, atol=1e-9)
        ok = ok and rng_g[r].pos == st.pos and (sg[r].steps_p1, sg[r].steps_p2, sg[r].steps_skipped) == (so.steps_p1, so.steps_p2, so.steps_skipped)
        bad += (not ok)
    assert bad == 0, "%d of %d restarts differ from the oracle beyond 1e-6" % (bad, R)


def test_cd_separable_kernel_edge_cases():
    """cd_lpc.cu: stream refills inside a pass (long phase-1 bisections), phase1=False, tiny n, the stuck-coordinate
    fast-forward, and a MAXCUT instance (flat objectives and ties go through the real stream)."""
    from oracle import oracle as orc
    from qcqp_b200 import engine
    # far-away starts: ~20 probes per coordinate -> several MT19937 refills per 32-coordinate pass
    forms, _ = GEN["bls"](n=100, m=150, seed=3)
    rs = np.random.RandomState(9)
    X0 = rs.randn(12, 100) * 300.0
    (Xg, fg, vg, sg, rng_g), out = _run_both(forms, X0, 40 + np.arange(12), 0)
    for r in range(12):
        xo, fo, vo, so, st = out[r]
        assert rng_g[r].pos == st.pos and (sg[r].steps_p1, sg[r].steps_p2) == (so.steps_p1, so.steps_p2)
        assert rel_close(fg[r], fo, rtol=1e-6) and rel_close(vg[r], vo, rtol=1e-6, atol=1e-10)
    # phase1=False from a nearly feasible point; n smaller than a warp
    forms, _ = GEN["bls"](n=7, m=12, seed=1)
    X0 = np.sign(rs.randn(6, 7)) * np.sqrt(1 + 4e-3 * rs.rand(6, 7))
    (Xg, fg, vg, sg, rng_g), out = _run_both(forms, X0, 70 + np.arange(6), 0, phase1=False)
    for r in range(6):
        xo, fo, vo, so, st = out[r]
        assert rng_g[r].pos == st.pos and sg[r].steps_p2 == so.steps_p2 and rel_close(fg[r], fo, rtol=1e-9)
    # stuck coordinate: fast-forward
    forms, _ = GEN["bls"](n=40, m=60, seed=1)
    x0 = np.ones(40); x0[17] = np.sqrt(1.01005)
    (Xg, fg, vg, sg, rng_g), out = _run_both(forms, x0[None, :], [5], 0, num_iters=30)
    assert sg[0].steps_skipped == out[0][3].steps_skipped > 0 and sg[0].steps_p1 == out[0][3].steps_p1 and np.array_equal(Xg[0], out[0][0])
    # MAXCUT: not bit-comparable in fast mode, but every run must end feasible with a sensible cut
    forms, info = GEN["maxcut"](n=60, p=0.1, seed=1)
    pack = engine.Pack(forms)
    X0 = rs.randn(32, 60)
    Xg, fg, vg, sg = pack.cd_improve(X0, engine.rng_states(seeds=np.arange(32)), num_iters=40)
    Xs, fs, vs, ss = pack.cd_improve(X0, engine.rng_states(seeds=np.arange(32)), num_iters=40, strict=1)
    assert all(s.status == 0 for s in sg) and vg.max() < 1e-2
    assert abs(np.mean(-fg) - np.mean(-fs)) < 0.05 * abs(np.mean(-fs))
    pack.close()


def test_cd_error_statuses():
    """A coordinate that no constraint touches makes the reference raise in phase 1 (qcqp.py:117): status EMPTY_MAX."""
    import scipy.sparse as sp
    from qcqp_b200 import engine
    n = 4
    forms = [(sp.identity(n, format="csr"), np.zeros(n), 0.0, None)]
    P = sp.csr_matrix(([1.0], ([0], [0])), shape=(n, n))
    forms.append((P, np.zeros(n), -1.0, "=="))
    pack = engine.Pack(forms)
    X, f0, mv, st = pack.cd_improve(np.full((1, n), 3.0), engine.rng_states(seeds=[1]))
    assert st[0].status == 1
    pack.close()


def test_phase1_fixed_point_is_fast_forwarded():
    """x_k^2 - 1 in (viol_tol, viol_tol + tol]: phase 1 can never move the coordinate (qcqp.py:122: the bisection loop does
    not start), so the reference burns all num_iters sweeps as no-ops.  The engine detects the fixed point, reports the
    steps it did not execute, and returns the same point and the same untouched RNG stream as the oracle."""
    from oracle import oracle as orc
    from qcqp_b200 import engine
    forms, _ = GEN["bls"](n=12, m=18, seed=1)
    x0 = np.ones(12)
    x0[3] = np.sqrt(1.01005)
    pack = engine.Pack(forms)
    rng = engine.rng_states(seeds=[5])
    X, f0, mv, st = pack.cd_improve(x0[None, :], rng, strict=True, num_iters=50)
    P = orc.Problem(forms)
    sto = orc.RngState.from_seed(5)
    xo, so = P.improve_cd(x0, sto, fast=False, num_iters=50)    # faithful mode executes every no-op sweep
    assert st[0].steps_skipped > 0 and st[0].steps_p1 + st[0].steps_skipped == so.steps_p1
    assert np.array_equal(X[0], xo) and rng[0].pos == sto.pos and st[0].ran_phase2 == so.ran_phase2 == 0
    pack.close()


# ---------------------------------------------------------------------------------------------------------
# cd_blk.cu: one CTA per restart (many-incidence sparse problems).  strict = 5 / 4 force it at any size.
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("threads", [128, 256, 512])
@pytest.mark.parametrize("gen,gargs,R,kw", [
    ("circle", dict(ncirc=3), 6, dict(num_iters=10)),
    ("circle", dict(ncirc=6), 4, dict(num_iters=5)),
    ("circle", dict(ncirc=20), 4, dict(num_iters=4)),      # 21 incidences per centre coordinate: register-sorted holes
    ("circle", dict(ncirc=40), 3, dict(num_iters=3)),      # 39 holes per centre coordinate: block sort + chunked scan
    ("circle", dict(ncirc=50), 2, dict(num_iters=2)),      # radius meets 1327 forms: coefficients recomputed per probe
    ("random", dict(n=6, m=4, seed=0), 6, dict(num_iters=20)),
    ("random", dict(n=9, m=12, seed=5, density=0.6), 6, dict(num_iters=15)),
    ("random", dict(n=12, m=80, seed=7, density=0.5), 4, dict(num_iters=6)),   # up to 80 mixed-relop constraints per coordinate
    ("maxcut", dict(n=25, p=0.2, seed=1), 4, dict(num_iters=20)),
])
def test_cd_blk_strict_matches_oracle(gen, gargs, R, kw, threads, monkeypatch):
    """The CTA-per-restart kernel reproduces the oracle's run like the warp-per-restart kernel does: x to 1e-9, identical
    stream positions and step counts, at every CTA width."""
    monkeypatch.setenv("QCQP_BLK_THREADS", str(threads))
    forms, _ = GEN[gen](**gargs)
    n = forms[0][1].size
    rs = np.random.RandomState(321)
    X0 = rs.randn(R, n) if gen != "circle" else np.abs(rs.randn(R, n)) * 3 + 0.5
    seeds = 2000 + np.arange(R)
    (Xg, fg, vg, sg, rng_g), out = _run_both(forms, X0, seeds, 5, **kw)
    for r in range(R):
        xo, fo, vo, so, st = out[r]
        assert sg[r].status == so.status
        assert (sg[r].steps_p1, sg[r].steps_p2, sg[r].steps_skipped) == (so.steps_p1, so.steps_p2, so.steps_skipped), (r, sg[r].steps_p1, so.steps_p1, sg[r].steps_p2, so.steps_p2)
        assert rng_g[r].pos == st.pos, r
        assert rel_close(Xg[r], xo, rtol=1e-9, atol=1e-9), r
        assert rel_close(fg[r], fo, rtol=1e-9, atol=1e-9)
        assert rel_close(vg[r], vo, rtol=1e-6, atol=1e-10)


def test_cd_blk_equals_warp_kernel_bitwise():
    """Both general kernels take the same decisions with the same arithmetic: identical x, f0, stream, statistics."""
    from qcqp_b200 import engine
    for gen, gargs, kw in [("circle", dict(ncirc=30), dict(num_iters=4)), ("random", dict(n=10, m=40, seed=3, density=0.7), dict(num_iters=8))]:
        forms, _ = GEN[gen](**gargs)
        n = forms[0][1].size
        rs = np.random.RandomState(11)
        X0 = rs.randn(5, n) if gen != "circle" else np.abs(rs.randn(5, n)) * 3 + 0.5
        pack = engine.Pack(forms)
        assert not pack_blk_default(pack)      # strict=1 runs the warp-per-restart kernel on these
        ra, rb = engine.rng_states(seeds=np.arange(5)), engine.rng_states(seeds=np.arange(5))
        Xa, fa, va, sa = pack.cd_improve(X0, ra, strict=1, **kw)
        Xb, fb, vb, sb = pack.cd_improve(X0, rb, strict=5, **kw)
        assert np.array_equal(Xa, Xb) and np.array_equal(fa, fb) and np.array_equal(va, vb)
        for r in range(5):
            assert ra[r].pos == rb[r].pos and (sa[r].steps_p1, sa[r].steps_p2, sa[r].updates_p2) == (sb[r].steps_p1, sb[r].steps_p2, sb[r].updates_p2)
        pack.close()


def pack_blk_default(pack):
    return pack.info.n_dense == 0 and pack.info.incidences >= 48 * pack.n


def test_cd_blk_default_dispatch_circle_60():
    """60 circles: 61 incidences per centre coordinate on average -> qcqp_cd_improve picks the CTA-per-restart kernel by
    itself (strict 0 and 1); fast mode within 1e-6 of the oracle, same step counts and stream."""
    forms, _ = GEN["circle"](ncirc=60)
    from qcqp_b200 import engine
    pack = engine.Pack(forms)
    assert pack_blk_default(pack)
    pack.close()
    rs = np.random.RandomState(4)
    X0 = np.abs(rs.randn(3, 121)) * 3 + 0.5
    for strict in (0, 1):
        (Xg, fg, vg, sg, rng_g), out = _run_both(forms, X0, 10 + np.arange(3), strict, num_iters=2)
        for r in range(3):
            xo, fo, vo, so, st = out[r]
            assert rng_g[r].pos == st.pos and (sg[r].steps_p1, sg[r].steps_p2) == (so.steps_p1, so.steps_p2)
            assert rel_close(fg[r], fo, rtol=1e-6, atol=1e-9) and rel_close(vg[r], vo, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("gen,gargs,R,kw,randn_start", [
    ("circle", dict(ncirc=200), 24, dict(num_iters=6), True),                        # C5's instance: 40 of 201 constraints kept per centre coordinate
    ("circle", dict(ncirc=70), 8, dict(num_iters=8), False),                         # phase 1 reaches feasibility for some restarts, phase 2 follows
    ("random", dict(n=14, m=220, seed=9, density=0.8), 6, dict(num_iters=6), True),  # > 128 kept constraints on most coordinates: CTA-wide fallback
    ("random", dict(n=12, m=100, seed=2, density=0.7), 6, dict(num_iters=8), True),  # 33..128 kept, mixed relops: several hole chunks per probe
])
def test_cd_blk_warp_probes_equal_cta_wide_probes(gen, gargs, R, kw, randn_start, monkeypatch):
    """Phase 1 of cd_blk_kernel with the compacted, warp-local, speculative probes (default) against every probe CTA-wide over the full
    constraint list (QCQP_BLK_WARP=0, the round-1 path that the oracle tests above also hold): identical bits of x, f0, maxviol,
    statistics and MT19937 position."""
    from qcqp_b200 import engine
    forms, _ = GEN[gen](**gargs)
    n = forms[0][1].size
    rs = np.random.RandomState(77)
    X0 = rs.randn(R, n) if randn_start else np.abs(rs.randn(R, n)) * 3 + 0.5
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("QCQP_BLK_WARP", mode)
        pack = engine.Pack(forms)
        rng = engine.rng_states(seeds=500 + np.arange(R))
        X, f0, mv, st = pack.cd_improve(X0, rng, strict=4, **kw)
        out[mode] = (X.copy(), f0.copy(), mv.copy(), [(s.steps_p1, s.steps_p2, s.updates_p1, s.updates_p2, s.sweeps_p1, s.sweeps_p2, s.status, s.steps_skipped) for s in st],
                     [r.pos for r in rng])
        pack.close()
    a, b = out["1"], out["0"]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert a[3] == b[3] and a[4] == b[4]
    assert sum(s[2] for s in a[3]) > 0          # phase 1 did move something
