"""Unit tests of the DEVICE 1-D solver (qcqp_b200/csrc/onevar.cuh) compiled for the host through the test-only shim
csrc/host_shim.cpp: golden quirk cases from the reference, then random cases against the oracle.  CPU only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "qcqp_b200", "csrc", "host_shim.cpp")
SO = os.path.join(ROOT, "qcqp_b200", "libqcqp_b200_hostshim.so")


@pytest.fixture(scope="module")
def shim():
    hdr = os.path.join(ROOT, "qcqp_b200", "csrc", "onevar.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", SRC, "-o", SO])
    L = C.CDLL(SO)
    L.qcqp_shim_onevar_qcqp.restype = C.c_int
    L.qcqp_shim_onevar_qcqp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_int32]
    L.qcqp_shim_feasible_intervals.restype = C.c_int
    L.qcqp_shim_feasible_intervals.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int32, C.c_double, C.c_void_p]
    L.qcqp_shim_phase1_bisect.restype = C.c_int
    L.qcqp_shim_phase1_bisect.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    L.qcqp_shim_uniform.restype = C.c_double
    L.qcqp_shim_uniform.argtypes = [C.c_void_p, C.c_double, C.c_double]
    L.qcqp_shim_choice.restype = C.c_int
    L.qcqp_shim_choice.argtypes = [C.c_void_p, C.c_int32]
    L.qcqp_shim_single_det.restype = C.c_int
    L.qcqp_shim_single_det.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def _solve(shim, f0, fs, s, st, force_general=0):
    f0a = np.ascontiguousarray(f0, dtype=np.float64)
    fa = np.ascontiguousarray([f[:3] for f in fs], dtype=np.float64).reshape(-1, 3)
    ra = np.ascontiguousarray([orc.RELOP_CODE[f[3]] for f in fs], dtype=np.int32)
    out = C.c_double()
    rc = shim.qcqp_shim_onevar_qcqp(f0a.ctypes.data, fa.ctypes.data, ra.ctypes.data, len(fs), float(s), C.byref(st), C.byref(out), force_general)
    return rc, out.value


def test_device_intervals_match_reference(shim, golden):
    for c in golden["intervals"]:
        out = np.empty(4)
        p, q, r, op = c["f"]
        n = shim.qcqp_shim_feasible_intervals(p, q, r, orc.RELOP_CODE[op], c["s"], out.ctypes.data)
        assert n == len(c["intervals"])
        for i, (a, b) in enumerate(c["intervals"]):
            assert out[2 * i] == a and out[2 * i + 1] == b


def test_device_solver_matches_reference_goldens(shim, golden):
    for c in golden["onevar"]:
        for mode in (0, 1, 2, 3):       # register path, sorted-event path, hole formulation, sort-free holes (cd_blk.cu phase 1)
            st = orc.RngState.from_seed(c["seed"])
            rc, x = _solve(shim, c["f0"], [tuple(f) for f in c["fs"]], c["s"], st, mode)
            if c["error"]:
                assert rc == -2
                continue
            if c["result"] is None:
                assert rc == 0, (mode, c)
            else:
                assert rc == 1 and x == c["result"], (mode, c)
            assert st.pos == c["rng"]["pos"], (mode, c)


def test_device_solver_matches_oracle_random(shim):
    """Wider net than the goldens: many constraints, duplicated constraints, shared endpoints, degenerate forms."""
    rs = np.random.RandomState(2024)
    for t in range(4000):
        m = int(rs.randint(0, 12)) if t % 7 else int(rs.randint(12, 60))
        fs = []
        for _ in range(m):
            kind = rs.randint(0, 7)
            if kind == 6:   # lattice coefficients -> exact endpoint coincidences between different constraints
                p, q, r = float(rs.randint(-2, 3)), float(rs.randint(-2, 3)), float(rs.randint(-4, 2))
            else:
                p = rs.randn() * (kind != 0) * (1e-5 if kind == 5 else 1.0)
                q = rs.randn() * (kind != 1)
                r = rs.randn() - 1.0
            if p == 0 and q == 0:
                q = 1.0     # (0, 0) forms are filtered by the caller (qcqp.py:116)
            fs.append((float(p), float(q), float(r), "==" if rs.rand() < 0.35 else "<="))
        if m > 1 and t % 5 == 0:
            fs[rs.randint(0, m)] = fs[rs.randint(0, m)]
        k0 = rs.randint(0, 5)
        f0 = (float(abs(rs.randn())) if k0 == 0 else (0.0 if k0 in (1, 2) else float(rs.randint(-2, 3))),
              0.0 if k0 == 2 else float(rs.randint(-3, 4)) if k0 == 4 else float(rs.randn()), float(rs.randn() * 10 ** rs.randint(0, 6)))
        s = float(rs.choice([0.0, 1e-4, 0.01, 0.3, 1.0, 2.5]) * (1 if t % 2 else abs(rs.randn())))
        st_a = orc.RngState.from_seed(t)
        st_b = orc.RngState.from_seed(t)
        try:
            want = orc.onevar_qcqp(f0, fs, s, st_a)
            err = False
        except OverflowError:
            err = True
        for force_general in (0, 1, 2, 3):    # register path (<= 1 two-interval constraint), sorted events, holes, sort-free holes
            st_b = orc.RngState.from_seed(t)
            rc, x = _solve(shim, f0, fs, s, st_b, force_general)
            if err:
                assert rc == -2, (t, f0, fs, s)
                continue
            if want is None:
                assert rc == 0, (t, f0, fs, s, x)
            else:
                assert rc == 1 and x == want, (t, f0, fs, s, x, want)
            assert st_a.pos == st_b.pos, (t, f0, fs, s)


def test_device_rng_transforms(shim):
    for seed in (3, 99):
        rs = np.random.RandomState(seed)
        st = orc.RngState.from_seed(seed)
        for t in range(2000):
            if t % 3:
                assert rs.uniform(-1.5, 2.25) == shim.qcqp_shim_uniform(C.byref(st), -1.5, 2.25)
            else:
                n = [1, 2, 3, 5, 8, 200][t % 6]
                assert int(rs.choice(n)) == shim.qcqp_shim_choice(C.byref(st), n)


def test_device_single_constraint_deterministic_choice(shim):
    """The separable fast path (cd_lpc.cu): with one constraint, the RNG-free decision either equals onevar_qcqp's result with
    the RNG untouched, or reports 'needs a draw' exactly when the reference draws."""
    rs = np.random.RandomState(77)
    for t in range(6000):
        kind = rs.randint(0, 7)
        if kind == 6:
            p, q, r = float(rs.randint(-2, 3)), float(rs.randint(-2, 3)), float(rs.randint(-4, 2))
        else:
            p = rs.randn() * (kind != 0) * (1e-5 if kind == 5 else 1.0); q = rs.randn() * (kind != 1); r = rs.randn() - 1.0
        if p == 0 and q == 0:
            q = 1.0
        rel = "==" if rs.rand() < 0.5 else "<="
        k0 = rs.randint(0, 5)
        f0 = (float(abs(rs.randn())) if k0 == 0 else (0.0 if k0 in (1, 2) else float(rs.randint(-2, 3))),
              0.0 if k0 == 2 else float(rs.randint(-3, 4)) if k0 == 4 else float(rs.randn()), float(rs.randn() * 10 ** rs.randint(0, 6)))
        s = float(rs.choice([0.0, 1e-4, 0.01, 0.3, 1.0, 2.5]) * (1 if t % 2 else abs(rs.randn())))
        st = orc.RngState.from_seed(t)
        pos0 = st.pos
        try:
            want = orc.onevar_qcqp(f0, [(p, q, r, rel)], s, st)
            err = False
        except OverflowError:
            err = True
        f0a = np.array(f0); fa = np.array([p, q, r]); out = C.c_double(); pieces = np.zeros(4); nC = C.c_int32()
        rc = shim.qcqp_shim_single_det(f0a.ctypes.data, fa.ctypes.data, orc.RELOP_CODE[rel], s, C.byref(out), pieces.ctypes.data, C.byref(nC))
        assert rc != -100, "finite-endpoint chooser disagrees with the general one"
        drew = err or st.pos != pos0
        if rc == 2:
            assert drew or (f0[0] == 0 and f0[1] == 0), (t, f0, (p, q, r, rel), s)
        elif rc == 1:
            assert not drew and want == out.value, (t, f0, (p, q, r, rel), s, want, out.value)
        else:
            assert want is None and not drew, (t, f0, (p, q, r, rel), s, want)


def test_hole_formulation_heavy_ties(shim):
    """The hole formulation (cd.cu's general path) on constraint sets built to collide: lattice coefficients, repeated constraints,
    '==' pairs sharing roots with '<=' holes, levels that make roots coincide.  Result and RNG consumption must equal the oracle's."""
    rs = np.random.RandomState(4242)
    for t in range(3000):
        m = int(rs.randint(2, 40))
        fs = []
        for _ in range(m):
            p = float(rs.choice([-2.0, -1.0, -1.0, -0.5, 0.0, 1.0, 2.0]))
            q = float(rs.randint(-3, 4))
            r = float(rs.randint(-6, 3)) * float(rs.choice([1.0, 0.5, 0.25]))
            if p == 0 and q == 0:
                q = 1.0
            fs.append((p, q, r, "==" if rs.rand() < 0.25 else "<="))
        for _ in range(rs.randint(0, 4)):
            fs[rs.randint(0, m)] = fs[rs.randint(0, m)]
        k0 = rs.randint(0, 4)
        f0 = (0.0, 0.0, 1.0) if k0 == 0 else (float(rs.randint(-2, 3)), float(rs.randint(-3, 4)), float(rs.randint(-3, 4)))
        s = float(rs.choice([0.0, 0.25, 0.5, 1.0, 2.0, 4.0, 7.0]))
        st_a = orc.RngState.from_seed(t)
        try:
            want = orc.onevar_qcqp(f0, fs, s, st_a)
            err = False
        except OverflowError:
            err = True
        for mode in (2, 3):                # sorted holes (cd.cu), sort-free holes (cd_blk.cu phase 1)
            st_b = orc.RngState.from_seed(t)
            rc, x = _solve(shim, f0, fs, s, st_b, mode)
            if err:
                assert rc == -2, (t, f0, fs, s)
                continue
            if want is None:
                assert rc == 0, (mode, t, f0, fs, s, x)
            else:
                assert rc == 1 and x == want, (mode, t, f0, fs, s, x, want)
            assert st_a.pos == st_b.pos, (mode, t, f0, fs, s)



def _ref_bisect(fs, ss, es, tol, st):
    """coord_descent_phase1's inner loop (qcqp.py:122-131) over the oracle's onevar_qcqp with the zero objective."""
    new_xi, new_viol, probes = 0.0, es, 0
    while es - ss > tol:
        s = (ss + es) / 2
        xi = orc.onevar_qcqp((0.0, 0.0, 0.0), fs, s, st)
        probes += 1
        if xi is None:
            ss = s
        else:
            new_xi, new_viol, es = xi, s, s
    return new_xi, new_viol, ss, es, probes


def test_phase1_bisection_by_solid_level_search_equals_the_reference_loop(shim):
    """cd_blk.cu's phase-1 bisection: instead of one probe per level, a search over the chain of levels in which a SOLID level (no
    open interval common to the feasible sets) certifies every lower level infeasible.  Scalar model (host_shim.cpp:
    qcqp_shim_phase1_bisect, the arithmetic of onevar.cuh) against the reference's loop over the oracle: same new_xi, new_viol,
    bracket and MT19937 position, for 2- and 4-level rounds, on circle-packing-like lists (concave constraints around centres plus
    a box), random mixed lists with equalities, and lattice lists built to tie (hidden pieces: non-solid levels that report nothing)."""
    rs = np.random.RandomState(99)
    tol = 1e-4
    saved, total, moved, hidden = 0, 0, 0, 0
    for t in range(1500):
        kind = t % 3
        fs = []
        if kind == 0:        # circle packing, one centre coordinate: -(x - c)^2 + w <= 0 for every other circle, r <= x <= 1 - r
            rad = abs(rs.randn()) * 0.3 + 0.02
            for _ in range(int(rs.randint(3, 60))):
                cx, dy = rs.randn() * 0.7 + 0.5, rs.randn() * 0.6
                fs.append((-1.0, 2 * cx, 4 * rad * rad - dy * dy - cx * cx, "<="))
            fs.append((0.0, -1.0, rad, "<=")); fs.append((0.0, 1.0, rad - 1.0, "<="))
        elif kind == 1:      # mixed random
            for _ in range(int(rs.randint(1, 25))):
                p = rs.randn() * (rs.rand() < 0.8); q = rs.randn(); r = rs.randn() - 0.5
                if p == 0 and q == 0:
                    q = 1.0
                fs.append((float(p), float(q), float(r), "==" if rs.rand() < 0.25 else "<="))
        else:                # lattice coefficients: exact coincidences of endpoints between constraints
            for _ in range(int(rs.randint(2, 20))):
                p = float(rs.choice([-2.0, -1.0, -1.0, -0.5, 0.0, 1.0, 2.0])); q = float(rs.randint(-3, 4))
                r = float(rs.randint(-6, 3)) * float(rs.choice([1.0, 0.5, 0.25]))
                if p == 0 and q == 0:
                    q = 1.0
                fs.append((p, q, r, "==" if rs.rand() < 0.2 else "<="))
            for _ in range(rs.randint(0, 3)):
                fs[rs.randint(0, len(fs))] = fs[rs.randint(0, len(fs))]
        ss, es = -tol, float(rs.choice([0.05, 0.3, 1.0, 2.0, 4.0, 7.0]) * (1.0 if kind == 2 else rs.rand() + 0.05))
        st_ref = orc.RngState.from_seed(t)
        try:
            want = _ref_bisect(fs, ss, es, tol, st_ref)
        except OverflowError:
            continue                                  # an unbounded piece: the reference raises (the kernels report it as an error status)
        fa = np.ascontiguousarray([f[:3] for f in fs], dtype=np.float64).reshape(-1, 3)
        ra = np.ascontiguousarray([orc.RELOP_CODE[f[3]] for f in fs], dtype=np.int32)
        for mode, nw in ((0, 1), (1, 2), (1, 4), (1, 8)):
            st = orc.RngState.from_seed(t)
            out = np.zeros(4)
            work = shim.qcqp_shim_phase1_bisect(fa.ctypes.data, ra.ctypes.data, len(fs), ss, es, tol, C.byref(st), mode, nw, out.ctypes.data)
            assert work >= 0, (t, mode, nw)
            assert st.pos == st_ref.pos, (t, mode, nw, fs, es)
            assert out[1] == want[1] and out[2] == want[2] and out[3] == want[3], (t, mode, nw, out, want)
            if want[1] != es:                         # some level was feasible: the point drawn there
                assert out[0] == want[0], (t, mode, nw, out, want)
            if mode == 1 and nw == 4:
                total += want[4]; saved += want[4] - min(work, want[4]); moved += int(want[1] != es)
    assert total > 10000 and moved > 300
    print("reference probes %d, levels the 4-wide search did not have to examine %d, brackets with a feasible level %d" % (total, saved, moved))


def test_feasible_sets_grow_with_the_level(shim):
    """The lemma under cd_blk.cu's compaction and its solid-level certificate: in floating point, as computed by
    get_feasible_intervals, the feasible set of a constraint at level s0 is contained in its feasible set at every s1 >= s0 -- for
    '<=' and '==' constraints, convex, concave, linear and degenerate, including levels one ulp apart and the discriminant's zero.
    In particular a concave constraint that is the whole line at s0 stays the whole line."""
    rs = np.random.RandomState(1234)
    EQ, LE = orc.RELOP_CODE["=="], orc.RELOP_CODE["<="]

    def sets(p, q, r, rel, s):
        out = np.zeros(4)
        n = shim.qcqp_shim_feasible_intervals(p, q, r, rel, s, out.ctypes.data)
        return [(out[2 * i], out[2 * i + 1]) for i in range(n)]

    whole = 0
    for t in range(40000):
        k = t % 5
        if k == 0:
            p, q, r = float(rs.choice([-2, -1, -0.5, 0.5, 1, 4])), float(rs.randint(-3, 4)), float(rs.randint(-6, 3)) * 0.5
        elif k == 1:
            p, q, r = rs.randn(), rs.randn(), rs.randn()
        elif k == 2:
            p, q, r = 0.0, float(rs.choice([-2.0, -1.0, 1.0, 0.5, 3.0])) * (1 if t % 2 else rs.rand() + 0.1), rs.randn()
        elif k == 3:
            p, q = float(rs.choice([-1.0, 1.0, -2.0, 0.7])), float(rs.randint(-4, 5))
            r = q * q / (4 * p) + float(rs.choice([0.0, 1e-16, -1e-16, 1e-9, -1e-9, 1e-3]))     # discriminant at / next to zero
        else:
            p, q, r = rs.randn() * 1e-4, rs.randn() * 1e-4, rs.randn()                            # around the 1e-4 classification tolerance
        rel = EQ if t % 7 == 0 else LE
        s0 = float(rs.choice([-1e-4, 0.0, 1e-4, 0.3, 1.0])) if t % 3 else float(rs.randn())
        ds = float(rs.choice([0.0, 1e-4, 0.1, 2.0])) if t % 4 else float(np.spacing(abs(s0)) * rs.randint(0, 4))
        s1 = s0 + ds
        A, B = sets(p, q, r, rel, s0), sets(p, q, r, rel, s1)
        for lo, hi in A:                               # every interval of the lower level lies inside one interval of the higher level
            assert any(blo <= lo and hi <= bhi for blo, bhi in B), (p, q, r, rel, s0, s1, A, B)
        if rel == LE and p < -1e-4 and A == [(-np.inf, np.inf)]:
            whole += 1
            assert B == [(-np.inf, np.inf)], (p, q, r, s0, s1)
    assert whole > 500
