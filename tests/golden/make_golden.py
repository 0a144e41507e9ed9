#!/usr/bin/env python
"""Mints tests/golden/golden.json by EXECUTING THE UNMODIFIED REFERENCE (/root/reference, behind the
cvxpy stub of oracle/ref_harness.py).  Run in the build container only:

    python tests/golden/make_golden.py

The reference has no test suite for this path ("parity unpinned" by its own tests), so these
vectors are what pins oracle/qcqp_oracle.c.  Problems are regenerated from (generator, args) by
qcqp_b200.problems; a checksum of the stacked data guards against generator drift.
"""
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from qcqp_b200 import problems as pb  # noqa: E402

warnings.filterwarnings("ignore")
u, q = rh.load()

GEN = {
    "bls": pb.boolean_least_squares,
    "maxcut": pb.maxcut,
    "beam": pb.beamforming,
    "circle": pb.circle_packing,
    "random": pb.random_qcqp,
}


def checksum(forms):
    acc = 0.0
    for j, (P, qv, r, _op) in enumerate(forms):
        P = P.tocsr()
        acc += (j + 1) * (float(np.abs(P.data).sum()) + float(np.abs(np.asarray(qv)).sum()) + abs(float(r)))
    return acc


def rng_tail():
    """Position of the process-global MT19937 stream after the call (and a peek at its next output)."""
    st = np.random.get_state()
    nxt = float(np.random.random_sample())
    np.random.set_state(st)
    return dict(pos=int(st[2]), has_gauss=int(st[3]), key_crc=int(np.bitwise_xor.reduce(st[1])), next_double=nxt)


def lst(a):
    return [float(v) for v in np.asarray(a, dtype=float).ravel()]


def cd_case(name, gen, gargs, x0_recipe, seed, kwargs, only_phase=None):
    forms, _ = GEN[gen](**gargs)
    prob = rh.make_form(u, forms)
    n = prob.n
    np.random.seed(seed)
    if x0_recipe == "randn":
        x0 = np.random.randn(n)
    elif x0_recipe == "sign":
        x0 = np.sign(np.random.randn(n))
    elif x0_recipe == "sign_perturbed":
        s = np.sign(np.random.randn(n))
        x0 = s * np.sqrt(1 + 5e-3 * (np.arange(n) + 1) / n)
    elif x0_recipe == "randn_x2":
        x0 = 2 * np.random.randn(n)
    elif x0_recipe == "circle":
        x0 = np.abs(np.random.randn(n)) * 3 + 0.5
        x0[0] = 0.3
    else:
        raise ValueError(x0_recipe)
    state0 = np.random.get_state()
    if only_phase == 1:
        x = q.coord_descent_phase1(x0, prob, kwargs.get("num_iters", 1000), kwargs.get("viol_tol", 1e-2), kwargs.get("tol", 1e-4))
    elif only_phase == 2:
        x = q.coord_descent_phase2(x0, prob, kwargs.get("num_iters", 1000), kwargs.get("viol_tol", 1e-2), kwargs.get("tol", 1e-4))
    else:
        x = q.improve_coord_descent(x0, prob, **kwargs)
    tail = rng_tail()
    untouched = bool(np.array_equal(state0[1], np.random.get_state()[1]) and state0[2] == np.random.get_state()[2])
    return dict(name=name, gen=gen, gargs=gargs, checksum=checksum(forms), seed=seed, x0=lst(x0), kwargs=kwargs,
                only_phase=only_phase, x=lst(x), f0=float(prob.f0.eval(x)), maxviol=float(max(prob.violations(x))),
                f0_start=float(prob.f0.eval(x0)), rng=tail, rng_untouched=untouched)


def onevar_cases():
    out = []

    def run(f0, fs, s, seed=3):
        np.random.seed(seed)
        o = u.OneVarQuadraticFunction(*f0)
        cs = [u.OneVarQuadraticFunction(*c) for c in fs]
        try:
            res = u.onevar_qcqp(o, cs, s)
            err = None
        except OverflowError as e:
            res, err = None, "OverflowError"
        return dict(f0=list(f0), fs=[list(c) for c in fs], s=s, seed=seed, result=(None if res is None else float(res)),
                    error=err, rng=rng_tail())

    # Q1..Q4 of SURVEY section 8c
    out.append(run((1, 0, 0), [(0, -1, 1, "<=")], 0))
    out.append(run((1, 0, 0), [(0, 1, 1, "<=")], 0))
    out.append(run((1, -6, 9), [(1, 0, -1, "<="), (1, 0, -1, "<=")], 0))
    out.append(run((1, -6, 9), [(1, 0, -1, "<=")], 0))
    out.append(run((1, -.2, .01), [(-1, 0, 1, "<="), (1, 0, -4, "<=")], 0))
    out.append(run((1, 0, 0), [], 0))                       # m = 0
    out.append(run((0, 0, 0), [(1, 0, -1, "==")], 0.1))     # flat objective: choice + uniform
    out.append(run((0, 1e-17, -5e4), [(1, 0, -1, "==")], 1e-4))  # absorbed ties -> choice over 4 endpoints
    out.append(run((-1, 0.3, 0), [(1, 0, -1, "==")], 0.05))
    out.append(run((0, 0, 0), [(0, 1, 1, "<=")], 0))        # uniform(-inf, a) -> OverflowError
    rs = np.random.RandomState(99)
    for t in range(400):
        m = int(rs.randint(1, 7))
        fs = []
        for _ in range(m):
            kind = rs.randint(0, 6)
            p = rs.randn() * (kind != 0) * (1.0 if kind != 5 else 1e-5)
            qq = rs.randn() * (kind != 1)
            r = rs.randn() - 1.0
            relop = "==" if rs.rand() < 0.35 else "<="
            fs.append((float(p), float(qq), float(r), relop))
        if t % 7 == 0 and m > 1:
            fs[1] = fs[0]  # duplicated constraint: coincident endpoints quirk
        kind0 = rs.randint(0, 4)
        f0 = (float(abs(rs.randn())) if kind0 == 0 else (0.0 if kind0 in (1, 2) else float(rs.randn())),
              0.0 if kind0 == 2 else float(rs.randn()), float(rs.randn()))
        s = float(abs(rs.randn()) * (0.5 if t % 3 else 3.0))
        out.append(run(f0, fs, s, seed=1000 + t))
    return out


def interval_cases():
    out = []
    cases = [((1, 0, -1, "=="), 0.1)]
    rs = np.random.RandomState(5)
    for _ in range(120):
        kind = rs.randint(0, 5)
        p = rs.randn() * (kind != 0) * (1e-5 if kind == 4 else 1.0)
        cases.append(((float(p), float(rs.randn() * (kind != 1)), float(rs.randn()), "==" if rs.rand() < 0.5 else "<="),
                      float(abs(rs.randn()))))
    for f, s in cases:
        I = u.get_feasible_intervals(u.OneVarQuadraticFunction(*f), s)
        out.append(dict(f=list(f), s=s, intervals=[[float(a), float(b)] for (a, b) in I]))
    return out


def onevar_func_cases():
    out = []
    for gen, gargs in (("random", dict(n=7, m=5, seed=3)), ("circle", dict(ncirc=3)), ("bls", dict(n=6, m=9, seed=1))):
        forms, _ = GEN[gen](**gargs)
        prob = rh.make_form(u, forms)
        rs = np.random.RandomState(17)
        x = rs.randn(prob.n)
        rows = []
        for j, f in enumerate([prob.f0] + prob.fs):
            for k in range(prob.n):
                g = f.get_onevar_func(x, k)
                rows.append([j, k, float(g.P), float(g.q), float(g.r)])
        evals = [float(f.eval(x)) for f in [prob.f0] + prob.fs]
        out.append(dict(gen=gen, gargs=gargs, checksum=checksum(forms), x=lst(x), rows=rows, evals=evals))
    return out


def onecons_and_admm_cases():
    out = dict(onecons=[], admm=[])
    forms, _ = GEN["beam"](n=20, m=5, l=2, seed=1)
    prob = rh.make_form(u, forms)
    np.random.seed(4)
    x0 = 2 * np.random.randn(40)
    for (z, j) in ((0.1 * x0, 0), (x0, 5), (x0, 2), (0.01 * x0, 6)):
        f = prob.fs[j]
        xp = u.onecons_qcqp(np.copy(z), f)
        out["onecons"].append(dict(gen="beam", gargs=dict(n=20, m=5, l=2, seed=1), j=j + 1, z=lst(z), x=lst(xp),
                                   fz=float(f.eval(z)), fx=float(f.eval(xp)), dist2=float(np.sum((xp - z) ** 2))))
    # equality-constrained projections on a random instance
    forms2, _ = GEN["random"](n=8, m=6, seed=11)
    prob2 = rh.make_form(u, forms2)
    rs = np.random.RandomState(2)
    for j in range(prob2.m):
        z = rs.randn(8) * 2
        f = prob2.fs[j]
        xp = u.onecons_qcqp(np.copy(z), f)
        out["onecons"].append(dict(gen="random", gargs=dict(n=8, m=6, seed=11), j=j + 1, z=lst(z), x=lst(xp),
                                   fz=float(f.eval(z)), fx=float(f.eval(xp)), dist2=float(np.sum((xp - z) ** 2))))

    def admm(name, gen, gargs, seed, scale, kwargs):
        forms, _ = GEN[gen](**gargs)
        pr = rh.make_form(u, forms)
        np.random.seed(seed)
        x0 = scale * np.random.randn(pr.n)
        calls = [0]
        orig = u.onecons_qcqp

        def counting(z, f, tol=1e-6):
            calls[0] += 1
            return orig(z, f, tol)
        q.onecons_qcqp = counting
        try:
            x = q.improve_admm(x0, pr, **kwargs)
        finally:
            q.onecons_qcqp = orig
        return dict(name=name, gen=gen, gargs=gargs, checksum=checksum(forms), seed=seed, scale=scale, x0=lst(x0),
                    kwargs=kwargs, x=lst(x), f0=float(pr.f0.eval(x)), maxviol=float(max(pr.violations(x))),
                    onecons_calls=calls[0])

    out["admm"].append(admm("G4", "beam", dict(n=20, m=5, l=2, seed=1), 4, 2.0, dict(rho=float(np.sqrt(7)))))
    out["admm"].append(admm("beam_nophase1", "beam", dict(n=12, m=4, l=2, seed=3), 8, 2.0, dict(rho=2.0, phase1=False, num_iters=200)))
    out["admm"].append(admm("beam_autorho", "beam", dict(n=8, m=3, l=2, seed=2), 5, 1.0, dict(num_iters=150)))
    out["admm"].append(admm("bls_admm", "bls", dict(n=8, m=12, seed=1), 6, 1.0, dict(rho=3.0, num_iters=120)))
    return out


def sdr_cases():
    out = []
    for n, seed in ((10, 2), (24, 9)):
        forms, _ = GEN["bls"](n=n, m=n + 5, seed=1)
        prob = rh.make_form(u, forms)
        Xs = pb.synthetic_sdr_solution(n, rank=4, seed=5)
        mu = np.asarray(Xs[:-1, -1]).flatten()
        import scipy.sparse as sp
        Sigma = np.asmatrix(Xs)[:-1, :-1] - mu * mu.T + 1e-8 * sp.identity(n)   # qcqp.py:394-395 verbatim
        np.random.seed(seed)
        draws = []
        for _ in range(3):
            x = np.random.multivariate_normal(mu, Sigma)
            draws.append(dict(x=lst(x), f0=float(prob.f0.eval(x)), maxviol=float(max(prob.violations(x)))))
        out.append(dict(n=n, gargs=dict(n=n, m=n + 5, seed=1), seed=seed, rank=4, xs_seed=5, draws=draws,
                        Sigma_sum=float(np.asarray(Sigma).sum()), rng=rng_tail()))
    return out


def better_cases():
    out = []
    forms, _ = GEN["random"](n=6, m=4, seed=21)
    prob = rh.make_form(u, forms)
    rs = np.random.RandomState(8)
    for t in range(40):
        x1 = rs.randn(6) * 0.3
        x2 = x1 + rs.randn(6) * (1e-5 if t % 2 else 0.3)
        w = prob.better(x1, x2)
        out.append(dict(x1=lst(x1), x2=lst(x2), pick=1 if w is x1 else 2))
    return dict(gen="random", gargs=dict(n=6, m=4, seed=21), cases=out)


def main():
    G = dict(meta=dict(numpy=np.__version__, note="generated by tests/golden/make_golden.py from /root/reference"))
    cd = []
    cd.append(cd_case("G1", "bls", dict(n=10, m=15, seed=1), "randn", 7, {}))
    cd.append(cd_case("G1_phase1", "bls", dict(n=10, m=15, seed=1), "randn", 7, {}, only_phase=1))
    cd.append(cd_case("G2", "bls", dict(n=10, m=15, seed=1), "sign", 7, dict(phase1=False)))
    cd.append(cd_case("G2p", "bls", dict(n=10, m=15, seed=1), "sign_perturbed", 7, dict(phase1=False)))
    cd.append(cd_case("G2pp", "bls", dict(n=40, m=60, seed=1), "sign_perturbed", 7, dict(phase1=False)))
    cd.append(cd_case("G3", "maxcut", dict(n=25, p=0.2, seed=1), "randn", 11, {}))
    cd.append(cd_case("bls20", "bls", dict(n=20, m=30, seed=1), "randn", 21, {}))
    cd.append(cd_case("bls30_s5", "bls", dict(n=30, m=45, seed=4), "randn", 5, {}))
    cd.append(cd_case("maxcut40", "maxcut", dict(n=40, p=0.15, seed=2), "randn", 12, dict(num_iters=30)))
    cd.append(cd_case("maxcut60", "maxcut", dict(n=60, p=0.1, seed=1), "randn", 13, dict(num_iters=25)))
    cd.append(cd_case("circle3_p1", "circle", dict(ncirc=3), "circle", 31, {}, only_phase=1))
    cd.append(cd_case("circle3", "circle", dict(ncirc=3), "circle", 31, dict(num_iters=12)))
    cd.append(cd_case("circle5", "circle", dict(ncirc=5), "randn", 33, dict(num_iters=6)))
    cd.append(cd_case("beam_cd", "beam", dict(n=6, m=3, l=2, seed=1), "randn_x2", 41, dict(num_iters=15)))
    cd.append(cd_case("random_a", "random", dict(n=6, m=4, seed=0), "randn", 51, dict(num_iters=20)))
    cd.append(cd_case("random_b", "random", dict(n=8, m=3, seed=7, eq_frac=0.0), "randn", 52, dict(num_iters=20)))
    cd.append(cd_case("random_c", "random", dict(n=5, m=6, seed=9, density=0.5), "randn", 53, dict(num_iters=20, viol_tol=0.05)))
    G["cd"] = cd
    G["onevar"] = onevar_cases()
    G["intervals"] = interval_cases()
    G["onevar_func"] = onevar_func_cases()
    G.update(onecons_and_admm_cases())
    G["sdr"] = sdr_cases()
    G["better"] = better_cases()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json")
    with open(path, "w") as fh:
        json.dump(G, fh)
    print("wrote", path, os.path.getsize(path), "bytes")
    for c in cd:
        print("%-12s f0=%.15g viol=%.6g pos=%d untouched=%s" % (c["name"], c["f0"], c["maxviol"], c["rng"]["pos"], c["rng_untouched"]))
    for c in G["admm"]:
        print("%-14s f0=%.15g viol=%.6g calls=%d" % (c["name"], c["f0"], c["maxviol"], c["onecons_calls"]))


if __name__ == "__main__":
    main()
