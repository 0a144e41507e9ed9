#!/usr/bin/env python
"""Mints tests/golden/golden_large.json: coordinate-descent runs of the UNMODIFIED reference at sizes nearer the bench
configuration (Boolean LS n = 100 and 150, MAXCUT n = 120, circle packing 8 circles), where one run of the Python reference
takes from tens of seconds to minutes.  Same recipe and record layout as make_golden.py's `cd` section.  Build container only:

    python tests/golden/make_golden_large.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402  (loads the reference behind the cvxpy stub)


def admm_c4(rho_index):
    """BASELINE configuration C4 (beamforming n=64 antennas: N=128, 32 constraints), x0 = 2 randn(128) after seed(4), one rho of
    the sweep rho = sqrt(32) 2^(k/2), k = -8..7 -- the start point and rho values of tests/golden/c4_admm_oracle.json."""
    import numpy as np
    u, q, rh = mg.u, mg.q, mg.rh
    forms, _ = mg.GEN["beam"](n=64, m=24, l=8, seed=1)
    pr = rh.make_form(u, forms)
    rho = float(np.sqrt(32) * 2.0 ** ((rho_index - 8) / 2.0))
    np.random.seed(4)
    x0 = 2 * np.random.randn(pr.n)
    calls = [0]
    orig = u.onecons_qcqp

    def counting(z, f, tol=1e-6):
        calls[0] += 1
        return orig(z, f, tol)
    q.onecons_qcqp = counting
    t = time.time()
    try:
        x = q.improve_admm(x0, pr, rho=rho)
    finally:
        q.onecons_qcqp = orig
    c = dict(name="c4_rho%d" % rho_index, rho_index=rho_index, rho=rho, checksum=mg.checksum(forms), x=mg.lst(x),
             f0=float(pr.f0.eval(x)), maxviol=float(max(pr.violations(x))), onecons_calls=calls[0], reference_seconds=time.time() - t)
    print("%-12s rho=%.4f f0=%.15g viol=%.6g calls=%d  (%.0f s of the Python reference)" % (c["name"], rho, c["f0"], c["maxviol"], calls[0], c["reference_seconds"]), flush=True)
    return c


def sdr_c2():
    """The sampler lines of QCQP.suggest (qcqp.py:394-401) at the C2 size: n = 1000, the declared synthetic X* of the GPU tests
    (rank 16, seed 5), two draws after np.random.seed(2)."""
    import numpy as np
    import scipy.sparse as sp
    n = 1000
    forms, _ = mg.GEN["bls"](n=n, m=1500, seed=1)
    prob = mg.rh.make_form(mg.u, forms)
    Xs = mg.pb.synthetic_sdr_solution(n, rank=16, seed=5)
    mu = np.asarray(Xs[:-1, -1]).flatten()
    Sigma = np.asmatrix(Xs)[:-1, :-1] - mu * mu.T + 1e-8 * sp.identity(n)   # qcqp.py:394-395 verbatim
    np.random.seed(2)
    draws = []
    t = time.time()
    for _ in range(2):
        x = np.random.multivariate_normal(mu, Sigma)
        draws.append(dict(x=mg.lst(x), f0=float(prob.f0.eval(x)), maxviol=float(max(prob.violations(x)))))
    print("sdr_c2: 2 draws in %.1f s of the Python reference, f0 = %.12g, %.12g" % (time.time() - t, draws[0]["f0"], draws[1]["f0"]), flush=True)
    return [dict(n=n, gargs=dict(n=n, m=1500, seed=1), seed=2, rank=16, xs_seed=5, draws=draws,
                 Sigma_sum=float(np.asarray(Sigma).sum()), rng=mg.rng_tail())]


def main():
    cases = [
        ("bls100", "bls", dict(n=100, m=150, seed=1), "randn", 61, {}),
        ("bls150_s3", "bls", dict(n=150, m=225, seed=3), "randn", 62, {}),
        ("maxcut120", "maxcut", dict(n=120, p=0.1, seed=2), "randn", 63, dict(num_iters=40)),
        ("circle8", "circle", dict(ncirc=8), "randn", 64, dict(num_iters=5)),
        # the bench configuration C2 itself: phase 1 and two phase-2 sweeps of one restart (~2 min per sweep in the reference)
        ("bls1000_2sweeps", "bls", dict(n=1000, m=1500, seed=1), "randn", 71, dict(num_iters=2)),
        # ... and one restart of it run to convergence (num_iters = 1000: about 20 phase-2 sweeps)
        ("bls1000_full", "bls", dict(n=1000, m=1500, seed=1), "randn", 74, {}),
        # C3's instance (MAXCUT n = 2000, p = 0.1, CSR objective): the phase-1 sweep of one restart (~10 min in the reference)
        ("maxcut2000_p1", "maxcut", dict(n=2000, p=0.1, seed=1), "randn", 72, dict(num_iters=1), 1),
        # ... and phase 1 followed by one phase-2 sweep (the exact-zero tests of SURVEY H5 at n = 2000)
        ("maxcut2000_1sweep_each", "maxcut", dict(n=2000, p=0.1, seed=1), "randn", 72, dict(num_iters=1)),
        # (C4's instance under COORD_DESCENT is not minted: the reference's own result is chaotic there -- a 3e-16 relative
        #  perturbation of x0 moves its one-sweep output by 0.69 and its stream position from 442 to 416; DESIGN.md "Conditioning")
        # C5's instance (circle packing, 200 circles: N = 401, 20 701 constraints): the first phase-1 sweep of one restart
        ("circle200_p1", "circle", dict(ncirc=200), "randn", 73, dict(num_iters=1), 1),
    ]
    only = sys.argv[1:]            # names to (re)mint; everything else is kept from the existing file
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_large.json")
    old = json.load(open(path)) if (only and os.path.exists(path)) else dict(cd=[], admm=[], sdr=[])
    out = [c for c in old["cd"] if c["name"] not in only]
    for name, gen, gargs, recipe, seed, kw, *phase in cases:
        if only and name not in only:
            continue
        t = time.time()
        c = mg.cd_case(name, gen, gargs, recipe, seed, kw, only_phase=(phase[0] if phase else None))
        c["reference_seconds"] = time.time() - t
        out.append(c)
        print("%-12s f0=%.15g viol=%.6g pos=%d  (%.0f s of the Python reference)" % (name, c["f0"], c["maxviol"], c["rng"]["pos"], c["reference_seconds"]), flush=True)
    import numpy as np
    admm = old["admm"] if only else [admm_c4(k) for k in (4, 8, 11)]
    sdr = old.get("sdr") if (only and "sdr_c2" not in only and old.get("sdr")) else sdr_c2()
    with open(path, "w") as fh:
        json.dump(dict(meta=dict(numpy=np.__version__, note="generated by tests/golden/make_golden_large.py from /root/reference"), cd=out, admm=admm, sdr=sdr), fh)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
