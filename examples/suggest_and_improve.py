#!/usr/bin/env python
"""The four problem families of the reference's examples/ directory, modelled with qcqp_b200.model (the cvxpy surface those
scripts use) and run through QCQP.suggest / QCQP.improve on the GPU.

    python examples/suggest_and_improve.py bls    --n 100 --m 150 --samples 256
    python examples/suggest_and_improve.py maxcut --n 200 --p 0.1 --samples 256
    python examples/suggest_and_improve.py beam   --n 20 --m 5 --l 2
    python examples/suggest_and_improve.py circle --n 10 --samples 64

`--samples S` runs S SDR draws (or S random starts for `circle`) as ONE batch and keeps the best point in the reference's
`better` order; S = 1 consumes the process-global np.random stream exactly as the reference does.  The DCCP / IPOPT lines of
the reference scripts are third-party wrappers and are not reproduced (QCQP.improve raises "... package is not installed.").
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import qcqp_b200.model as cvx                                                  # noqa: E402
from qcqp_b200 import QCQP, RANDOM, SDR, COORD_DESCENT, ADMM                   # noqa: E402


def bls(args):
    """minimize ||Ax - b||^2 over x in {-1, +1}^n  (reference: examples/boolean_least_squares.py)."""
    A = np.random.randn(args.m, args.n)
    b = np.random.randn(args.m, 1)
    x = cvx.Variable(args.n)
    prob = cvx.Problem(cvx.Minimize(cvx.sum_squares(A*x - b)), [cvx.square(x) == 1])
    return prob, [x], SDR, [[COORD_DESCENT], [COORD_DESCENT, ADMM]]


def maxcut(args):
    """maximize the cut weight of a random graph  (reference: examples/maxcut.py)."""
    U = np.triu(np.random.uniform(size=(args.n, args.n)), 1)
    W = ((U + U.T + np.eye(args.n)) < args.p).astype(float)
    x = cvx.Variable(args.n)
    prob = cvx.Problem(cvx.Maximize(0.25*(cvx.sum_entries(W) - cvx.quad_form(x, W))), [cvx.square(x) == 1])
    return prob, [x], SDR, [[COORD_DESCENT]]


def beam(args):
    """secondary-user multicast beamforming, real form  (reference: examples/secondary_user_beamforming.py)."""
    n, m, l = args.n, args.m, args.l
    HR, HI = np.random.randn(m, n), np.random.randn(m, n)
    GR, GI = np.random.randn(l, n), np.random.randn(l, n)
    A, B = np.hstack((HR, HI)), np.hstack((-HI, HR))
    C, D = np.hstack((GR, GI)), np.hstack((-GI, GR))
    x = cvx.Variable(2*n)
    cons = [cvx.square(A*x) + cvx.square(B*x) >= args.tau, cvx.square(C*x) + cvx.square(D*x) <= args.eta]
    prob = cvx.Problem(cvx.Minimize(cvx.sum_squares(x)), cons)
    return prob, [x], SDR, [[COORD_DESCENT], [ADMM], [COORD_DESCENT]]


def circle(args):
    """largest common radius of n circles in a box  (reference: examples/circle_packing.py)."""
    X = cvx.Variable(2, args.n)
    r = cvx.Variable()
    side = 10
    cons = [X >= r, X <= side - r, r >= 0]
    for i in range(args.n):
        for j in range(i + 1, args.n):
            cons.append(cvx.square(2*r) <= cvx.sum_squares(X[:, i] - X[:, j]))
    return cvx.Problem(cvx.Maximize(r), cons), [r, X], RANDOM, [[COORD_DESCENT]]


FAMILIES = {"bls": bls, "maxcut": maxcut, "beam": beam, "circle": circle}


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("family", choices=sorted(FAMILIES))
    ap.add_argument("--n", type=int, default=10)
    ap.add_argument("--m", type=int, default=15)
    ap.add_argument("--l", type=int, default=2)
    ap.add_argument("--p", type=float, default=0.2)
    ap.add_argument("--tau", type=float, default=20.0)
    ap.add_argument("--eta", type=float, default=2.0)
    ap.add_argument("--samples", type=int, default=1)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--num-iters", type=int, default=1000)
    args = ap.parse_args()

    np.random.seed(args.seed)
    prob, variables, suggest, stages = FAMILIES[args.family](args)
    qcqp = QCQP(prob)
    print("%s: n = %d variables, m = %d constraints" % (args.family, qcqp.n, qcqp.qcqp_form.m))

    f, v = qcqp.suggest(suggest, samples=args.samples)
    if suggest == SDR:
        print("SDR bound: %.3f" % qcqp.sdr_bound)
    print("suggest(%s): objective %.3f, violation %.3g" % (suggest, f, v))
    for stage in stages:
        kw = dict(num_iters=args.num_iters)
        if args.samples > 1:
            kw["seed"] = args.seed
        if stage == [ADMM] and args.family == "beam":
            kw["rho"] = np.sqrt(args.m + args.l)
        f, v = qcqp.improve(stage, **kw)
        print("improve(%s): objective %.3f, violation %.3g" % (" + ".join(stage), f, v))
    for var in variables:
        print("%s =\n%s" % (var.name, np.array2string(np.asarray(var.value), precision=4, threshold=40)))


if __name__ == "__main__":
    main()
