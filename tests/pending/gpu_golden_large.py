"""NOT COLLECTED (the file name does not match test_*.py): GPU parity tests against tests/golden/golden_large.json -- runs of the
unmodified reference near and at the bench sizes -- written at the end of round 1 when no GPU time was left to run them once.
Next round: run `python -m pytest tests/pending/gpu_golden_large.py -m gpu -q -p no:cacheprovider` on a B200, fix what the first
contact shows, then move the file to tests/test_gpu_golden_large.py.  Until then the chain is: the GPU is held to the oracle at
these sizes (tests/test_gpu_full_size.py, tests/test_gpu_cd.py) and the oracle to this file (tests/test_oracle_golden.py)."""
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from helpers import forms_of, rel_close  # noqa: E402

pytestmark = pytest.mark.gpu


def _large():
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "golden", "golden_large.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("strict", [True, False])
def test_cd_large_goldens(strict):
    """strict: SciPy's summation order, expected bit-level agreement; production mode: the north star's 1e-6 on (f0, maxviol).
    MAXCUT in production mode is left to its properties (exact-zero tests depend on the summation order)."""
    from qcqp_b200 import engine
    for c in _large()["cd"]:
        if c["only_phase"] is not None:
            continue            # the C ABI runs improve_coord_descent as a whole; phase-only goldens pin the oracle
        if not strict and c["gen"] == "maxcut":
            continue
        forms, _ = forms_of(c)
        pack = engine.Pack(forms)
        x0 = np.array(c["x0"])
        rs = np.random.RandomState(c["seed"]); rs.standard_normal(len(x0))
        rng = engine.rng_states(states=[rs.get_state()])
        X, f0, mv, st = pack.cd_improve(x0[None, :], rng, strict=strict, **c["kwargs"])
        assert st[0].status == 0, c["name"]
        rt = 1e-9 if strict else 1e-6
        assert rel_close(f0[0], c["f0"], rtol=rt, atol=rt), (c["name"], f0[0], c["f0"])
        assert rel_close(mv[0], c["maxviol"], rtol=1e-6, atol=1e-9), (c["name"], mv[0], c["maxviol"])
        assert rng[0].pos == c["rng"]["pos"], (c["name"], rng[0].pos)
        pack.close()


@pytest.mark.parametrize("kernel", ["res", "run"])
def test_admm_c4_reference_runs(kernel, monkeypatch):
    from qcqp_b200 import engine, problems as pb
    if kernel == "run":
        monkeypatch.setenv("QCQP_ADMM_KERNEL", "run")
    L = _large()
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "golden", "c4_admm_oracle.json")) as fh:
        x0 = np.array(json.load(fh)["x0"])
    forms, _ = pb.beamforming(n=64, m=24, l=8, seed=1)
    pack = engine.Pack(forms)
    rhos = np.array([c["rho"] for c in L["admm"]])
    X, f0, mv, st = pack.admm_improve(x0[None, :], rhos)
    for k, c in enumerate(L["admm"]):
        assert rel_close(f0[k, 0], c["f0"], rtol=1e-6, atol=1e-9), (c["name"], f0[k, 0], c["f0"])
        assert rel_close(mv[k, 0], c["maxviol"], rtol=1e-6, atol=1e-8), c["name"]
        assert st[k].onecons_calls == c["onecons_calls"], (c["name"], st[k].onecons_calls)
    pack.close()


def test_sdr_sampler_c2_reference_draws():
    from qcqp_b200 import engine, problems as pb
    for c in _large()["sdr"]:
        forms, _ = pb.boolean_least_squares(**c["gargs"])
        pack = engine.Pack(forms)
        mu, _Sigma, F = engine.sdr_factor(pb.synthetic_sdr_solution(c["n"], rank=c["rank"], seed=c["xs_seed"]))
        rs = np.random.RandomState(c["seed"])
        Z = np.stack([rs.standard_normal(c["n"]) for _ in c["draws"]])
        X, f0, mv = pack.sdr_sample_eval(mu, F, Z=Z)
        for i, d in enumerate(c["draws"]):
            assert rel_close(X[i], d["x"], rtol=1e-9, atol=1e-10)
            assert rel_close(f0[i], d["f0"], rtol=1e-9) and rel_close(mv[i], d["maxviol"], rtol=1e-9)
        pack.close()
