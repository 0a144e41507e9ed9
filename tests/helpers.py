"""Shared helpers for the parity tests."""
import numpy as np

from qcqp_b200 import problems as pb

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


def forms_of(case):
    forms, info = GEN[case["gen"]](**case["gargs"])
    if "checksum" in case:
        assert abs(checksum(forms) - case["checksum"]) <= 1e-12 * max(1.0, abs(case["checksum"])), "generator drifted from the golden file"
    return forms, info


def rel_close(a, b, rtol=1e-6, atol=1e-9):
    a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
    return np.all(np.abs(a - b) <= atol + rtol * np.maximum(np.abs(a), np.abs(b)))
