"""CPU-side checks of the product: the C-ABI library loads and exports every symbol include/qcqp_b200.h declares, fails
loudly without a device, the host logic (form flattening, sharding, best-pick order) is right, and the N>1 path works
over gloo with world_size 2."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from qcqp_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "qcqp_b200.h")).read()
    declared = set(re.findall(r"\b(qcqp_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), "libqcqp_b200.so does not export %s" % name
    assert set(_lib.EXPORTS) == declared, (set(_lib.EXPORTS) ^ declared)
    assert b"sm_100a" in _lib.load().qcqp_version()


def test_struct_layouts_match_header():
    from qcqp_b200 import _lib
    assert C.sizeof(_lib.RngState) == 624 * 4 + 16
    assert C.sizeof(_lib.CdStats) == 56 and C.sizeof(_lib.CdParams) == 40 and C.sizeof(_lib.AdmmStats) == 24


def test_no_device_fails_loudly():
    """No CPU fallback: without a GPU every entry point reports QCQP_ERR_NO_DEVICE instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from qcqp_b200 import engine, problems as pb
    forms, _ = pb.boolean_least_squares(6, 9)
    with pytest.raises(Exception, match="no CUDA device"):
        engine.Pack(forms)
    with pytest.raises(Exception, match="no CUDA device"):
        engine.best([1.0], [0.0])


def test_c_abi_rejects_bad_arguments():
    """Error convention of include/qcqp_b200.h: negative status, no exception across the ABI, message through
    qcqp_last_error.  Argument checks come before any device work, so they can be exercised without a GPU."""
    from qcqp_b200 import _lib
    L = _lib.load()
    INVALID = -1
    x = (C.c_double * 4)()
    out = C.c_int32(7)
    prm = _lib.CdParams(10, 1e-2, 1e-4, 1, 0, 0)
    calls = {
        "qcqp_eval": lambda: L.qcqp_eval(None, x, 1, x, x, None),
        "qcqp_cd_improve": lambda: L.qcqp_cd_improve(None, C.byref(prm), x, 1, x, x, x, x, None),
        "qcqp_sdr_sample_eval": lambda: L.qcqp_sdr_sample_eval(None, x, x, None, 0, 1, x, x, x),
        "qcqp_best": lambda: L.qcqp_best(x, x, 0, 1e-4, C.byref(out)),
        "qcqp_best(null)": lambda: L.qcqp_best(None, x, 4, 1e-4, C.byref(out)),
        "qcqp_best(tol)": lambda: L.qcqp_best(x, x, 4, 0.0, C.byref(out)),
        "qcqp_pack_create": lambda: L.qcqp_pack_create(None, C.byref(C.c_void_p())),
        "qcqp_pack_get_info": lambda: L.qcqp_pack_get_info(None, C.byref(_lib.PackInfo())),
    }
    for name, call in calls.items():
        rc = call()
        assert rc == INVALID, (name, rc, L.qcqp_last_error())
        assert L.qcqp_last_error(), name
    assert out.value == 7                       # outputs untouched on failure
    L.qcqp_pack_destroy(None)                   # destroying a null handle is a no-op
    with pytest.raises(Exception, match=r"\[status -1\]"):
        _lib.check(L.qcqp_best(x, x, 0, 1e-4, C.byref(out)))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "qcqp_b200")
    for base, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert "qcqp_oracle" not in txt and "libqcqp_oracle" not in txt, f


def test_flatten_forms_layout():
    from qcqp_b200 import engine, problems as pb
    forms, _ = pb.circle_packing(3)
    f = engine.flatten_forms(forms)
    assert f["n"] == 7 and f["m"] == len(forms) - 1
    assert f["p_ptr"][0] == 0 and f["p_ptr"][-1] == len(f["p_val"]) and np.all(np.diff(f["p_ptr"]) >= 0)
    for j in range(f["m"] + 1):
        a, b = f["p_ptr"][j], f["p_ptr"][j + 1]
        key = f["p_row"][a:b].astype(np.int64) * f["n"] + f["p_col"][a:b]
        assert np.all(np.diff(key) > 0)
        qa, qb = f["q_ptr"][j], f["q_ptr"][j + 1]
        dense_q = np.zeros(f["n"]); dense_q[f["q_idx"][qa:qb]] = f["q_val"][qa:qb]
        assert np.array_equal(dense_q, np.asarray(forms[j][1], dtype=float))
    assert not np.any(f["p_val"] == 0.0)


def test_rng_state_roundtrip():
    from qcqp_b200 import engine
    rs = np.random.RandomState(5); rs.standard_normal(3)
    arr = engine.rng_states(states=[rs.get_state()])
    back = engine.rng_state_tuple(arr[0])
    rs2 = np.random.RandomState(0); rs2.set_state(back)
    assert rs2.uniform() == rs.uniform()


def test_shard_ranges_cover_everything():
    from qcqp_b200.dist import shard_range
    for total in (0, 1, 7, 256, 1024, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_local_best_is_the_better_fold():
    """local_best == folding `best = better(best, x_r)` over r with QCQPForm.better (utilities.py:135-146), which returns
    its SECOND argument on an exact tie -- hence "later index wins"."""
    from qcqp_b200.dist import local_best

    def better_returns_first(mv1, f1, mv2, f2, tol=1e-4):
        v1, v2 = int(mv1 / tol), int(mv2 / tol)
        if v1 < v2: return True
        if v2 < v1: return False
        return f1 < f2

    rs = np.random.RandomState(3)
    for t in range(200):
        R = int(rs.randint(1, 40))
        f0 = np.round(rs.randn(R), 1)
        mv = np.abs(rs.randn(R)) * 10.0 ** rs.randint(-6, 0, size=R)
        best = 0
        for r in range(1, R):
            if not better_returns_first(mv[best], f0[best], mv[r], f0[r]):
                best = r
        assert local_best(f0, mv)[2] == best


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from qcqp_b200.dist import shard_range, local_best, global_best, broadcast_point
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
rs = np.random.RandomState(0)
R, n = 37, 5
f0 = np.round(rs.randn(R), 1); mv = np.abs(rs.randn(R)) * 1e-3; X = rs.randn(R, n)
lo, hi = shard_range(R, rank, 2)
b, f, i = local_best(f0[lo:hi], mv[lo:hi])
gb, gf, gi = global_best(b, f, lo + i if i >= 0 else -1)
want = local_best(f0, mv)
assert (gb, gf, gi) == want, ((gb, gf, gi), want)
owner = 0 if gi < shard_range(R, 0, 2)[1] else 1
x = broadcast_point(X[gi] if rank == owner else np.zeros(n), owner, n)
assert np.array_equal(x, X[gi])
# the same pick with the winner's point riding on the single all-gather
gb2, gf2, gi2, xw = global_best(b, f, lo + i if i >= 0 else -1, x=X[lo + i])
assert (gb2, gf2, gi2) == want and np.array_equal(xw, X[gi])
# a rank with nothing to offer (empty shard / every restart failed) cannot win and does not hang the others
e = global_best(np.iinfo(np.int64).max, np.inf, -1, x=np.zeros(n)) if rank == 1 else global_best(b, f, lo + i, x=X[lo + i])
assert e[2] == shard_range(R, 0, 2)[0] + local_best(f0[:shard_range(R, 0, 2)[1]], mv[:shard_range(R, 0, 2)[1]])[2]
dist.barrier(); dist.destroy_process_group()
print("OK", rank)
'''


def test_gloo_world_size_2_best_pick(tmp_path):
    """The N>1 path (sharded restarts + best-pick reduction + winner broadcast) on two CPU ranks over gloo."""
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + (os.getpid() % 2000))
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "OK %d" % r in o, o


_GLOO_FACADE_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
import qcqp_b200 as Q
import qcqp_b200.model as cvx
from qcqp_b200 import engine, problems as pb
from qcqp_b200.dist import local_best
from helpers import OraclePack
engine.Pack = OraclePack                                   # no GPU here: the facade's host logic over the CPU oracle
engine.best = lambda f0, mv, tol=1e-4: local_best(f0, mv, tol)[2]
world = int(sys.argv[4])

def build():
    np.random.seed(1)
    A = np.random.randn(18, 12); b = np.random.randn(18, 1)
    x = cvx.Variable(12)
    return x, Q.QCQP(cvx.Problem(cvx.Minimize(cvx.sum_squares(A*x - b)), [cvx.square(x) == 1]))

def flow(qc, S):
    out = []
    np.random.seed(5)
    out.append(qc.suggest(Q.RANDOM, samples=S))
    out.append(qc.improve(Q.COORD_DESCENT, seed=100, num_iters=40))
    out.append((float(qc.best_index), float(qc.x.sum())))
    qc.set_sdr_solution(pb.synthetic_sdr_solution(12, rank=3, seed=5))
    np.random.seed(6)
    out.append(qc.suggest(Q.SDR, samples=S))
    out.append(qc.improve(Q.COORD_DESCENT, seed=7, num_iters=40))
    np.random.seed(8)
    out.append(qc.suggest_improve(samples=S, seed=300, num_iters=40))
    out.append((float(qc.best_index), float(qc.x.sum())))
    return np.array(out)

# single process first: the whole batch on one rank
x1, q1 = build()
want = {S: flow(q1, S) for S in (7, 2, 1)}
x_want = x1.value.copy()
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=world)
x2, q2 = build()
for S in (7, 2, 1):                                        # S = 2 < world = 3 leaves a rank with an empty shard; S = 1 is replicated
    got = flow(q2, S)
    assert np.array_equal(got, want[S]), (S, got, want[S])
    if S > 1:
        lo, hi, total = q2._shard
        assert total == S and q2.X.shape[0] == hi - lo
assert np.array_equal(x2.value, x_want)                    # every rank ends with the same written-back point
dist.barrier(); dist.destroy_process_group()
print("OK", sys.argv[3])
'''


def test_gloo_sharded_facade_matches_single_process(tmp_path):
    """SURVEY 8e at the facade: under a process group a batch of restarts / draws is sharded over the ranks and the best point
    is picked by one reduction; objective, violation, winner index and point must equal the single-process run bit for bit,
    whatever the number of ranks (restart r always owns the stream seed + r).  Three gloo ranks, the oracle as the pack."""
    script = tmp_path / "wf.py"
    script.write_text(_GLOO_FACADE_WORKER)
    port = str(31600 + (os.getpid() % 2000))
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r), "3"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(3)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "OK %d" % r in o, o


def test_host_relaxations():
    """qcqp_b200/relax.py (host NumPy SDP, setup code): the unit-diagonal mixing method agrees with the general ADMM solver,
    solutions are PSD and feasible, and the values are valid bounds."""
    import itertools
    from qcqp_b200 import relax, problems as pb
    from qcqp_b200.forms import QCQPForm
    forms, _ = pb.maxcut(12, 0.4, seed=3)
    F = QCQPForm.from_tuples(forms)
    X1, v1 = relax.solve_sdr(F)
    N = 13
    E = np.zeros((N, N)); E[-1, -1] = 1
    A_eq = [E] + [relax.homogeneous_form(f) for f in F.fs]
    X2, v2, info = relax._admm_sdp(relax.homogeneous_form(F.f0), A_eq, [1.0] + [0.0] * 12, [], [])
    assert abs(v1 - v2) < 1e-4 * abs(v1)
    assert np.linalg.eigvalsh(X1).min() > -1e-9 and np.abs(np.diag(X1) - 1).max() < 1e-9
    forms, _ = pb.boolean_least_squares(8, 12, seed=2)
    F = QCQPForm.from_tuples(forms)
    X, v = relax.solve_sdr(F)
    P0 = np.asarray(forms[0][0].todense()); q0 = forms[0][1]; r0 = forms[0][2]
    best = min(np.array(s) @ P0 @ np.array(s) + q0 @ np.array(s) + r0 for s in itertools.product([-1.0, 1.0], repeat=8))
    assert v <= best + 1e-6
    xs, vs = relax.solve_spectral(F)
    assert vs <= v + 1e-5
    forms, _ = pb.beamforming(6, 3, 2, seed=1)
    F = QCQPForm.from_tuples(forms)
    X, v = relax.solve_sdr(F)
    assert np.linalg.eigvalsh(X).min() > -1e-7 and abs(X[-1, -1] - 1) < 1e-6
    assert max(np.sum(relax.homogeneous_form(f) * X) for f in F.fs) < 1e-4


def test_integration_stub_is_executable_against_the_reference():
    """INTEGRATION.md's `qcqp/b200.py` (the binding a reference maintainer would add) is run as printed against the unmodified
    reference's QCQPForm: it must flatten the forms and reach qcqp_pack_create, which -- with no GPU in the build container --
    answers QCQP_ERR_NO_DEVICE.  Build container only (the reference tree does not travel)."""
    import torch
    from oracle import ref_harness as rh
    from qcqp_b200 import _lib, problems as pb
    if not rh.available() or torch.cuda.is_available():
        pytest.skip("needs the reference tree and no GPU")
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"```python\n(# qcqp/b200.py.*?)```", md, re.S).group(1)
    code = code.replace('C.CDLL("libqcqp_b200.so")', "C.CDLL(%r)" % _lib.LIB_PATH)
    ns = {}
    exec(code, ns)
    u, _q = rh.load()
    for forms in (pb.boolean_least_squares(10, 15, seed=1)[0], pb.circle_packing(4)[0]):
        prob = rh.make_form(u, forms)
        with pytest.raises(Exception, match="no CUDA device"):
            ns["improve_coord_descent"](np.random.RandomState(0).randn(prob.n), prob)


def test_reference_window_runner_and_clean_cpu_arm(tmp_path):
    """bench.py's CPU arms: (i) oracle/ref_python.py times windows of the UNMODIFIED reference's coord_descent_phase2 (here from
    /root/reference or baseline/_ref, whichever exists) and counts exactly the coordinate steps asked for; (ii) building a
    benchmark instance and running the C port does not import qcqp_b200 nor map libqcqp_b200.so (the reference arm's record
    must be free of the product library)."""
    from oracle import ref_python as rp
    if rp.ref_root() is None:
        pytest.skip("reference package not present (neither baseline/_ref nor /root/reference)")
    pool = rp.ReferencePool("boolean_least_squares", dict(n=40, m=60, seed=1), procs=2)
    steps, wall, secs = pool.window(5)
    pool.close()
    assert steps == 10 and wall > 0 and len(secs) == 2
    code = r'''
import sys
sys.path.insert(0, %r)
import bench
cfg = dict(bench.CONFIGS["c2"]); cfg["gargs"] = dict(n=24, m=36, seed=1)
forms, info, Xstar = bench.build_problem(cfg)
wk, dt = bench.port_sample(cfg, forms, Xstar, 8, 2)
assert wk > 0
assert not any(m == "qcqp_b200" or m.startswith("qcqp_b200.") for m in sys.modules), [m for m in sys.modules if "qcqp" in m]
assert "libqcqp_b200" not in open("/proc/self/maps").read()
print("CLEAN")
''' % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "CLEAN" in out.stdout, out.stdout + out.stderr


def test_bench_strong_scaling_shards_partition_one_batch():
    """bench.py c3 / c5: the per-rank inputs of a strong-scaling run are slices of ONE batch (same draws, same seeds whatever N)."""
    import bench
    cfg = dict(bench.CONFIGS["c3"]); cfg["gargs"] = dict(n=12, p=0.3, seed=1); cfg["restarts"] = 10
    forms, _i, _x = bench.build_problem(cfg, want_sdr=False)
    whole, seeds = bench.starts_for(cfg, forms, None, 0, 10, 0, False)
    parts = [bench.starts_for(cfg, forms, None, *bench.shard(10, r, 4), r, False) for r in range(4)]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), whole) and np.array_equal(np.concatenate([p[1] for p in parts]), seeds)
    cfg5 = dict(bench.CONFIGS["c5"]); cfg5["gargs"] = dict(ncirc=3); cfg5["restarts"] = 6
    forms5, _i, _x = bench.build_problem(cfg5)
    w5, s5 = bench.starts_for(cfg5, forms5, None, 0, 6, 0, False)
    p5 = [bench.starts_for(cfg5, forms5, None, *bench.shard(6, r, 4), r, False) for r in range(4)]
    assert np.array_equal(np.concatenate([p[0] for p in p5]), w5) and np.array_equal(w5[2], np.random.RandomState(2).randn(7))
