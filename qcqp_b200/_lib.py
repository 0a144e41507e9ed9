"""ctypes binding of libqcqp_b200.so (include/qcqp_b200.h).  There is no CPU fallback: if the library is missing the
import of this module fails loudly, and if no CUDA device is visible every compute entry point returns
QCQP_ERR_NO_DEVICE, which surfaces as an Exception."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QCQP_B200_LIB") or os.path.join(_HERE, "libqcqp_b200.so")   # override: an instrumented build (tools/blk_prof.py)

RELOP_NONE, RELOP_LE, RELOP_EQ = 0, 1, 2
RELOP_CODE = {None: RELOP_NONE, "<=": RELOP_LE, "==": RELOP_EQ}

RUN_OK, RUN_EMPTY_MAX, RUN_UNBOUNDED_UNIFORM = 0, 1, 2


class PackDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("m", C.c_int32),
                ("p_ptr", C.c_void_p), ("p_row", C.c_void_p), ("p_col", C.c_void_p), ("p_val", C.c_void_p),
                ("q_ptr", C.c_void_p), ("q_idx", C.c_void_p), ("q_val", C.c_void_p),
                ("r", C.c_void_p), ("relop", C.c_void_p), ("dense_min_fill", C.c_double)]


class PackInfo(C.Structure):
    _fields_ = [("n", C.c_int32), ("m", C.c_int32), ("n_dense", C.c_int32), ("max_incidence", C.c_int32),
                ("incidences", C.c_int64), ("nnz_offdiag", C.c_int64), ("device_bytes", C.c_int64),
                ("bytes_per_sweep_phase2", C.c_double), ("bytes_per_sweep_phase1", C.c_double),
                ("separable", C.c_int32), ("pad_", C.c_int32)]


class RngState(C.Structure):
    """np.random.RandomState state, field for field."""
    _fields_ = [("key", C.c_uint32 * 624), ("pos", C.c_int32), ("has_gauss", C.c_int32), ("gauss", C.c_double)]


class CdParams(C.Structure):
    _fields_ = [("num_iters", C.c_int32), ("viol_tol", C.c_double), ("tol", C.c_double), ("phase1", C.c_int32),
                ("strict", C.c_int32), ("refresh_every", C.c_int32)]


class CdStats(C.Structure):
    _fields_ = [("steps_p1", C.c_int64), ("steps_p2", C.c_int64), ("updates_p1", C.c_int64), ("updates_p2", C.c_int64),
                ("sweeps_p1", C.c_int32), ("sweeps_p2", C.c_int32), ("status", C.c_int32), ("ran_phase2", C.c_int32),
                ("steps_skipped", C.c_int64)]


class AdmmParams(C.Structure):
    _fields_ = [("num_iters", C.c_int32), ("viol_lim", C.c_double), ("tol", C.c_double), ("phase1", C.c_int32)]


class AdmmStats(C.Structure):
    _fields_ = [("iters_p1", C.c_int32), ("iters_p2", C.c_int32), ("onecons_calls", C.c_int64), ("status", C.c_int32),
                ("pad_", C.c_int32)]


EXPORTS = {
    # name: (restype, argtypes)
    "qcqp_last_error": (C.c_char_p, []),
    "qcqp_device_count": (C.c_int, []),
    "qcqp_version": (C.c_char_p, []),
    "qcqp_pack_create": (C.c_int, [C.POINTER(PackDesc), C.POINTER(C.c_void_p)]),
    "qcqp_pack_destroy": (None, [C.c_void_p]),
    "qcqp_pack_get_info": (C.c_int, [C.c_void_p, C.POINTER(PackInfo)]),
    "qcqp_pack_reserve": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "qcqp_eval": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qcqp_eval_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qcqp_cd_improve": (C.c_int, [C.c_void_p, C.POINTER(CdParams), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "qcqp_cd_improve_device": (C.c_int, [C.c_void_p, C.POINTER(CdParams), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qcqp_cd_get_timing": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "qcqp_cd_get_counters": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "qcqp_admm_pack_eig": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qcqp_admm_improve": (C.c_int, [C.c_void_p, C.POINTER(AdmmParams), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qcqp_admm_improve_device": (C.c_int, [C.c_void_p, C.POINTER(AdmmParams), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                           C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qcqp_sdr_sample_eval": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int32, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "qcqp_sdr_sample_eval_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int32, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p]),
    "qcqp_sdr_prefetch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "qcqp_sdr_cd_pipeline": (C.c_int, [C.c_void_p, C.POINTER(CdParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "qcqp_best": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p]),
    "qcqp_best_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qcqp_comm_unique_id": (C.c_int, [C.c_void_p]),
    "qcqp_comm_create": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "qcqp_comm_destroy": (None, [C.c_void_p]),
    "qcqp_best_multi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_int64,
                                  C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p,
                                  C.c_void_p]),
    "qcqp_probe_l2_bandwidth": (C.c_int, [C.c_int64, C.c_int32, C.POINTER(C.c_double)]),
    "qcqp_probe_fp64_peaks": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """Loads libqcqp_b200.so or raises: the engine has no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "or `make -C qcqp_b200/csrc` (nvcc, sm_100a). qcqp_b200 has no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(L, name)     # AttributeError if the library does not export what the header declares
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = load().qcqp_last_error()
        raise Exception((msg.decode() if msg else "qcqp_b200 error") + " [status %d]" % rc)
