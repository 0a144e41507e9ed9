"""Host-side handle on the B200 engine: builds a pack from quadratic forms and calls the C ABI.

``forms = [(P, q, r, relop), ...]`` is the reference's QCQPForm content (utilities.py:318-347): forms[0] the
objective in minimise form (relop None), then the scalar constraints with relop '<=' or '=='."""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _lib
from ._lib import AdmmParams, AdmmStats, CdParams, CdStats, PackDesc, PackInfo, RngState, RELOP_CODE, check


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def flatten_forms(forms):
    """[(P, q, r, relop)] -> the arrays of qcqp_pack_desc (COO per form sorted by (row, col), sparse q)."""
    n = int(np.asarray(forms[0][1]).size)
    p_ptr = [0]; q_ptr = [0]
    rows, cols, vals, qi, qv = [], [], [], [], []
    for (P, q, _r, _op) in forms:
        M = sp.csr_matrix(P, dtype=np.float64)
        if M.shape != (n, n):
            raise Exception("quadratic form has shape %s, expected (%d, %d)" % (M.shape, n, n))
        M.sum_duplicates(); M.eliminate_zeros(); M.sort_indices()
        coo = M.tocoo()
        rows.append(coo.row.astype(np.int32)); cols.append(coo.col.astype(np.int32)); vals.append(coo.data.astype(np.float64))
        p_ptr.append(p_ptr[-1] + coo.nnz)
        qa = np.asarray(q.todense() if sp.issparse(q) else q, dtype=np.float64).ravel()
        nz = np.flatnonzero(qa)
        qi.append(nz.astype(np.int32)); qv.append(qa[nz])
        q_ptr.append(q_ptr[-1] + len(nz))
    cat = lambda xs, dt: np.ascontiguousarray(np.concatenate(xs) if xs else np.zeros(0), dtype=dt)
    return dict(
        n=n, m=len(forms) - 1,
        p_ptr=np.ascontiguousarray(p_ptr, dtype=np.int64), p_row=cat(rows, np.int32), p_col=cat(cols, np.int32), p_val=cat(vals, np.float64),
        q_ptr=np.ascontiguousarray(q_ptr, dtype=np.int64), q_idx=cat(qi, np.int32), q_val=cat(qv, np.float64),
        r=np.ascontiguousarray([float(f[2]) for f in forms], dtype=np.float64),
        relop=np.ascontiguousarray([RELOP_CODE[f[3]] for f in forms], dtype=np.int32),
    )


def rng_states(seeds=None, states=None):
    """An array of qcqp_rng_state: one MT19937 stream per restart, seeded like np.random.seed(seed_r)."""
    if states is not None:
        src = states
    else:
        src = [np.random.RandomState(int(s)).get_state() for s in seeds]
    arr = (RngState * len(src))()
    for i, st in enumerate(src):
        _name, key, pos, has_gauss, gauss = st
        C.memmove(arr[i].key, np.ascontiguousarray(key, dtype=np.uint32).ctypes.data, 624 * 4)
        arr[i].pos, arr[i].has_gauss, arr[i].gauss = int(pos), int(has_gauss), float(gauss)
    return arr


def rng_state_tuple(st):
    return ("MT19937", np.frombuffer(st.key, dtype=np.uint32).copy(), int(st.pos), int(st.has_gauss), float(st.gauss))


def rng_states_as_tensor_bytes(arr):
    """The raw bytes of a qcqp_rng_state array as a uint8 numpy array (for staging into a device tensor)."""
    return np.frombuffer(arr, dtype=np.uint8).copy()


class Pack:
    """The stacked (P_j, q_j, r_j, relop_j) resident in HBM; mirror of the reference's QCQPForm."""

    def __init__(self, forms, dense_min_fill=0.0):
        L = _lib.load()
        self.forms = forms
        self._flat = flatten_forms(forms)
        f = self._flat
        self.n, self.m = f["n"], f["m"]
        desc = PackDesc(f["n"], f["m"], _ptr(f["p_ptr"]).value, _ptr(f["p_row"]).value, _ptr(f["p_col"]).value, _ptr(f["p_val"]).value,
                        _ptr(f["q_ptr"]).value, _ptr(f["q_idx"]).value, _ptr(f["q_val"]).value, _ptr(f["r"]).value,
                        _ptr(f["relop"]).value, float(dense_min_fill))
        h = C.c_void_p()
        check(L.qcqp_pack_create(C.byref(desc), C.byref(h)))
        self._h = h
        info = PackInfo()
        check(L.qcqp_pack_get_info(self._h, C.byref(info)))
        self.info = info
        self._has_eig = False

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().qcqp_pack_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    # ---- evaluation ------------------------------------------------------------------------------------
    def eval(self, X, want_viol=False):
        X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, self.n)
        R = X.shape[0]
        f0 = np.empty(R); mv = np.empty(R)
        viol = np.empty((R, self.m)) if want_viol else None
        check(_lib.load().qcqp_eval(self._h, _ptr(X), R, _ptr(f0), _ptr(mv), _ptr(viol) if want_viol else None))
        return (f0, mv, viol) if want_viol else (f0, mv)

    # ---- coordinate descent ----------------------------------------------------------------------------
    def cd_improve(self, X0, rng, num_iters=1000, viol_tol=1e-2, tol=1e-4, phase1=True, strict=False, refresh_every=0):
        """improve_coord_descent for each row of X0 (qcqp.py:181-192). rng: array of RngState, advanced in place.
        strict: 0/False fast (cached dense row dots), 1/True bit-exact sequential dots, 2 fresh parallel dots.
        Returns (X, f0, maxviol, stats)."""
        X0 = np.ascontiguousarray(X0, dtype=np.float64).reshape(-1, self.n)
        R = X0.shape[0]
        if len(rng) != R:
            raise Exception("need one RNG state per restart")
        X = np.empty_like(X0); f0 = np.empty(R); mv = np.empty(R)
        stats = (CdStats * R)()
        prm = CdParams(int(num_iters), float(viol_tol), float(tol), int(bool(phase1)), int(strict), int(refresh_every))
        check(_lib.load().qcqp_cd_improve(self._h, C.byref(prm), _ptr(X0), R, C.cast(rng, C.c_void_p), _ptr(X), _ptr(f0), _ptr(mv),
                                          C.cast(stats, C.c_void_p)))
        return X, f0, mv, stats

    # ---- ADMM ------------------------------------------------------------------------------------------
    def set_eig(self, lmb, Q, qhat):
        lmb = np.ascontiguousarray(lmb, dtype=np.float64); Q = np.ascontiguousarray(Q, dtype=np.float64)
        qhat = np.ascontiguousarray(qhat, dtype=np.float64)
        check(_lib.load().qcqp_admm_pack_eig(self._h, _ptr(lmb), _ptr(Q), _ptr(qhat)))
        self._has_eig = True

    def compute_eig(self):
        """Host-side setup the reference caches on f.eigh (utilities.py:160-166), with the same NumPy calls, so the
        device works from bit-identical (lambda, Q, Q^T q)."""
        n, m = self.n, self.m
        lmb = np.empty((m, n)); Q = np.empty((m, n, n)); qhat = np.empty((m, n))
        for i in range(m):
            P = sp.csr_matrix(self.forms[i + 1][0])
            Psymm = (P + P.T) / 2.
            lmb[i], Q[i] = np.linalg.eigh(np.asarray(Psymm.todense()))
            q = self.forms[i + 1][1]
            qa = np.asarray(q.todense() if sp.issparse(q) else q, dtype=np.float64).ravel()
            qhat[i] = Q[i].T.dot(qa)
        self.set_eig(lmb, Q, qhat)

    def compute_eig_device(self):
        """SURVEY 8(f) f-3, opt-in: the m eigendecompositions of utilities.py:160-162 as ONE batched `eigh` on the GPU (torch ->
        cuSOLVER; a library call, setup code outside every timed region) instead of m LAPACK calls on the host.  (lambda, Q) then
        differ from LAPACK's in the last bits and in the basis of degenerate eigenspaces (the n - 2 noise eigenvalues of a rank-2
        constraint): onecons_qcqp's bracket depends on them (SURVEY a-11), so results agree with the host path to ~1e-9, not bit
        for bit -- tests/test_gpu_eval_sdr_admm.py holds the C4-shaped problem to 1e-6 on (f0, maxviol).  The default stays host."""
        import torch
        n, m = self.n, self.m
        Ps = np.stack([np.asarray(((sp.csr_matrix(f[0]) + sp.csr_matrix(f[0]).T) / 2.).todense()) for f in self.forms[1:]])
        qs = np.stack([np.asarray(f[1].todense() if sp.issparse(f[1]) else f[1], dtype=np.float64).ravel() for f in self.forms[1:]])
        dP = torch.from_numpy(Ps).cuda()
        lmb, Q = torch.linalg.eigh(dP)                                    # [m][n], [m][n][n]
        qhat = torch.bmm(Q.transpose(1, 2), torch.from_numpy(qs).cuda().unsqueeze(2)).squeeze(2)
        self.set_eig(lmb.cpu().numpy(), Q.cpu().numpy(), qhat.cpu().numpy())

    def zinv_device(self, rhos):
        """inverse of 2 (P0 + rho m I) for every rho as one batched Cholesky solve on the GPU (f-3, opt-in; the reference hands
        the matrix to SuperLU, qcqp.py:224-227)."""
        import torch
        P0 = torch.from_numpy(np.asarray(sp.csr_matrix(self.forms[0][0]).todense())).cuda()
        eye = torch.eye(self.n, dtype=torch.float64, device=P0.device)
        A = torch.stack([2 * (P0 + float(r) * self.m * eye) for r in np.atleast_1d(rhos)])
        Lc = torch.linalg.cholesky(A)
        return np.ascontiguousarray(torch.cholesky_inverse(Lc).cpu().numpy())

    def zinv(self, rho):
        """inverse of 2 (P0 + rho m I): the matrix qcqp.py:226-227 factorises for the z-update."""
        P0 = np.asarray(sp.csr_matrix(self.forms[0][0]).todense())
        return np.ascontiguousarray(np.linalg.inv(2 * (P0 + rho * self.m * np.eye(self.n))))

    def admm_improve(self, X0, rhos, num_iters=1000, viol_lim=1e4, tol=1e-2, phase1=True, setup="host"):
        """improve_admm for every (rho, start) pair (qcqp.py:254-285). Returns (X[K][R][n], f0[K][R], maxviol[K][R], stats).
        setup="device": the eigendecompositions and the per-rho inverses are computed on the GPU (f-3, see compute_eig_device)."""
        if not self._has_eig:
            self.compute_eig_device() if setup == "device" else self.compute_eig()
        X0 = np.ascontiguousarray(X0, dtype=np.float64).reshape(-1, self.n)
        R = X0.shape[0]
        rhos = np.ascontiguousarray(np.atleast_1d(rhos), dtype=np.float64)
        K = len(rhos)
        Zinv = self.zinv_device(rhos) if setup == "device" else np.ascontiguousarray(np.stack([self.zinv(r) for r in rhos]))
        X = np.empty((K, R, self.n)); f0 = np.empty((K, R)); mv = np.empty((K, R))
        stats = (AdmmStats * (K * R))()
        prm = AdmmParams(int(num_iters), float(viol_lim), float(tol), int(bool(phase1)))
        check(_lib.load().qcqp_admm_improve(self._h, C.byref(prm), _ptr(rhos), _ptr(Zinv), K, _ptr(X0), R, _ptr(X), _ptr(f0), _ptr(mv),
                                            C.cast(stats, C.c_void_p)))
        return X, f0, mv, stats

    # ---- SDR sampler -----------------------------------------------------------------------------------
    def sdr_sample_eval(self, mu, F, Z=None, S=None, seed=0):
        """x_s = mu + z_s F and (f0, maxviol) per draw (qcqp.py:396-401). Z: [S][n] standard normals (parity) or None
        (device Philox stream from `seed`, S draws)."""
        mu = np.ascontiguousarray(mu, dtype=np.float64); F = np.ascontiguousarray(F, dtype=np.float64)
        if Z is not None:
            Z = np.ascontiguousarray(Z, dtype=np.float64).reshape(-1, self.n)
            S = Z.shape[0]
        X = np.empty((S, self.n)); f0 = np.empty(S); mv = np.empty(S)
        check(_lib.load().qcqp_sdr_sample_eval(self._h, _ptr(mu), _ptr(F), _ptr(Z) if Z is not None else None, int(seed), int(S),
                                               _ptr(X), _ptr(f0), _ptr(mv)))
        return X, f0, mv


    def sdr_prefetch(self, Z):
        """Starts the upload of the standard normals of a LATER sdr_cd_pipeline(Z=Z) call (same array object, pinned for the copy to
        be asynchronous) and returns: issued before the call on the current batch, the copy hides behind that call's kernels."""
        if not (isinstance(Z, np.ndarray) and Z.dtype == np.float64 and Z.flags["C_CONTIGUOUS"]):
            raise Exception("sdr_prefetch needs the C-contiguous float64 array that will be passed to sdr_cd_pipeline")
        check(_lib.load().qcqp_sdr_prefetch(self._h, _ptr(Z), int(Z.size // self.n)))

    def sdr_cd_pipeline(self, seeds, mu=None, F=None, Z=None, S=None, seed=0, want_draws=False, want_rng=False, out=None,
                        num_iters=1000, viol_tol=1e-2, tol=1e-4, phase1=True, strict=False, refresh_every=0):
        """S draws x_s = mu + z_s F (qcqp.py:394-401), improve_coord_descent of every draw with the stream of
        np.random.seed(seeds[s]) (qcqp.py:181-192), best pick (utilities.py:135-146) -- one call, the draws stay on the device.
        mu / F None: the factor cached on the pack by an earlier call.  out: optional preallocated (X, f0, maxviol) arrays
        (e.g. pinned).  Returns dict(X, f0, maxviol, stats, best[, X0, f0_draw, maxviol_draw][, rng])."""
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        if Z is not None:
            Z = np.ascontiguousarray(Z, dtype=np.float64).reshape(-1, self.n)
            S = Z.shape[0]
        S = int(S if S is not None else len(seeds))
        if len(seeds) != S:
            raise Exception("need one seed per draw")
        if mu is not None:
            mu = np.ascontiguousarray(mu, dtype=np.float64); F = np.ascontiguousarray(F, dtype=np.float64)
        if out is not None:
            X, f0, mv = out
        else:
            X = np.empty((S, self.n)); f0 = np.empty(S); mv = np.empty(S)
        stats = (CdStats * S)()
        X0 = np.empty((S, self.n)) if want_draws else None
        fd = np.empty(S) if want_draws else None; vd = np.empty(S) if want_draws else None
        rng = (RngState * S)() if want_rng else None
        bi = C.c_int32(-1)
        prm = CdParams(int(num_iters), float(viol_tol), float(tol), int(bool(phase1)), int(strict), int(refresh_every))
        check(_lib.load().qcqp_sdr_cd_pipeline(
            self._h, C.byref(prm), _ptr(mu) if mu is not None else None, _ptr(F) if mu is not None else None,
            _ptr(Z) if Z is not None else None, int(seed), S, _ptr(seeds), _ptr(X0) if want_draws else None,
            _ptr(fd) if want_draws else None, _ptr(vd) if want_draws else None, _ptr(X), _ptr(f0), _ptr(mv),
            C.cast(stats, C.c_void_p), C.cast(rng, C.c_void_p) if want_rng else None, C.cast(C.byref(bi), C.c_void_p)))
        res = dict(X=X, f0=f0, maxviol=mv, stats=stats, best=int(bi.value))
        if want_draws:
            res.update(X0=X0, f0_draw=fd, maxviol_draw=vd)
        if want_rng:
            res["rng"] = rng
        return res


class Comm:
    """NCCL communicator owned by the library (qcqp_comm_*): the best pick across GPUs without torch.distributed on the data
    path.  Rank 0 calls Comm.unique_id() and hands the 128 bytes to the other ranks; every rank then builds Comm(rank, n, id)."""

    @staticmethod
    def unique_id():
        buf = (C.c_char * 128)()
        check(_lib.load().qcqp_comm_unique_id(C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def __init__(self, rank, nranks, unique_id):
        if len(unique_id) != 128:
            raise Exception("the NCCL unique id is 128 bytes")
        h = C.c_void_p()
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        check(_lib.load().qcqp_comm_create(int(rank), int(nranks), C.cast(buf, C.c_void_p), C.byref(h)))
        self._h, self.rank, self.nranks = h, int(rank), int(nranks)

    def best(self, d_f0, d_maxviol, d_X, R, n, index_offset=0, d_xbest=0, tol=1e-4, stream=0):
        """Device pointers (ints) of this rank's f0[R], maxviol[R], X[R][n]; returns (global index, rank, f0, maxviol) of the best
        restart over all ranks in QCQPForm.better order; its point is written to d_xbest (device pointer, may be 0)."""
        bi, br, bf, bv = C.c_int64(-1), C.c_int32(-1), C.c_double(0), C.c_double(0)
        check(_lib.load().qcqp_best_multi(self._h, d_f0, d_maxviol, d_X, int(R), int(n), float(tol), int(index_offset), C.byref(bi), C.byref(br),
                                          C.byref(bf), C.byref(bv), d_xbest, stream))
        return int(bi.value), int(br.value), float(bf.value), float(bv.value)

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().qcqp_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def best(f0, maxviol, tol=1e-4):
    """Index of the best point in QCQPForm.better order (utilities.py:135-146)."""
    f0 = np.ascontiguousarray(f0, dtype=np.float64).ravel(); mv = np.ascontiguousarray(maxviol, dtype=np.float64).ravel()
    out = C.c_int32(-1)
    check(_lib.load().qcqp_best(_ptr(f0), _ptr(mv), len(f0), float(tol), C.byref(out)))
    return int(out.value)


def sdr_factor(Xstar, eps=1e-8, corrected=False):
    """(mu, Sigma, F) from the relaxed solution X* as QCQP.suggest builds them (qcqp.py:394-395), F being the factor
    np.random.multivariate_normal derives from Sigma by SVD, so that a draw is mu + standard_normal(n) @ F.
    corrected=False keeps the reference's row-broadcast `mu*mu.T`; True uses the intended outer product."""
    Xs = np.asarray(Xstar, dtype=np.float64)
    n = Xs.shape[0] - 1
    mu = Xs[:-1, -1].copy()
    if corrected:
        Sigma = Xs[:-1, :-1] - np.outer(mu, mu) + eps * np.eye(n)
    else:
        Sigma = Xs[:-1, :-1] - mu * mu.T + eps * np.eye(n)
    _u, s, vt = np.linalg.svd(Sigma)
    return mu, Sigma, np.ascontiguousarray(np.sqrt(s)[:, None] * vt)
