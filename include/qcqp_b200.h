/*
 * qcqp_b200.h -- C ABI of libqcqp_b200.so, the B200 (sm_100a) engine behind the Suggest-and-Improve
 * hot path of cvxgrp/qcqp.
 *
 * The reference has no FFI: its seam is the Python function boundary QCQP.suggest / QCQP._improve
 * dispatch through.  Each entry point below replaces one of those functions;
 * the reference line it stands in for is cited on the declaration.  INTEGRATION.md shows the ctypes
 * binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative qcqp_status; qcqp_last_error() (thread-local)
 *     holds the message.  No exceptions cross the boundary.
 *   - all floating point is IEEE binary64; matrices are row-major; index arrays are int32 unless noted.
 *   - "host" entry points take caller-owned host buffers and return after the stream is synchronised.
 *     "_device" entry points take device pointers (cudaMalloc'ed / torch tensors' data_ptr()) and a
 *     cudaStream_t passed as void*; they only enqueue work -- after qcqp_pack_reserve, or once the pack's
 *     grow-on-demand workspaces have reached the batch size (the first call at a new size allocates).
 *   - a pack is bound to the CUDA device that was current when it was created; it is not thread-safe
 *     (neither is the reference: it caches state on prob/f, utilities.py:46,129-130).
 */
#ifndef QCQP_B200_H
#define QCQP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    QCQP_OK = 0,
    QCQP_ERR_INVALID = -1,   /* bad argument */
    QCQP_ERR_CUDA = -2,      /* CUDA runtime error (message has the cudaError string) */
    QCQP_ERR_NOMEM = -3,     /* host or device allocation failed */
    QCQP_ERR_CAPACITY = -4,  /* problem exceeds a shared-memory capacity of the kernels */
    QCQP_ERR_NO_DEVICE = -5, /* no CUDA device: the engine has no CPU fallback */
    QCQP_ERR_NCCL = -6       /* NCCL could not be loaded, or an NCCL call failed (multi-GPU best pick) */
} qcqp_status;

enum { QCQP_RELOP_NONE = 0, QCQP_RELOP_LE = 1, QCQP_RELOP_EQ = 2 };

/* per-restart status written by the improve kernels (mirrors the Python exceptions of the reference) */
enum {
    QCQP_RUN_OK = 0,
    QCQP_RUN_EMPTY_MAX = 1,        /* qcqp.py:117  max() of an empty list: a coordinate in no constraint, phase 1 */
    QCQP_RUN_UNBOUNDED_UNIFORM = 2 /* utilities.py:267  np.random.uniform with an infinite bound (OverflowError) */
};

typedef struct qcqp_pack qcqp_pack; /* opaque: the stacked (P_j, q_j, r_j, relop_j), j = 0..m, laid out in HBM */

/*
 * Problem description: what get_qcqp_form (utilities.py:318-347) produces, flattened.
 * Form j = 0 is the objective in minimise form; j = 1..m are the scalar constraints.
 * P_j is symmetric (utilities.py:333,345).  Entries of one form are sorted by (row, col), no duplicates,
 * no explicit zeros.
 */
typedef struct {
    int32_t n;              /* variables */
    int32_t m;              /* scalar constraints */
    const int64_t* p_ptr;   /* [m+2]   entries of P_j are p_row/p_col/p_val[p_ptr[j] .. p_ptr[j+1]) */
    const int32_t* p_row;
    const int32_t* p_col;
    const double* p_val;
    const int64_t* q_ptr;   /* [m+2]   nonzeros of q_j, sorted by index */
    const int32_t* q_idx;
    const double* q_val;
    const double* r;        /* [m+1] */
    const int32_t* relop;   /* [m+1]  QCQP_RELOP_*; relop[0] = NONE */
    double dense_min_fill;  /* a form with nnz >= fill * n^2 (and n >= 16) is stored as a dense n x ld matrix;
                               <= 0 selects the default 0.25 */
} qcqp_pack_desc;

typedef struct {
    int32_t n, m;
    int32_t n_dense;        /* forms stored dense */
    int32_t max_incidence;  /* max over k of the number of forms that involve x_k */
    int64_t incidences;     /* INC of SURVEY 8d: sum over k */
    int64_t nnz_offdiag;    /* off-diagonal nonzeros of the sparse forms */
    int64_t device_bytes;   /* HBM held by the pack */
    double bytes_per_sweep_phase2; /* algorithmic bytes of one restart-sweep, streaming model */
    double bytes_per_sweep_phase1; /* same without the objective (phase 1 never reads P_0) */
    int32_t separable;      /* 1: every constraint touches one coordinate, one constraint per coordinate -> cd_lpc kernel */
    int32_t pad_;
} qcqp_pack_info;

/* NumPy RandomState (MT19937) state, field for field: np.random.get_state() -> (key, pos, has_gauss, cached_gaussian) */
typedef struct {
    uint32_t key[624];
    int32_t pos;
    int32_t has_gauss;
    double gauss;
} qcqp_rng_state;

/* kwargs of improve_coord_descent (qcqp.py:181-185) */
typedef struct {
    int32_t num_iters;      /* 1000 */
    double viol_tol;        /* 1e-2 */
    double tol;             /* 1e-4 */
    int32_t phase1;         /* 1 */
    int32_t strict;         /* 0 (default): fast -- dense row dots come from a cached g = P x kept current by an axpy on every
                                  move (the rows still stream through the TMA ring; the dot leaves the critical path).
                               1: strict -- every row dot summed sequentially in column order with separately rounded
                                  multiply/add, like SciPy's csr_matvec (bit-exact t1; slow).
                               2: fresh -- warp-parallel fma dot of the staged row at every step.
                               3: the general kernel in fast mode even when the problem is separable (A/B measurements).
                               With 0, separable problems (every constraint touches one coordinate, one constraint per
                               coordinate: Boolean LS, MAXCUT) run the lane-per-coordinate kernel (csrc/cd_lpc.cu). */
    int32_t refresh_every;  /* recompute the cached f_j(x) from scratch every this many phase-2 sweeps; 0 = 64 */
} qcqp_cd_params;

typedef struct {
    int64_t steps_p1;       /* coordinate steps executed in phase 1 / phase 2 */
    int64_t steps_p2;
    int64_t updates_p1;     /* accepted moves */
    int64_t updates_p2;
    int32_t sweeps_p1;      /* outer iterations entered (qcqp.py:110 / :160) */
    int32_t sweeps_p2;
    int32_t status;         /* QCQP_RUN_* */
    int32_t ran_phase2;
    int64_t steps_skipped;  /* phase-1 steps NOT executed: once a full sweep changes nothing (no move, no RNG draw) every
                               later sweep of qcqp.py:110 is the same no-op, so the engine jumps to the end of the loop */
} qcqp_cd_stats;

/* kwargs of improve_admm (qcqp.py:254-259); rho is per run, see qcqp_admm_improve */
typedef struct {
    int32_t num_iters;      /* 1000 */
    double viol_lim;        /* 1e4 */
    double tol;             /* 1e-2 */
    int32_t phase1;         /* 1 */
} qcqp_admm_params;

typedef struct {
    int32_t iters_p1;
    int32_t iters_p2;
    int64_t onecons_calls;
    int32_t status;
    int32_t pad_;
} qcqp_admm_stats;

/* ---- library ------------------------------------------------------------------------------------------- */
const char* qcqp_last_error(void);
int qcqp_device_count(void);
const char* qcqp_version(void);

/* ---- pack: replaces the QCQPForm / QuadraticFunction containers (utilities.py:41-46, 122-131) -------------- */
int qcqp_pack_create(const qcqp_pack_desc* desc, qcqp_pack** out);
void qcqp_pack_destroy(qcqp_pack* pack);
int qcqp_pack_get_info(const qcqp_pack* pack, qcqp_pack_info* info);
/* Sizes every grow-on-demand device buffer of the pack for batches of up to R restarts / draws (and K rho values), so that the
 * `_device` calls that follow only enqueue work: no cudaMalloc / cudaFree (an implicit device-wide synchronisation, illegal during
 * CUDA-graph capture) inside a stream of work.  Without it the first call at a new size grows the buffers.  A pack is
 * single-stream: its workspaces are shared by its calls (the reference's seam is not re-entrant either). */
int qcqp_pack_reserve(qcqp_pack* pack, int32_t R, int32_t K);

/* ---- batched evaluation: QuadraticFunction.eval / QCQPForm.violations for R points
 *      (utilities.py:49-50, 56-62, 133-134; the (f0, max violation) pair of qcqp.py:399-401, 415-417) ---------- */
int qcqp_eval(qcqp_pack* pack, const double* X /*[R][n]*/, int32_t R, double* f0 /*[R]*/, double* maxviol /*[R]*/,
              double* viol /*[R][m] or NULL*/);
int qcqp_eval_device(qcqp_pack* pack, const double* dX, int32_t R, double* df0, double* dmaxviol, double* dviol, void* stream);

/* ---- coordinate descent: improve_coord_descent(x0, prob, **kwargs) for R independent restarts
 *      (qcqp.py:181-192 -> coord_descent_phase1 :101-148, coord_descent_phase2 :152-178,
 *       get_onevar_func utilities.py:99-105, onevar_qcqp :241-288, get_feasible_intervals :198-232).
 *      Restart r consumes its own MT19937 stream rng[r] exactly as the reference consumes np.random. -------- */
int qcqp_cd_improve(qcqp_pack* pack, const qcqp_cd_params* params, const double* X0 /*[R][n]*/, int32_t R,
                    qcqp_rng_state* rng /*[R] in/out*/, double* X /*[R][n]*/, double* f0 /*[R]*/, double* maxviol /*[R]*/,
                    qcqp_cd_stats* stats /*[R] or NULL*/);
int qcqp_cd_improve_device(qcqp_pack* pack, const qcqp_cd_params* params, const double* dX0, int32_t R,
                           qcqp_rng_state* drng, double* dX, double* df0, double* dmaxviol, qcqp_cd_stats* dstats,
                           void* stream);

/* Device time (ms, CUDA events on the launching stream) of the launches behind the last qcqp_cd_improve* call:
 * ms[0] phase-1 kernel, ms[1] G = X P0 GEMM, ms[2] phase-2 kernel, ms[3] batched (f0, maxviol).  *count = 4 when the call took
 * the separable dense-objective path (the only one split into launches), else 0.  Call after synchronising the stream. */
int qcqp_cd_get_timing(qcqp_pack* pack, double* ms /*[4]*/, int32_t* count);

/* Device counters of the last qcqp_cd_improve* call (zeroed at its start; kernels that do not count leave zeros):
 * out[0] rows of P_0 applied to the cached g = P_0 x (one per accepted phase-2 move), out[1] 32 x 32 diagonal blocks of P_0
 * staged by TMA, out[2] rows read by from-scratch refreshes of g, out[3] bytes the phase-2 kernel requested from L2.
 * *count = 4.  Call after synchronising the stream.  bench.py derives the L2 roofline of the phase-2 kernel from out[3]. */
int qcqp_cd_get_counters(qcqp_pack* pack, uint64_t* out /*[4]*/, int32_t* count);

/* ---- consensus ADMM: improve_admm(x0, prob, rho=...) for K rho values x R starts
 *      (qcqp.py:254-285 -> admm_phase1 :195-212, admm_phase2 :215-251, onecons_qcqp utilities.py:149-196,
 *       QCQPForm.better :135-146).
 *      The host supplies what the reference caches: eigh of each constraint (utilities.py:160-166) and, per rho,
 *      the inverse of 2(P0 + rho m I) that qcqp.py:226-227 factorises. ------------------------------------------ */
int qcqp_admm_pack_eig(qcqp_pack* pack, const double* lambda /*[m][n]*/, const double* Q /*[m][n][n]*/,
                       const double* qhat /*[m][n] = Q_i^T q_i*/);
int qcqp_admm_improve(qcqp_pack* pack, const qcqp_admm_params* params, const double* rhos /*[K]*/,
                      const double* Zinv /*[K][n][n]*/, int32_t K, const double* X0 /*[R][n]*/, int32_t R,
                      double* X /*[K][R][n]*/, double* f0 /*[K][R]*/, double* maxviol /*[K][R]*/,
                      qcqp_admm_stats* stats /*[K][R] or NULL*/);
int qcqp_admm_improve_device(qcqp_pack* pack, const qcqp_admm_params* params, const double* drhos, const double* dZinv,
                             int32_t K, const double* dX0, int32_t R, double* dX, double* df0, double* dmaxviol,
                             qcqp_admm_stats* dstats, void* stream);

/* ---- SDR randomized rounding: x_s = mu + z_s F, then (f0, max violation), for S draws
 *      (qcqp.py:394-401; F = sqrt(s)[:,None]*Vt is the factor np.random.multivariate_normal builds).
 *      Z != NULL: the caller's standard normals (parity with np.random.standard_normal).
 *      Z == NULL: device Philox + Box-Muller from `seed`. --------------------------------------------------------- */
int qcqp_sdr_sample_eval(qcqp_pack* pack, const double* mu /*[n]*/, const double* F /*[n][n]*/, const double* Z /*[S][n] or NULL*/,
                         uint64_t seed, int32_t S, double* X /*[S][n]*/, double* f0 /*[S]*/, double* maxviol /*[S]*/);
int qcqp_sdr_sample_eval_device(qcqp_pack* pack, const double* dmu, const double* dF, const double* dZ, uint64_t seed,
                                int32_t S, double* dX, double* df0, double* dmaxviol, void* stream);

/* ---- Suggest-and-Improve pipeline for S draws in ONE call: SDR randomized rounding (qcqp.py:394-401) -> improve_coord_descent of
 *      every draw (qcqp.py:181-192) -> best pick in the `better` order (utilities.py:135-146).  What a user of the reference
 *      writes as `for s in range(S): qcqp.suggest(SDR); qcqp.improve(COORD_DESCENT)`; here the draws never leave the device.
 *      mu / F == NULL: the factor cached on the pack by the last call that supplied one (the reference caches mu and Sigma on
 *      self, qcqp.py:394-395).  Z as in qcqp_sdr_sample_eval.  seeds[s]: restart s consumes the stream of
 *      np.random.seed(seeds[s]) (MT19937 init_genrand, built on the device).  X0 / f0_draw / maxviol_draw (the draws and their
 *      (f, v)), stats, rng_out (final stream states) and best_idx may be NULL. -------------------------------------------- */
int qcqp_sdr_cd_pipeline(qcqp_pack* pack, const qcqp_cd_params* params, const double* mu /*[n] or NULL*/,
                         const double* F /*[n][n] or NULL*/, const double* Z /*[S][n] or NULL*/, uint64_t seed, int32_t S,
                         const uint32_t* seeds /*[S]*/, double* X0 /*[S][n] or NULL*/, double* f0_draw /*[S] or NULL*/,
                         double* maxviol_draw /*[S] or NULL*/, double* X /*[S][n]*/, double* f0 /*[S]*/, double* maxviol /*[S]*/,
                         qcqp_cd_stats* stats /*[S] or NULL*/, qcqp_rng_state* rng_out /*[S] or NULL*/, int32_t* best_idx /*or NULL*/);
/* Batch after batch: starts the upload of the standard normals Z [S][n] (pinned host memory, for the copy to be asynchronous) of a
 * LATER qcqp_sdr_cd_pipeline call on a private stream and returns; issued before the call on the current batch, the copy runs beside
 * that call's kernels.  The later call recognises Z by its address and S and waits for the copy instead of uploading; any other Z
 * is uploaded as usual.  Z must stay unchanged until that call returns.  At most two prefetches are outstanding (the third
 * replaces the first).  No reference counterpart: the reference draws on the host, one sample per suggest() (qcqp.py:394-401). */
int qcqp_sdr_prefetch(qcqp_pack* pack, const double* Z /*[S][n]*/, int32_t S);

/* ---- best pick: argmin in the QCQPForm.better order (utilities.py:135-146): lexicographic on
 *      (int(maxviol / tol), f0), later index wins exact ties. ---------------------------------------------------- */
int qcqp_best(const double* f0, const double* maxviol, int32_t R, double tol, int32_t* best_idx);
int qcqp_best_device(const double* df0, const double* dmaxviol, int32_t R, double tol, int32_t* dbest_idx,
                     int64_t* dbest_bucket, double* dbest_f0, void* stream);

/* ---- best pick ACROSS GPUs (QCQPForm.better, utilities.py:135-146, over the restarts of every rank).
 *      One process per GPU; restarts are sharded contiguously with no collective on the data path; this is the only exchange:
 *      ONE NCCL all-gather of (f0, maxviol, global index, x) per rank, then the same `better` fold over the ranks.  NCCL is
 *      bound at run time (dlopen of libnccl.so.2; QCQP_NCCL_LIB overrides the name).
 *        qcqp_comm_unique_id  rank 0 creates the 128-byte NCCL id; the caller hands it to the other ranks (MPI, a file, a socket).
 *        qcqp_comm_create     collective over the ranks; binds the communicator to the CUDA device that is current.
 *        qcqp_best_multi      collective.  df0 / dmaxviol / dX: this rank's R restarts on the device (R may be 0);
 *                             index_offset: global index of its first restart (ranks own increasing, disjoint ranges, so that
 *                             "the later index wins an exact tie" holds across ranks as it does inside qcqp_best).
 *                             Returns, identically on every rank: the winner's global index, its rank, (f0, maxviol), and its
 *                             point in dx_best (device, [n]; may be NULL).  Synchronises `stream` before returning. ---------- */
typedef struct qcqp_comm qcqp_comm;
int qcqp_comm_unique_id(void* id128 /* out: 128 bytes */);
int qcqp_comm_create(int32_t rank, int32_t nranks, const void* id128, qcqp_comm** out);
void qcqp_comm_destroy(qcqp_comm* comm);
int qcqp_best_multi(qcqp_comm* comm, const double* df0 /*[R]*/, const double* dmaxviol /*[R]*/, const double* dX /*[R][n] or NULL*/,
                    int32_t R, int32_t n, double tol, int64_t index_offset, int64_t* best_index, int32_t* best_rank,
                    double* best_f0, double* best_maxviol, double* dx_best /*[n] device or NULL*/, void* stream);

/* ---- measurement helpers (no counterpart in the reference): the ceilings bench.py reports its roofline fractions against,
 *      measured on the device that is current, with CUDA events.  The path is FP64 and its matrices are L2-resident, so the
 *      driver's HBM-copy / bf16-GEMM peaks do not bound it. ------------------------------------------------------------- */
int qcqp_probe_l2_bandwidth(int64_t bytes /* buffer size, must stay L2-resident (e.g. 32 MiB) */, int32_t iters,
                            double* gb_per_s /* sustained L2 -> SM read bandwidth */);
int qcqp_probe_fp64_peaks(double* dmma_tflops /* mma.sync.m8n8k4.f64 from registers */, double* dfma_tflops /* FMA pipe */);

#ifdef __cplusplus
}
#endif
#endif /* QCQP_B200_H */
