// common.cuh -- shared declarations of libqcqp_b200: error plumbing, the device-side pack view, warp helpers,
// mbarrier / bulk-copy (TMA) PTX wrappers for sm_100a.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/qcqp_b200.h"

namespace qcqp {

// ---------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define QCQP_CUDA_TRY(expr)                                                                                 \
    do {                                                                                                    \
        cudaError_t err__ = (expr);                                                                         \
        if (err__ != cudaSuccess)                                                                           \
            return ::qcqp::fail(QCQP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err__));      \
    } while (0)

// ---------------------------------------------------------------------------------------------------------
// device view of a pack (plain pointers; passed to kernels by value)
//
// HBM layout (DESIGN.md "data layout"):
//   coordinate-major incidence, for the CD sweep:
//     inc_ptr[n+1]; for incidence e in [inc_ptr[k], inc_ptr[k+1]):  (forms in ascending j; objective first)
//       inc_form[e]  = j | relop_j << 28 | dense << 31     inc_t2[e] = P_j[k,k]     inc_qk[e] = q_j[k]
//       off-diagonal entries of row k of P_j: row_col/row_val[inc_rbeg[e] .. inc_rbeg[e] + inc_rlen[e])  (sorted by
//       column); for a form stored dense inc_rlen[e] = -1 and inc_rbeg[e] = its dense slot
//   form-major COO, for f_j(x) from scratch:
//     f_ptr[m+2]; entries (f_row, f_col, f_val) sorted by (row, col);  q_ptr[m+2]; (q_idx, q_val); r[m+1]; relop[m+1]
//   dense forms: dense_P[slot][n][ld] row-major, ld = n rounded up to even (16-byte rows for cp.async.bulk)
// ---------------------------------------------------------------------------------------------------------
constexpr uint32_t INC_FORM_MASK = 0x0fffffffu;
constexpr int INC_RELOP_SHIFT = 28;
constexpr uint32_t INC_DENSE_BIT = 0x80000000u;

// one step of f_j(x) = sum_i ((P x)_i + q_i) x_i + r in SciPy's order: y += val * x[col] (col >= 0) or y += val (the q_i entry, col = -1);
// row >= 0 closes row i: acc += y * x[row], y = 0
struct EvOp { double val; int col; int row; };

struct PackView {
    int n, m, n_dense, ld;
    int max_inc;     // max incidences of one coordinate
    int max_inc_small;   // max incidences over the coordinates that have at most 1024 (their coefficient scratch lives in shared memory)
    int max_two;     // max number of two-interval-capable incidences of one coordinate (holes the 1-D sweep line can need)
    const int* inc_ptr;
    const uint32_t* inc_form;
    const double* inc_t2;
    const double* inc_qk;
    const int* inc_rbeg;
    const int* inc_rlen;
    const int* row_col;
    const double* row_val;
    const long long* f_ptr;
    const int* f_row;
    const int* f_col;
    const double* f_val;
    const long long* q_ptr;
    const int* q_idx;
    const double* q_val;
    // evaluation programs of the sparse forms: the operations of QuadraticFunction.eval in the reference's order, 16 bytes each, so that
    // a lane streams them with address-independent loads (forms_eval.cuh); ev_ptr[m+2]
    const long long* ev_ptr;
    const struct EvOp* ev_op;
    const double* r;
    const int* relop;
    const int* dense_slot;   // [m+1]: slot or -1
    const int* dense_form;   // [n_dense]: form of slot
    const double* dense_P;   // [n_dense][n][ld]
    // ADMM (set by qcqp_admm_pack_eig)
    const double* eig_lambda;  // [m][n]
    const double* eig_Q;       // [m][n][n]   Q_i[a][b]: component a of eigenvector b
    const double* eig_Qt;      // [m][n][n]   transposed copy, so both Q^T v and Q xhat read rows coalesced
    const double* eig_qhat;    // [m][n]
};

// Separable view (cd_lpc.cu): every constraint touches exactly one coordinate and every coordinate has exactly one
// constraint (Boolean least squares, MAXCUT, any per-coordinate quadratic constraint).  Then the one-variable restriction of
// coordinate k's constraint is the constant triple (c_p[k], c_q[k], c_r[k]) -- t0 = f_j(z) = r_j exactly, as the reference
// computes it -- and coordinates are independent of each other in phase 1.
struct LpcView {
    const double* c_p;     // [n] P_j[k,k]
    const double* c_q;     // [n] q_j[k]
    const double* c_r;     // [n] r_j
    const int* c_rel;      // [n] relop_j
    const double* o_diag;  // [n] P_0[k,k]
    const double* o_q;     // [n] q_0[k]
    const int* o_rbeg;     // [n] objective row k (off-diagonal) in row_col/row_val; unused when the objective is dense
    const int* o_rlen;     // [n]
    const unsigned char* o_inc;  // [n] 1 when the objective involves x_k structurally
    const double* cst6;    // [n][6] (c_p, c_q, c_r, c_rel, o_diag, o_q) of coordinate k side by side (cd_lpc2.cu stages them by TMA)
    double o_r;
    int obj_dense;
};

}  // namespace qcqp

// the opaque handle of the C ABI
struct qcqp_pack {
    qcqp::PackView v;
    qcqp_pack_info info;
    int device;
    std::vector<void*> allocs;   // every cudaMalloc owned by the pack
    // workspace, grown on demand
    void* ws;
    size_t ws_bytes;
    // device staging arena of the host-buffer entry points (grow-only, so steady-state calls do not cudaMalloc)
    void* io;
    size_t io_bytes;
    // scratch of the batched eval / SDR sampler (row-dot partials, device-generated normals)
    void* ws2;
    size_t ws2_bytes;
    // SDR factor of the last call that supplied one (the reference caches mu / Sigma on self, qcqp.py:394-395)
    double* sdr_mu;
    double* sdr_F;
    bool sdr_ok;
    bool has_eig;
    int objective_dense;
    bool lpc_ok;
    qcqp::LpcView lpc;
    // CUDA events around the launches of the last qcqp_cd_improve* call (see qcqp_cd_get_timing)
    cudaEvent_t ev[6];
    bool ev_ok;
    int ev_count;
    // TMA tensor map of the dense objective matrix (CUtensorMap, built on first use by cd_lpc2.cu): 0 not built, 1 ok, -1 failed
    alignas(64) unsigned char tmap[128];
    int tmap_state;
    // device counters of the last qcqp_cd_improve* call (qcqp_cd_get_counters): [0] rows of P_0 applied, [1] diagonal blocks
    // fetched, [2] rows read by from-scratch refreshes, [3] bytes requested from L2, [4..7] reserved
    unsigned long long* d_ctr;
    // host-buffer entry points: device-visible alias of the caller's PINNED result array X (else null).  A kernel that can deliver
    // every restart's point itself, as the restart finishes, writes through it and sets x_mirror_done; the read-back of X then
    // overlaps the tail of the launch instead of following it.
    double* x_mirror;
    bool x_mirror_done;
    // qcqp_sdr_prefetch: two device staging buffers for the standard normals of FUTURE qcqp_sdr_cd_pipeline calls, filled on a private
    // non-blocking stream while the current call computes; a pipeline call whose Z pointer matches a filled slot waits for that slot's
    // event instead of uploading
    double* zpre[2];
    size_t zpre_cap[2];
    const double* zpre_src[2];
    int zpre_S[2];
    unsigned long long zpre_seq[2], zpre_count;   // order of the prefetches: a call takes the OLDEST slot that matches
    cudaEvent_t zpre_ev[2];
    cudaStream_t zpre_stream;
    int zpre_next;
    // pinned staging block of the host-buffer pipeline's small results (f0, maxviol, statistics, best index): one read-back
    void* h_small;
    size_t h_small_cap;
};

namespace qcqp {

int ensure_workspace(qcqp_pack* p, size_t bytes);
int ensure_io(qcqp_pack* p, size_t bytes);
int ensure_workspace2(qcqp_pack* p, size_t bytes);
int num_sms(int device);
int max_smem_optin(int device);

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------------------
// warp helpers
// ---------------------------------------------------------------------------------------------------------
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { double w = __shfl_xor_sync(FULL, v, o); v = (w > v) ? w : v; }
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double bcast(double v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ int bcast_i(int v, int src) { return __shfl_sync(FULL, v, src); }

// ---------------------------------------------------------------------------------------------------------
// mbarrier + bulk async copy (TMA unit, non-tensor form) -- SASS: SYNCS.* / UBLKCP
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
// global -> shared bulk copy; bytes and both addresses must be multiples of 16
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
#endif  // __CUDACC__

}  // namespace qcqp
