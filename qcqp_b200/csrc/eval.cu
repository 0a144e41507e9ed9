// eval.cu -- batched (f0, max violation) of R points (QuadraticFunction.eval utilities.py:49-50, violation :56-62,
// QCQPForm.violations :133-134; the pair returned by QCQP.suggest / QCQP._improve, qcqp.py:399-401, 415-417),
// and the best-pick reduction in QCQPForm.better order (utilities.py:135-146).
#include "common.cuh"
#include "forms_eval.cuh"
#include "onevar.cuh"

namespace qcqp {

// one warp per point; x staged in shared memory
__global__ void eval_kernel(PackView P, const double* __restrict__ X, int R, double* __restrict__ f0, double* __restrict__ maxviol,
                            double* __restrict__ viol)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int n = P.n, m = P.m;
    double* x = reinterpret_cast<double*>(smem) + (size_t)warp * ((n + 1) & ~1);
    for (int r = blockIdx.x * wpb + warp; r < R; r += gridDim.x * wpb) {
        __syncwarp();
        for (int i = lane; i < n; i += 32) x[i] = X[(size_t)r * n + i];
        __syncwarp();
        double mv = -QCQP_INF;
        double fobj = 0.0;
        const PackView& Pr = P;
        double* vrow = viol ? viol + (size_t)r * m : nullptr;
        eval_forms(P, x, 0, m, false, lane, [&](int j, double v) {
            if (j == 0) fobj = v;
            else {
                double vv = violation_of(Pr.relop[j], v);
                if (vrow) vrow[j - 1] = vv;
                mv = (vv > mv) ? vv : mv;
            }
        });
        mv = warp_max(mv);
        fobj = warp_sum(fobj);   // f_0 was produced by exactly one lane; the others hold 0
        if (lane == 0) {
            f0[r] = fobj;
            maxviol[r] = (m > 0) ? mv : 0.0;
        }
    }
}

int gemm_quadform_launch(int S, int n, const double* dX, const double* dPj, int ld, double* dpart, cudaStream_t stream);
int gemm_col_blocks(int n);

// Batched path for packs with dense forms: x_s' P_j x_s of every dense form as a tiled FP64 GEMM with a row-dot epilogue
// (gemm.cu), then this kernel -- one warp per point -- adds q_j.x + r_j, evaluates the sparse forms and reduces.
// sep: separable pack (LpcView) and no per-constraint output: constraint k is (c_p x_k + c_q) x_k + c_r read from the
// coordinate-major arrays -- the same operations in the same order as the generic one-entry form, without its index chasing.
__global__ void eval_finish_kernel(PackView P, LpcView V, int sep, const double* __restrict__ X, int R, const double* __restrict__ part,
                                   int ncb, double* __restrict__ f0, double* __restrict__ maxviol, double* __restrict__ viol)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int n = P.n, m = P.m;
    double* x = reinterpret_cast<double*>(smem) + (size_t)warp * ((n + 1) & ~1);
    for (int r = blockIdx.x * wpb + warp; r < R; r += gridDim.x * wpb) {
        __syncwarp();
        for (int i = lane; i < n; i += 32) x[i] = X[(size_t)r * n + i];
        __syncwarp();
        double mv = -QCQP_INF, fobj = 0.0;
        const PackView& Pr = P;
        double* vrow = viol ? viol + (size_t)r * m : nullptr;
        auto sink = [&](int j, double v) {
            if (j == 0) fobj = v;
            else {
                double vv = violation_of(Pr.relop[j], v);
                if (vrow) vrow[j - 1] = vv;
                mv = (vv > mv) ? vv : mv;
            }
        };
        if (sep) {
            eval_forms(P, x, 0, 0, false, lane, sink, /*skip_dense=*/true);      // a sparse objective, if any
            for (int k = lane; k < n; k += 32) {
                const double xk = x[k];
                const double vv = violation_of(V.c_rel[k], (V.c_p[k] * xk + V.c_q[k]) * xk + V.c_r[k]);
                mv = (vv > mv) ? vv : mv;
            }
        } else {
            eval_forms(P, x, 0, m, false, lane, sink, /*skip_dense=*/true);
        }
        for (int d = 0; d < P.n_dense; d++) {
            const int j = P.dense_form[d];
            double acc = 0.0;
            for (long long e = P.q_ptr[j] + lane; e < P.q_ptr[j + 1]; e += 32) acc = fma(P.q_val[e], x[P.q_idx[e]], acc);
            acc = warp_sum(acc);
            double quad = 0.0;
            for (int cb = 0; cb < ncb; cb++) quad += part[((size_t)d * ncb + cb) * R + r];   // fixed order: reproducible
            if (lane == 0) sink(j, quad + acc + P.r[j]);
        }
        mv = warp_max(mv);
        fobj = warp_sum(fobj);
        if (lane == 0) {
            f0[r] = fobj;
            maxviol[r] = (m > 0) ? mv : 0.0;
        }
    }
}

int eval_launch(qcqp_pack* p, const double* dX, int R, double* df0, double* dmv, double* dviol, cudaStream_t stream)
{
    if (R <= 0) return QCQP_OK;
    const int n = p->v.n;
    if (p->v.n_dense > 0 && R >= 32) {
        const int nd = p->v.n_dense, ncb = gemm_col_blocks(n), ld = p->v.ld;
        int rc = ensure_workspace2(p, (size_t)nd * ncb * R * 8);
        if (rc != QCQP_OK) return rc;
        double* part = (double*)p->ws2;
        for (int d = 0; d < nd; d++) {
            rc = gemm_quadform_launch(R, n, dX, p->v.dense_P + (size_t)d * n * ld, ld, part + (size_t)d * ncb * R, stream);
            if (rc != QCQP_OK) return rc;
        }
        const int wpb = 4;
        size_t smem = (size_t)wpb * ((n + 1) & ~1) * 8;
        if (smem > (size_t)max_smem_optin(p->device)) return fail(QCQP_ERR_CAPACITY, "qcqp_eval: n too large for the shared-memory staging of x");
        QCQP_CUDA_TRY(cudaFuncSetAttribute(eval_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        eval_finish_kernel<<<(R + wpb - 1) / wpb, wpb * 32, smem, stream>>>(p->v, p->lpc, (p->lpc_ok && !dviol) ? 1 : 0, dX, R, part, ncb, df0, dmv,
                                                                            dviol);
        QCQP_CUDA_TRY(cudaGetLastError());
        return QCQP_OK;
    }
    const int wpb = 4;
    size_t smem = (size_t)wpb * ((n + 1) & ~1) * 8;
    if (smem > (size_t)max_smem_optin(p->device)) return fail(QCQP_ERR_CAPACITY, "qcqp_eval: n too large for the shared-memory staging of x");
    QCQP_CUDA_TRY(cudaFuncSetAttribute(eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = (R + wpb - 1) / wpb;
    eval_kernel<<<blocks, wpb * 32, smem, stream>>>(p->v, dX, R, df0, dmv, dviol);
    QCQP_CUDA_TRY(cudaGetLastError());
    return QCQP_OK;
}

// ---------------------------------------------------------------------------------------------------------
// best pick: lexicographic min on (int(maxviol / tol), f0); among exact ties the LATER index wins, which is what
// folding `best = better(best, x_r)` over r = 0..R-1 does (better returns its second argument on a tie).
// Single CTA: R is a restart count (thousands), the whole reduction is a few microseconds.
// ---------------------------------------------------------------------------------------------------------
struct BestKey { long long bucket; double f; int idx; };

__device__ __forceinline__ bool best_before(const BestKey& a, const BestKey& b)
{
    // true when a beats b
    if (a.idx < 0) return false;
    if (b.idx < 0) return true;
    if (a.bucket != b.bucket) return a.bucket < b.bucket;
    if (a.f != b.f) return a.f < b.f;
    return a.idx > b.idx;
}

__global__ void best_kernel(const double* __restrict__ f0, const double* __restrict__ maxviol, int R, double tol, int* best_idx,
                            long long* best_bucket, double* best_f0)
{
    __shared__ BestKey sh[32];
    BestKey me;
    me.idx = -1; me.bucket = 0; me.f = 0.0;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        BestKey c;
        c.idx = r;
        c.bucket = (long long)(maxviol[r] / tol);   // int(max(violations) / tol): truncation toward zero
        c.f = f0[r];
        if (c.f != c.f) continue;                   // NaN objective never wins (f1 < f2 is False both ways)
        if (best_before(c, me)) me = c;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) {
        BestKey ot;
        ot.bucket = __shfl_xor_sync(FULL, me.bucket, o);
        ot.f = __shfl_xor_sync(FULL, me.f, o);
        ot.idx = __shfl_xor_sync(FULL, me.idx, o);
        if (best_before(ot, me)) me = ot;
    }
    if (lane == 0) sh[warp] = me;
    __syncthreads();
    if (warp == 0) {
        int nw = blockDim.x >> 5;
        BestKey v;
        v.idx = -1; v.bucket = 0; v.f = 0.0;
        if (lane < nw) v = sh[lane];
        for (int o = 16; o > 0; o >>= 1) {
            BestKey ot;
            ot.bucket = __shfl_xor_sync(FULL, v.bucket, o);
            ot.f = __shfl_xor_sync(FULL, v.f, o);
            ot.idx = __shfl_xor_sync(FULL, v.idx, o);
            if (best_before(ot, v)) v = ot;
        }
        if (lane == 0) {
            *best_idx = v.idx;
            if (best_bucket) *best_bucket = v.bucket;
            if (best_f0) *best_f0 = v.f;
        }
    }
}

int best_launch(const double* df0, const double* dmv, int R, double tol, int* dbest, long long* dbucket, double* dbf, cudaStream_t stream)
{
    best_kernel<<<1, 1024, 0, stream>>>(df0, dmv, R, tol, dbest, dbucket, dbf);
    QCQP_CUDA_TRY(cudaGetLastError());
    return QCQP_OK;
}

}  // namespace qcqp
