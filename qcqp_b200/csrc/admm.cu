// admm.cu -- consensus ADMM with one-constraint projections (improve_admm qcqp.py:254-285; admm_phase1 :195-212;
// admm_phase2 :215-251; onecons_qcqp utilities.py:149-196; QCQPForm.better :135-146), for K rho values x R starts.
//
// Mapping: one CTA per (rho, start) run.  Inside an iteration the m projections are independent (the reference's
// "TODO: parallel x/u-updates", qcqp.py:234): warps take constraints round-robin; a projection is two coalesced
// GEMV passes over Q_i / Q_i^T (lanes own output components) around a warp-shuffle bisection on the multiplier.
// z, the z-update right-hand side and the best point live in shared memory; xs/us [m][n] per run in HBM/L2.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "forms_eval.cuh"
#include "onevar.cuh"

namespace qcqp {

struct AdmmK {
    int num_iters;
    double viol_lim, tol;
    int phase1;
};

constexpr int ADMM_THREADS = 512;   // 16 warps: 16 projections in flight per run; 128 registers each fill the file
constexpr int ADMM_WARPS = ADMM_THREADS / 32;

struct AdmmSmem {
    double* z; double* last_z; double* bestx; double* x1; double* rhs; double* q0; double* x0;
    double* wbuf;    // [warps][3][npad]: v, zhat, xhat
    double* red;     // [warps + 2]
};

// (f0(x), max violation(x)) with the whole CTA: warps take blocks of 32 forms
__device__ __forceinline__ void block_eval(const PackView& P, const double* x, double* red, double* f0_out, double* mv_out)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double mv = -QCQP_INF, fobj = 0.0;
    for (int base = warp * 32; base <= P.m; base += ADMM_WARPS * 32) {
        int hi = base + 31 < P.m ? base + 31 : P.m;
        eval_forms(P, x, base, hi, false, lane, [&](int j, double v) {
            if (j == 0) fobj = v;
            else {
                double vv = violation_of(P.relop[j], v);
                mv = (vv > mv) ? vv : mv;
            }
        });
    }
    mv = warp_max(mv);
    fobj = warp_sum(fobj);
    __syncthreads();
    if (lane == 0) red[warp] = mv;
    if (warp == 0 && lane == 0) red[ADMM_WARPS] = fobj;
    __syncthreads();
    double r = red[0];
    for (int i = 1; i < ADMM_WARPS; i++) r = red[i] > r ? red[i] : r;
    *mv_out = r;
    *f0_out = red[ADMM_WARPS];
    __syncthreads();
}

// QCQPForm.better on cached (maxviol, f0) pairs: true when the FIRST argument is returned
__device__ __forceinline__ bool better_first(double mv1, double f1, double mv2, double f2)
{
    long long v1 = (long long)(mv1 / 1e-4), v2 = (long long)(mv2 / 1e-4);
    if (v1 < v2) return true;
    if (v2 < v1) return false;
    return f1 < f2;
}

// onecons_qcqp(v, f_i) by one warp; v, zhat, xhat are this warp's smem buffers; out -> xs_i (global)
__device__ __forceinline__ void project_one(const PackView& P, int i /*0-based constraint*/, double* v, double* zhat, double* xhat,
                                            double* out, int lane)
{
    const int n = P.n;
    const int j = i + 1;
    if (P.relop[j] == QCQP_RELOP_LE) {
        double fv = (P.dense_slot[j] >= 0) ? eval_dense_form_warp(P, j, v, lane)
                  : ((P.f_ptr[j + 1] - P.f_ptr[j]) > EVAL_LONG_FORM ? eval_sparse_form_warp(P, j, v, lane)
                                                                   : bcast(eval_sparse_form_seq(P, j, v), 0));
        if (fv <= 0) {
            for (int a = lane; a < n; a += 32) out[a] = v[a];
            return;
        }
    }
    const double* lam = P.eig_lambda + (size_t)i * n;
    const double* qh = P.eig_qhat + (size_t)i * n;
    const double* Q = P.eig_Q + (size_t)i * n * n;
    const double* Qt = P.eig_Qt + (size_t)i * n * n;
    const double r = P.r[j];
    // zhat = Q^T v : zhat_b = sum_a Q[a][b] v_a ; lanes own b
    for (int b0 = 0; b0 < n; b0 += 128) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        const int b = b0 + lane;
        const bool v0 = b < n, v1 = b + 32 < n, v2 = b + 64 < n, v3 = b + 96 < n;
        for (int a = 0; a < n; a++) {
            const double va = v[a];
            const double* Qa = Q + (size_t)a * n;
            if (v0) a0 = fma(Qa[b], va, a0);
            if (v1) a1 = fma(Qa[b + 32], va, a1);
            if (v2) a2 = fma(Qa[b + 64], va, a2);
            if (v3) a3 = fma(Qa[b + 96], va, a3);
        }
        if (v0) zhat[b] = a0;
        if (v1) zhat[b + 32] = a1;
        if (v2) zhat[b + 64] = a2;
        if (v3) zhat[b + 96] = a3;
    }
    __syncwarp();
    // bracket of the multiplier (utilities.py:176-186)
    double s = -QCQP_INF, e = QCQP_INF;
    for (int t = lane; t < n; t += 32) {
        double l = lam[t];
        if (l > 0) { double c = -1. / l; s = c > s ? c : s; }
        if (l < 0) { double c = -1. / l; e = c < e ? c : e; }
    }
    s = warp_max(s);
    e = -warp_max(-e);
    auto phi = [&](double nu) {
        double a = 0.0, b = 0.0;
        for (int t = lane; t < n; t += 32) {
            double xh = -((nu * qh[t] - 2 * zhat[t]) / (2 * (1 + nu * lam[t])));
            a = fma(lam[t], xh * xh, a);
            b = fma(qh[t], xh, b);
        }
        return warp_sum(a) + warp_sum(b) + r;
    };
    int guard = 0;
    if (s == -QCQP_INF) { s = -1.; while (phi(s) <= 0 && ++guard < 4096) s *= 2.; }
    if (e == QCQP_INF) { e = 1.; while (phi(e) >= 0 && ++guard < 8192) e *= 2.; }
    while (e - s > 1e-6) {
        double mid = (s + e) / 2.;
        double ph = phi(mid);
        if (ph > 0) s = mid;
        else if (ph < 0) e = mid;
        else { s = e = mid; break; }
    }
    const double nu = (s + e) / 2.;
    for (int t = lane; t < n; t += 32) xhat[t] = -((nu * qh[t] - 2 * zhat[t]) / (2 * (1 + nu * lam[t])));
    __syncwarp();
    // out = Q xhat : out_a = sum_b Qt[b][a] xhat_b ; lanes own a
    for (int a0i = 0; a0i < n; a0i += 128) {
        double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
        const int a = a0i + lane;
        const bool v0 = a < n, v1 = a + 32 < n, v2 = a + 64 < n, v3 = a + 96 < n;
        for (int b = 0; b < n; b++) {
            const double xb = xhat[b];
            const double* Qb = Qt + (size_t)b * n;
            if (v0) c0 = fma(Qb[a], xb, c0);
            if (v1) c1 = fma(Qb[a + 32], xb, c1);
            if (v2) c2 = fma(Qb[a + 64], xb, c2);
            if (v3) c3 = fma(Qb[a + 96], xb, c3);
        }
        if (v0) out[a] = c0;
        if (v1) out[a + 32] = c1;
        if (v2) out[a + 64] = c2;
        if (v3) out[a + 96] = c3;
    }
}

__global__ void __launch_bounds__(ADMM_THREADS) admm_kernel(PackView P, AdmmK prm, const double* __restrict__ rhos,
                                                             const double* __restrict__ Zinv, int K, const double* __restrict__ X0, int R,
                                                             double* __restrict__ X, double* __restrict__ f0_out, double* __restrict__ mv_out,
                                                             qcqp_admm_stats* stats, double* ws)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int n = P.n, m = P.m;
    const int npad = (n + 1) & ~1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int run = blockIdx.x;
    const int kk = run / R, rr = run % R;
    const double rho = rhos[kk];
    const double* Zi = Zinv + (size_t)kk * n * n;
    double* sm = reinterpret_cast<double*>(smem);
    double* z = sm; double* last_z = z + npad; double* bestx = last_z + npad; double* x1 = bestx + npad;
    double* rhs = x1 + npad; double* q0 = rhs + npad; double* x0 = q0 + npad;
    double* wbuf = x0 + npad;
    double* red = wbuf + (size_t)ADMM_WARPS * 3 * npad;
    double* xs = ws + (size_t)run * 2 * m * n;
    double* us = xs + (size_t)m * n;
    double* myv = wbuf + (size_t)warp * 3 * npad;

    for (int a = tid; a < n; a += ADMM_THREADS) { x0[a] = X0[(size_t)rr * n + a]; q0[a] = 0.0; }
    __syncthreads();
    for (long long e = P.q_ptr[0] + tid; e < P.q_ptr[1]; e += ADMM_THREADS) q0[P.q_idx[e]] = P.q_val[e];
    __syncthreads();

    qcqp_admm_stats st;
    st.iters_p1 = st.iters_p2 = 0; st.onecons_calls = 0; st.status = 0; st.pad_ = 0;

    auto projections = [&]() {
        // xs[i] = onecons_qcqp(z + us[i], f_i); us[i] += z - xs[i]     (qcqp.py:205-209, 235-238)
        for (int i = warp; i < m; i += ADMM_WARPS) {
            double* xi = xs + (size_t)i * n;
            double* ui = us + (size_t)i * n;
            for (int a = lane; a < n; a += 32) myv[a] = z[a] + ui[a];
            __syncwarp();
            project_one(P, i, myv, myv + npad, myv + 2 * npad, xi, lane);
            __syncwarp();
            for (int a = lane; a < n; a += 32) ui[a] += z[a] - xi[a];
            __syncwarp();
        }
        __syncthreads();
    };
    auto init_state = [&](const double* from) {
        for (int a = tid; a < n; a += ADMM_THREADS) z[a] = from[a];
        for (size_t t = tid; t < (size_t)m * n; t += ADMM_THREADS) { xs[t] = from[t % n]; us[t] = 0.0; }
        __syncthreads();
    };

    double f_x0, mv_x0;
    block_eval(P, x0, red, &f_x0, &mv_x0);
    double f_x1 = f_x0, mv_x1 = mv_x0;
    for (int a = tid; a < n; a += ADMM_THREADS) x1[a] = x0[a];
    __syncthreads();

    if (prm.phase1) {
        init_state(x0);
        double fz = f_x0, mvz = mv_x0;
        for (int t = 0; t < prm.num_iters; t++) {
            if (mvz < prm.tol) break;
            st.iters_p1++;
            for (int a = tid; a < n; a += ADMM_THREADS) {
                double sx = 0.0, su = 0.0;
                for (int i = 0; i < m; i++) sx = sx + xs[(size_t)i * n + a];
                for (int i = 0; i < m; i++) su = su + us[(size_t)i * n + a];
                z[a] = (sx - su) / m;
            }
            __syncthreads();
            projections();
            st.onecons_calls += m;
            block_eval(P, z, red, &fz, &mvz);
        }
        // x1 = better(x0, z)
        if (!better_first(mv_x0, f_x0, mvz, fz)) {
            for (int a = tid; a < n; a += ADMM_THREADS) x1[a] = z[a];
            f_x1 = fz; mv_x1 = mvz;
        }
        __syncthreads();
    }

    // ---- phase 2 ----
    init_state(x1);
    for (int a = tid; a < n; a += ADMM_THREADS) bestx[a] = x1[a];
    double f_best = f_x1, mv_best = mv_x1;
    bool have_last = false;
    __syncthreads();
    for (int t = 0; t < prm.num_iters; t++) {
        st.iters_p2++;
        for (int a = tid; a < n; a += ADMM_THREADS) {
            double sx = 0.0, su = 0.0;
            for (int i = 0; i < m; i++) sx = sx + xs[(size_t)i * n + a];
            for (int i = 0; i < m; i++) su = su + us[(size_t)i * n + a];
            rhs[a] = 2 * rho * (sx - su) - q0[a];
        }
        __syncthreads();
        // z = (2 (P0 + rho m I))^{-1} rhs : warp per row, lanes over columns
        for (int a = warp; a < n; a += ADMM_WARPS) {
            const double* Za = Zi + (size_t)a * n;
            double acc = 0.0;
            for (int b = lane; b < n; b += 32) acc = fma(Za[b], rhs[b], acc);
            acc = warp_sum(acc);
            if (lane == 0) z[a] = acc;
        }
        __syncthreads();
        projections();
        st.onecons_calls += m;
        if (have_last) {
            double part = 0.0;
            for (int a = tid; a < n; a += ADMM_THREADS) { double d = last_z[a] - z[a]; part = fma(d, d, part); }
            part = warp_sum(part);
            if (lane == 0) red[warp] = part;
            __syncthreads();
            double tot = 0.0;
            for (int i = 0; i < ADMM_WARPS; i++) tot += red[i];
            __syncthreads();
            if (sqrt(tot) < prm.tol) break;
        }
        for (int a = tid; a < n; a += ADMM_THREADS) last_z[a] = z[a];
        have_last = true;
        double fz, mvz;
        block_eval(P, z, red, &fz, &mvz);
        if (mvz > prm.viol_lim) break;
        // bestx = better(z, bestx)
        if (better_first(mvz, fz, mv_best, f_best)) {
            for (int a = tid; a < n; a += ADMM_THREADS) bestx[a] = z[a];
            f_best = fz; mv_best = mvz;
        }
        __syncthreads();
    }
    // x2 = better(x1, bestx)
    const bool take1 = better_first(mv_x1, f_x1, mv_best, f_best);
    const double* res = take1 ? x1 : bestx;
    __syncthreads();
    for (int a = tid; a < n; a += ADMM_THREADS) X[(size_t)run * n + a] = res[a];
    if (tid == 0) {
        f0_out[run] = take1 ? f_x1 : f_best;
        mv_out[run] = take1 ? mv_x1 : mv_best;
        if (stats) stats[run] = st;
    }
}

int admm_res_try(qcqp_pack* p, const qcqp_admm_params* prm, const double* drhos, const double* dZinv, int K, const double* dX0, int R, double* dX,
                 double* df0, double* dmv, qcqp_admm_stats* dstats, cudaStream_t stream, bool* used);

int admm_launch(qcqp_pack* p, const qcqp_admm_params* prm, const double* drhos, const double* dZinv, int K, const double* dX0, int R,
                double* dX, double* df0, double* dmv, qcqp_admm_stats* dstats, cudaStream_t stream)
{
    const int runs = K * R;
    if (runs <= 0) return QCQP_OK;
    const PackView& v = p->v;
    if (v.m <= 0) return fail(QCQP_ERR_INVALID, "qcqp_admm_improve: the problem has no constraints");
    // constraint-resident kernel (admm_res.cu) whenever Q_i fits one SM's shared memory and m CTAs fit the GPU;
    // QCQP_ADMM_KERNEL=run forces the one-CTA-per-run kernel below (A/B runs, tests)
    const char* force = getenv("QCQP_ADMM_KERNEL");
    if (!(force && strcmp(force, "run") == 0)) {
        bool used = false;
        int rc0 = admm_res_try(p, prm, drhos, dZinv, K, dX0, R, dX, df0, dmv, dstats, stream, &used);
        if (used || rc0 != QCQP_OK) return rc0;
    }
    const int npad = (v.n + 1) & ~1;
    size_t smem = ((size_t)7 * npad + (size_t)ADMM_WARPS * 3 * npad + ADMM_WARPS + 4) * 8;
    if (smem > (size_t)max_smem_optin(p->device)) return fail(QCQP_ERR_CAPACITY, "qcqp_admm_improve: n too large for shared memory");
    int rc = ensure_workspace(p, (size_t)runs * 2 * v.m * v.n * 8);
    if (rc != QCQP_OK) return rc;
    AdmmK k;
    k.num_iters = prm->num_iters; k.viol_lim = prm->viol_lim; k.tol = prm->tol; k.phase1 = prm->phase1;
    QCQP_CUDA_TRY(cudaFuncSetAttribute(admm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    admm_kernel<<<runs, ADMM_THREADS, smem, stream>>>(v, k, drhos, dZinv, K, dX0, R, dX, df0, dmv, dstats, (double*)p->ws);
    QCQP_CUDA_TRY(cudaGetLastError());
    return QCQP_OK;
}

}  // namespace qcqp
