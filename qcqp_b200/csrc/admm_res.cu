// admm_res.cu -- consensus ADMM with the eigenbasis of every constraint RESIDENT in shared memory (improve_admm
// qcqp.py:254-285; admm_phase1 :195-212; admm_phase2 :215-251; onecons_qcqp utilities.py:149-196; QCQPForm.better :135-146).
//
// admm.cu maps one CTA to one (rho, start) run, so every projection re-reads its Q_i (n x n doubles: 128 KB at n = 128) from
// L2, and a rho sweep of 16 runs occupies 16 of the 148 SMs.  The projections of one ADMM iteration are independent across
// constraints AND across runs, and all runs share the same Q_i.  Here the grid is m x G CTAs (cooperative launch):
//   CTA (i, g) keeps Q_i, lambda_i, Q_i^T q_i in shared memory for the whole solve and serves constraint i for the runs of
//   group g:   S1  z -> [Q^T z, Q^T (z + u_i)] (one pass over Q for all runs of a batch), f_i(z) and the early-out test from
//                  the rotated vectors (f_i(v) = sum lambda vhat^2 + qhat.vhat + r), multiplier bisection (one warp per run),
//                  x_i = Q xhat, u_i += z - x_i, publish d_i = x_i - u_i and the violation of z
//   barrier over the m CTAs of the group (one atomic counter per group, acquire/release)
//   the run's HOME CTA (local run index mod m):  S2  D = sum_i d_i, the reference's loop control (violation / step-length /
//                  viol_lim tests, best-so-far in the `better` order, phase 1 -> phase 2), next z = D/m or Zinv (2 rho D - q0)
//   barrier, next iteration.
// Per-run control state lives with the home CTA only; the other CTAs see a command word and z.
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"
#include "forms_eval.cuh"
#include "onevar.cuh"

namespace qcqp {

constexpr int RES_THREADS = 512;
constexpr int RES_WARPS = RES_THREADS / 32;
constexpr int RES_RB = 4;   // runs per GEMM batch (2 * RES_RB accumulators per thread)

// command word of a run, written by its home CTA: where this iteration's z is (home-written z at phase starts -- u is reset then --
// or the z rows all CTAs of the group computed during the previous S2) and which phase the run is in
enum { CMD_IDLE = 0, CMD_HOME_P1 = 1, CMD_HOME_P2 = 2, CMD_SPEC_P1 = 3, CMD_SPEC_P2 = 4 };

struct ResK {
    int num_iters;
    double viol_lim, tol;
    int phase1;
    int K, R, runs, G, rpg;   // rpg: runs per group
    int S, nb32;              // a-slices of the GEMM passes, n rounded up to a multiple of 32
    int ldq;                  // row stride of the resident Q_i: n + 1 (FMA passes) or = 4 mod 16 (DMMA fragments conflict-free)
    int no_mma;               // QCQP_ADMM_NO_MMA=1: FMA-pipe GEMM passes even when n is a multiple of 8 (A/B runs)
};

// per-run block in the workspace (doubles unless noted): see res_run_doubles()
struct RunView {
    double* z;        // [n]   z written by the home CTA (phase starts)
    double* zs;       // [2][n] z rows computed by all CTAs of the group, by step parity
    double* d;        // [m][n] x_i - u_i
    double* viol;     // [m]   violation of z for constraint i
    double* x0;       // [n]
    double* x1;       // [n]
    double* bestx;    // [n]
    double* last_z;   // [n]
    double* sc;       // [8] scalars: f_x0, mv_x0, f_x1, mv_x1, f_best, mv_best
    int* ic;          // [8] ints: cmd, phase, t, have_last, iters_p1, iters_p2, calls_lo, calls_hi
};
__host__ __device__ inline size_t res_run_doubles(int n, int m) { return (size_t)n * (7 + m) + m + 8 + 4; }
__device__ __forceinline__ RunView res_run_view(double* ws, int run, int n, int m)
{
    RunView v;
    double* b = ws + (size_t)run * res_run_doubles(n, m);
    v.z = b; v.zs = v.z + n; v.d = v.zs + 2 * n; v.viol = v.d + (size_t)m * n; v.x0 = v.viol + m; v.x1 = v.x0 + n; v.bestx = v.x1 + n; v.last_z = v.bestx + n;
    v.sc = v.last_z + n; v.ic = reinterpret_cast<int*>(v.sc + 8);
    return v;
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// barrier over the m CTAs of one group: monotone counter, generation = target / m
__device__ __forceinline__ void group_barrier(unsigned* ctr, unsigned target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        while (ld_acquire_u32(ctr) < target) { }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ bool res_better_first(double mv1, double f1, double mv2, double f2)
{
    long long v1 = (long long)(mv1 / 1e-4), v2 = (long long)(mv2 / 1e-4);
    if (v1 < v2) return true;
    if (v2 < v1) return false;
    return f1 < f2;
}

// exact (f0(x), max violation(x)) with the whole CTA, forms from HBM (start / result points only)
__device__ __forceinline__ void res_block_eval(const PackView& P, const double* x, double* red, double* f0_out, double* mv_out)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double mv = -QCQP_INF, fobj = 0.0;
    for (int base = warp * 32; base <= P.m; base += RES_WARPS * 32) {
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
    if (warp == 0 && lane == 0) red[RES_WARPS] = fobj;
    __syncthreads();
    double r = red[0];
    for (int i = 1; i < RES_WARPS; i++) r = red[i] > r ? red[i] : r;
    *mv_out = r;
    *f0_out = red[RES_WARPS];
    __syncthreads();
}

struct ResSmem {
    double* Q;      // [n][n + 1]  Q_i[a][b]: component a of eigenvector b (row pad: conflict-free by row and by column)
    double* lam;    // [n]
    double* qh;     // [n]
    double* u;      // [rpg][n]
    double* zs;     // [RB][n]   z of the batch; reused for x = Q xhat
    double* vs;     // [RB][n]   z + u
    double* zh;     // [RB][n]   Q^T z
    double* vh;     // [RB][n]   Q^T v
    double* xh;     // [RB][n]   xhat(nu)   (DMMA passes: [n][RB], run-interleaved)
    double* zvT;    // [n][2 RB] (z_0, v_0, .., z_3, v_3) per component: B fragments of the first DMMA pass
    double* part;   // [S][2 RB][n] partial sums of the GEMM passes; also scratch of the home stage (rhs, z, red)
    double* q0s;    // [n] q_0 dense
    int* cmds;      // [rpg] command of every run of my group, as read in S1 of this step
    double* hs;     // [2 * homes] f0(z) and |z - last_z|^2 of the runs this CTA is home of (S1 -> S2)
    int* flag;      // [RB] early-out / command
};

// D(8x8) += A(8x4, row) * B(4x8, col) in FP64 on the tensor pipe (DMMA).  Fragments: a = A[lane/4][lane%4], b = B[lane%4][lane/4],
// c0/c1 = C[lane/4][2 (lane%4) + {0,1}].  The only dense contraction of this kernel -- the two passes over the resident Q_i.
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// the multiplier search of onecons_qcqp (utilities.py:168-195) by one warp; zhat = Q^T v, result xhat.
// A lane keeps its (at most RES_EPL) eigen-components in registers for the ~65 evaluations of phi.
constexpr int RES_EPL = 6;   // n <= 192 (shared memory holds Q_i up to n ~ 165)

template <int NEL>
__device__ __forceinline__ void res_bisect_t(const double* lam, const double* qh, const double* zhat, double* xhat, int xstride, double r, int n, int lane)
{
    double L[NEL], Qh[NEL], Z2[NEL];
#pragma unroll
    for (int u = 0; u < NEL; u++) {
        const int t = lane + 32 * u;
        const bool ok = t < n;
        L[u] = ok ? lam[t] : 0.0; Qh[u] = ok ? qh[t] : 0.0; Z2[u] = ok ? 2 * zhat[t] : 0.0;   // a padded slot adds exactly 0 to phi
    }
    double s = -QCQP_INF, e = QCQP_INF;
#pragma unroll
    for (int u = 0; u < NEL; u++) {
        const double l = L[u];
        if (l > 0) { double c = -1. / l; s = c > s ? c : s; }
        if (l < 0) { double c = -1. / l; e = c < e ? c : e; }
    }
    s = warp_max(s);
    e = -warp_max(-e);
    // phi(nu) = sum lambda xhat^2 + qhat.xhat + r (utilities.py:169-175).  Only its SIGN steers the search, so the quotient
    // is taken with a reciprocal refined to full precision (1-2 ulp) instead of the 125-cycle IEEE division (straight-line code:
    // the NEL quotients overlap), and the two sums share one shuffle reduction.  A denominator outside the reciprocal's safe
    // range sends the whole warp through the exact division.  The returned xhat((s+e)/2) below always uses the exact division.
    double L2x[NEL];
#pragma unroll
    for (int u = 0; u < NEL; u++) L2x[u] = 2 * L[u];
    // xhat = -(nu qhat - 2 zhat) / (2 + nu 2 lambda) through the refined reciprocal: 9 FP64 instructions per component
    auto quot = [&](double nu, int u) {
        const double num = fma(nu, Qh[u], -Z2[u]), den = fma(nu, L2x[u], 2.0);
        double y;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(den));
        double e1 = fma(-den, y, 1.0); y = fma(y, e1, y);
        e1 = fma(-den, y, 1.0); y = fma(y, e1, y);
        return -(num * y);
    };
    auto exact_sum = [&](double nu) {
        double acc = 0.0;
#pragma unroll
        for (int u = 0; u < NEL; u++) {
            const double xh = -((nu * Qh[u] - Z2[u]) / (2 * (1 + nu * L[u])));
            acc = fma(fma(L[u], xh, Qh[u]), xh, acc);
        }
        return acc;
    };
    // a denominator outside the reciprocal's range (zero, denormal, overflow) shows up as a non-finite sum: redo exactly
    auto phi = [&](double nu) {
        double acc = 0.0;
#pragma unroll
        for (int u = 0; u < NEL; u++) { const double xh = quot(nu, u); acc = fma(fma(L[u], xh, Qh[u]), xh, acc); }
        acc = warp_sum(acc);
        if (!(fabs(acc) < QCQP_INF)) acc = warp_sum(exact_sum(nu));
        return acc + r;
    };
    // three multipliers at once (one bisection step and both of its possible successors): the 3 NEL quotient chains and the
    // three shuffle reductions overlap, so two levels of the search cost little more than one
    auto phi3 = [&](double nu0, double nu1, double nu2, double& o0, double& o1, double& o2) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
        for (int u = 0; u < NEL; u++) {
            const double x0 = quot(nu0, u), x1 = quot(nu1, u), x2 = quot(nu2, u);
            a0 = fma(fma(L[u], x0, Qh[u]), x0, a0);
            a1 = fma(fma(L[u], x1, Qh[u]), x1, a1);
            a2 = fma(fma(L[u], x2, Qh[u]), x2, a2);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a0 += __shfl_xor_sync(FULL, a0, o); a1 += __shfl_xor_sync(FULL, a1, o); a2 += __shfl_xor_sync(FULL, a2, o);
        }
        if (!(fabs(a0) < QCQP_INF) || !(fabs(a1) < QCQP_INF) || !(fabs(a2) < QCQP_INF)) {
            a0 = warp_sum(exact_sum(nu0)); a1 = warp_sum(exact_sum(nu1)); a2 = warp_sum(exact_sum(nu2));
        }
        o0 = a0 + r; o1 = a1 + r; o2 = a2 + r;
    };
    int guard = 0;
    if (s == -QCQP_INF) { s = -1.; while (phi(s) <= 0 && ++guard < 4096) s *= 2.; }
    if (e == QCQP_INF) { e = 1.; while (phi(e) >= 0 && ++guard < 8192) e *= 2.; }
    while (e - s > 1e-6) {
        // the reference's loop (utilities.py:187-195), two iterations per pass: mid, then (s+mid)/2 or (mid+e)/2 -- the same
        // expressions the sequential loop would evaluate next, so the decisions are identical
        const double mid = (s + e) / 2.;
        const double midl = (s + mid) / 2., midr = (mid + e) / 2.;
        double ph, phl, phr;
        phi3(mid, midl, midr, ph, phl, phr);
        if (ph > 0) s = mid;
        else if (ph < 0) e = mid;
        else { s = e = mid; break; }
        if (!(e - s > 1e-6)) break;
        const double ph2 = (ph > 0) ? phr : phl;
        const double mid2 = (s + e) / 2.;
        if (ph2 > 0) s = mid2;
        else if (ph2 < 0) e = mid2;
        else { s = e = mid2; break; }
    }
    const double nu = (s + e) / 2.;
    for (int t = lane; t < n; t += 32) xhat[(size_t)t * xstride] = -((nu * qh[t] - 2 * zhat[t]) / (2 * (1 + nu * lam[t])));
}

__device__ __forceinline__ void res_bisect(const double* lam, const double* qh, const double* zhat, double* xhat, int xstride, double r, int n, int lane)
{
    const int nel = (n + 31) >> 5;
    if (nel <= 1) res_bisect_t<1>(lam, qh, zhat, xhat, xstride, r, n, lane);
    else if (nel == 2) res_bisect_t<2>(lam, qh, zhat, xhat, xstride, r, n, lane);
    else if (nel <= 4) res_bisect_t<4>(lam, qh, zhat, xhat, xstride, r, n, lane);
    else res_bisect_t<RES_EPL>(lam, qh, zhat, xhat, xstride, r, n, lane);
}

__global__ void __launch_bounds__(RES_THREADS, 1) admm_res_kernel(const __grid_constant__ PackView P, ResK prm, const double* __restrict__ rhos,
                                                                   const double* __restrict__ Zinv, const double* __restrict__ X0,
                                                                   double* __restrict__ X, double* __restrict__ f0_out,
                                                                   double* __restrict__ mv_out, qcqp_admm_stats* stats, double* ws,
                                                                   unsigned* bars)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = P.n, m = P.m;
    const int ldq = prm.ldq;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ci = blockIdx.x % m;          // my constraint (0-based)
    const int g = blockIdx.x / m;           // my run group
    const int run0 = g * prm.rpg;
    const int nrun = (prm.runs - run0 < prm.rpg) ? (prm.runs - run0) : prm.rpg;
    const int S = prm.S, nb32 = prm.nb32;
    ResSmem sm;
    {
        double* p = reinterpret_cast<double*>(smem_raw);
        sm.Q = p; p += (size_t)n * ldq;
        sm.lam = p; p += n;
        sm.qh = p; p += n;
        sm.u = p; p += (size_t)prm.rpg * n;
        sm.zs = p; p += RES_RB * n;
        sm.vs = p; p += RES_RB * n;
        sm.zh = p; p += RES_RB * n;
        sm.vh = p; p += RES_RB * n;
        sm.xh = p; p += RES_RB * n;
        sm.zvT = p; p += 2 * RES_RB * n;
        sm.part = p; p += (size_t)S * 2 * RES_RB * n + 2 * n + RES_WARPS + 8;
        sm.hs = p; p += 2 * ((prm.rpg + m - 1) / m) + 2;
        sm.q0s = p; p += n;
        sm.cmds = reinterpret_cast<int*>(p); p += (prm.rpg + 1) / 2 + 1;
        sm.flag = reinterpret_cast<int*>(p);
    }
    unsigned* bar = bars + g;
    unsigned bar_target = 0;
    const int j = ci + 1;                   // form index of my constraint
    const int relop = P.relop[j];
    const double rj = P.r[j];

    // ---- resident data: Q_i (padded rows), lambda_i, Q_i^T q_i ----
    {
        const double* Qg = P.eig_Q + (size_t)ci * n * n;
        for (int e = tid; e < n * n; e += RES_THREADS) { const int a = e / n, b = e - a * n; sm.Q[a * ldq + b] = Qg[e]; }
        for (int t = tid; t < n; t += RES_THREADS) { sm.lam[t] = P.eig_lambda[(size_t)ci * n + t]; sm.qh[t] = P.eig_qhat[(size_t)ci * n + t]; }
        for (int t = tid; t < prm.rpg * n; t += RES_THREADS) sm.u[t] = 0.0;
        for (int t = tid; t < n; t += RES_THREADS) sm.q0s[t] = 0.0;
        __syncthreads();
        for (long long e = P.q_ptr[0] + tid; e < P.q_ptr[1]; e += RES_THREADS) sm.q0s[P.q_idx[e]] = P.q_val[e];
    }
    __syncthreads();

    // scratch of the home stage inside `part`
    double* h_rhs = sm.part;
    double* h_z = sm.part + n;
    double* h_red = sm.part + 2 * n;
    double* h_part = sm.part + 2 * n + RES_WARPS + 8;   // [S][n] shares of D

    // z_{t+1} of a run from D = sum_i (x_i - u_i): phase 1 D/m, phase 2 Zinv (2 rho D - q0); writes rv.z
    auto next_z = [&](const RunView& rv, int run, int phase, bool from_point, const double* point) {
        const int kk = run / prm.R;
        // D[a]: thread (slice, a) adds a contiguous share of the m terms, the S shares are added in slice order
        {
            const int slice = tid / nb32, a = tid - slice * nb32;
            if (slice < S && a < n) {
                const int i0 = (int)((long long)slice * m / S), i1 = (int)((long long)(slice + 1) * m / S);
                double D = 0.0;
                if (from_point) { for (int i = i0; i < i1; i++) D = D + point[a]; }      // xs = [x_init] * m, us = 0
                else {
                    int i = i0;
                    for (; i + 4 <= i1; i += 4) {
                        const double d0 = __ldcg(rv.d + (size_t)i * n + a), d1 = __ldcg(rv.d + (size_t)(i + 1) * n + a);
                        const double d2 = __ldcg(rv.d + (size_t)(i + 2) * n + a), d3 = __ldcg(rv.d + (size_t)(i + 3) * n + a);
                        D = (((D + d0) + d1) + d2) + d3;
                    }
                    for (; i < i1; i++) D = D + __ldcg(rv.d + (size_t)i * n + a);
                }
                h_part[slice * n + a] = D;
            }
        }
        __syncthreads();
        for (int a = tid; a < n; a += RES_THREADS) {
            double D = 0.0;
            for (int sl = 0; sl < S; sl++) D = D + h_part[sl * n + a];
            h_rhs[a] = (phase == 1) ? D : 2 * rhos[kk] * D;
        }
        __syncthreads();
        if (phase == 1) {
            for (int a = tid; a < n; a += RES_THREADS) rv.z[a] = h_rhs[a] / m;
        } else {
            for (long long e = P.q_ptr[0] + tid; e < P.q_ptr[1]; e += RES_THREADS) h_rhs[P.q_idx[e]] -= P.q_val[e];   // q_0 indices are distinct
            __syncthreads();
            const double* Zi = Zinv + (size_t)kk * n * n;
            // z = Zinv rhs: a warp owns rows warp, warp + 16, ...; four rows in flight per pass (their loads overlap)
            for (int a = warp; a < n; a += 4 * RES_WARPS) {
                double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int r4 = 0; r4 < 4; r4++) {
                    const int aa = a + r4 * RES_WARPS;
                    if (aa < n) {
                        const double* Za = Zi + (size_t)aa * n;
                        for (int b = lane; b < n; b += 32) acc[r4] = fma(Za[b], h_rhs[b], acc[r4]);
                    }
                }
#pragma unroll
                for (int r4 = 0; r4 < 4; r4++) {
                    const int aa = a + r4 * RES_WARPS;
                    const double v = warp_sum(acc[r4]);
                    if (aa < n && lane == 0) rv.z[aa] = v;
                }
            }
        }
        __syncthreads();
    };
    auto finish_run = [&](const RunView& rv, int run) {
        // x2 = better(x1, bestx); the returned pair is a fresh exact evaluation of that point
        const bool take1 = res_better_first(rv.sc[3], rv.sc[2], rv.sc[5], rv.sc[4]);
        const double* res = take1 ? rv.x1 : rv.bestx;
        for (int a = tid; a < n; a += RES_THREADS) { h_z[a] = res[a]; X[(size_t)run * n + a] = res[a]; }
        __syncthreads();
        double f, mv;
        res_block_eval(P, h_z, h_red, &f, &mv);
        if (tid == 0) {
            f0_out[run] = f; mv_out[run] = mv;
            if (stats) {
                qcqp_admm_stats st;
                st.iters_p1 = rv.ic[4]; st.iters_p2 = rv.ic[5];
                st.onecons_calls = (long long)m * ((long long)rv.ic[4] + (long long)rv.ic[5]);
                st.status = 0; st.pad_ = 0;
                stats[run] = st;
            }
            rv.ic[0] = CMD_IDLE; rv.ic[1] = 3;
        }
        __syncthreads();
    };
    auto enter_phase2 = [&](const RunView& rv, int run) {
        // bestx = x1, z = x1, xs = [x1] * m, us = 0 (qcqp.py:218-222)
        for (int a = tid; a < n; a += RES_THREADS) rv.bestx[a] = rv.x1[a];
        if (tid == 0) { rv.sc[4] = rv.sc[2]; rv.sc[5] = rv.sc[3]; rv.ic[1] = 2; rv.ic[2] = 0; rv.ic[3] = 0; }
        __syncthreads();
        if (prm.num_iters <= 0) { finish_run(rv, run); return; }
        if (tid == 0) { rv.ic[2] = 1; rv.ic[5] = 1; rv.ic[0] = CMD_HOME_P2; }
        next_z(rv, run, 2, true, rv.x1);
    };

    // ---- initial state of the runs this CTA is home of ----
    for (int lr = ci; lr < nrun; lr += m) {
        const int run = run0 + lr;
        const RunView rv = res_run_view(ws, run, n, m);
        const int rr = run % prm.R;
        for (int a = tid; a < n; a += RES_THREADS) { const double v = X0[(size_t)rr * n + a]; rv.x0[a] = v; rv.x1[a] = v; h_z[a] = v; }
        __syncthreads();
        double f, mv;
        res_block_eval(P, h_z, h_red, &f, &mv);
        if (tid == 0) {
            rv.sc[0] = f; rv.sc[1] = mv; rv.sc[2] = f; rv.sc[3] = mv;
            for (int q = 0; q < 8; q++) rv.ic[q] = 0;
        }
        __syncthreads();
        if (prm.phase1 && prm.num_iters > 0 && !(mv < prm.tol)) {
            if (tid == 0) { rv.ic[1] = 1; rv.ic[2] = 1; rv.ic[4] = 1; rv.ic[0] = CMD_HOME_P1; }
            next_z(rv, run, 1, true, rv.x0);
        } else {
            enter_phase2(rv, run);   // x1 = better(x0, z = x0) is x0 itself
        }
    }
    bar_target += m;
    group_barrier(bar, bar_target);

    const bool use_mma = (n & 7) == 0 && !prm.no_mma;
    int par = 0;   // step parity: S1 reads zs[par], S2 writes zs[par ^ 1]
    for (;;) {
        // =========================== S1: my constraint, every active run of my group ===========================
        int any_active = 0;
        for (int b0 = 0; b0 < nrun; b0 += RES_RB) {
            const int nb = (nrun - b0 < RES_RB) ? (nrun - b0) : RES_RB;
            // A: commands, z, v = z + u
            if (tid < RES_RB) {
                int cmd = CMD_IDLE;
                if (tid < nb) cmd = __ldcg(res_run_view(ws, run0 + b0 + tid, n, m).ic);
                sm.flag[tid] = cmd;
                if (tid < nb) sm.cmds[b0 + tid] = cmd;
            }
            __syncthreads();
            int act = 0;
            for (int q = 0; q < nb; q++) act |= (sm.flag[q] != CMD_IDLE) << q;
            if (act == 0) continue;
            any_active = 1;
            for (int e = tid; e < RES_RB * n; e += RES_THREADS) {
                const int q = e / n, a = e - q * n;
                double zv = 0.0, uv = 0.0;
                if ((act >> q) & 1) {
                    const RunView rq = res_run_view(ws, run0 + b0 + q, n, m);
                    const bool home_z = (sm.flag[q] == CMD_HOME_P1 || sm.flag[q] == CMD_HOME_P2);
                    zv = __ldcg((home_z ? rq.z : rq.zs + (size_t)par * n) + a);
                    if (home_z) sm.u[(size_t)(b0 + q) * n + a] = 0.0;   // xs = [x_init] * m, us = 0 at a phase start
                    uv = sm.u[(size_t)(b0 + q) * n + a];
                }
                sm.zs[e] = zv; sm.vs[e] = zv + uv;
                if (use_mma) { sm.zvT[a * (2 * RES_RB) + 2 * q] = zv; sm.zvT[a * (2 * RES_RB) + 2 * q + 1] = zv + uv; }
            }
            __syncthreads();
            // B: [zhat, vhat] = Q^T [z, v] for the whole batch.  n a multiple of 8: FP64 tensor-core tiles -- a warp owns 8 outputs b,
            //    the 8 columns are (z_0, v_0, .., z_3, v_3), k walks a in steps of 4.  Otherwise: thread (slice, b) walks its slice of a.
            if (use_mma) {
                const int gq = lane >> 2, tq = lane & 3;
                for (int mt = warp; mt < (n >> 3); mt += RES_WARPS) {
                    const int bb0 = mt << 3;
                    // four independent accumulator chains (k = a0, a0 + 4, a0 + 8, a0 + 12): the DMMA latency overlaps
                    double c0[4] = {0.0, 0.0, 0.0, 0.0}, c1[4] = {0.0, 0.0, 0.0, 0.0};
                    int a0 = 0;
                    for (; a0 + 16 <= n; a0 += 16) {
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            dmma_m8n8k4(c0[u], c1[u], sm.Q[(a0 + 4 * u + tq) * ldq + bb0 + gq], sm.zvT[(a0 + 4 * u + tq) * (2 * RES_RB) + gq]);
                    }
                    for (; a0 < n; a0 += 4) dmma_m8n8k4(c0[0], c1[0], sm.Q[(a0 + tq) * ldq + bb0 + gq], sm.zvT[(a0 + tq) * (2 * RES_RB) + gq]);
                    sm.zh[tq * n + bb0 + gq] = (c0[0] + c0[1]) + (c0[2] + c0[3]);
                    sm.vh[tq * n + bb0 + gq] = (c1[0] + c1[1]) + (c1[2] + c1[3]);
                }
            } else {
                const int slice = tid / nb32, b = tid - slice * nb32;
                double acc[2 * RES_RB];
#pragma unroll
                for (int q = 0; q < 2 * RES_RB; q++) acc[q] = 0.0;
                if (slice < S && b < n) {
                    const int a0 = (int)((long long)slice * n / S), a1 = (int)((long long)(slice + 1) * n / S);
                    for (int a = a0; a < a1; a++) {
                        const double qv = sm.Q[a * ldq + b];
#pragma unroll
                        for (int q = 0; q < RES_RB; q++) {
                            acc[2 * q] = fma(qv, sm.zs[q * n + a], acc[2 * q]);
                            acc[2 * q + 1] = fma(qv, sm.vs[q * n + a], acc[2 * q + 1]);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 2 * RES_RB; q++) sm.part[((size_t)slice * 2 * RES_RB + q) * n + b] = acc[q];
                }
            }
            __syncthreads();
            if (!use_mma) {
                for (int e = tid; e < 2 * RES_RB * n; e += RES_THREADS) {
                    const int q = e / n, b = e - q * n;
                    double s = 0.0;
                    for (int sl = 0; sl < S; sl++) s += sm.part[((size_t)sl * 2 * RES_RB + q) * n + b];
                    if (q & 1) sm.vh[(q >> 1) * n + b] = s; else sm.zh[(q >> 1) * n + b] = s;
                }
                __syncthreads();
            }
            // C: one warp per run: violation of z, early-out test, multiplier bisection.  Meanwhile two of the idle warps per
            //    run do what the run's home CTA needs from z alone: the objective f0(z) (exact, from the stored form) and
            //    |z - last_z|^2 (qcqp.py:241)
            if (warp < nb && ((act >> warp) & 1)) {
                const double* zh = sm.zh + warp * n;
                const double* vh = sm.vh + warp * n;
                double a1 = 0.0, b1 = 0.0, a2 = 0.0, b2 = 0.0;
                for (int t = lane; t < n; t += 32) {
                    const double l = sm.lam[t], qq = sm.qh[t], zt = zh[t], vt = vh[t];
                    a1 = fma(l, zt * zt, a1); b1 = fma(qq, zt, b1);
                    a2 = fma(l, vt * vt, a2); b2 = fma(qq, vt, b2);
                }
                const double fz = warp_sum(a1) + warp_sum(b1) + rj;
                const double fv = warp_sum(a2) + warp_sum(b2) + rj;
                if (lane == 0) res_run_view(ws, run0 + b0 + warp, n, m).viol[ci] = violation_of(relop, fz);
                int early = 0;
                if (relop == QCQP_RELOP_LE && fv <= 0) early = 1;   // onecons_qcqp returns z + u itself (utilities.py:157-158)
                else if (use_mma) res_bisect(sm.lam, sm.qh, vh, sm.xh + warp, RES_RB, rj, n, lane);
                else res_bisect(sm.lam, sm.qh, vh, sm.xh + warp * n, 1, rj, n, lane);
                if (lane == 0) sm.flag[warp] = early ? -1 : sm.flag[warp];
            } else if (warp >= RES_RB && warp < RES_RB + 2 * nb) {
                const int q = (warp - RES_RB) >> 1, job = (warp - RES_RB) & 1;
                const int lr = b0 + q;
                if (((act >> q) & 1) && (lr % m) == ci) {
                    const double* zq = sm.zs + q * n;
                    if (job == 0) {
                        double fz = 0.0;
                        eval_forms(P, zq, 0, 0, false, lane, [&](int, double v) { fz = v; });
                        fz = warp_sum(fz);
                        if (lane == 0) sm.hs[2 * (lr / m)] = fz;
                    } else {
                        const double* lz = res_run_view(ws, run0 + lr, n, m).last_z;
                        double part = 0.0;
                        for (int a = lane; a < n; a += 32) { const double dd = lz[a] - zq[a]; part = fma(dd, dd, part); }
                        part = warp_sum(part);
                        if (lane == 0) sm.hs[2 * (lr / m) + 1] = part;
                    }
                }
            }
            __syncthreads();
            // D: x = Q xhat for the runs that were projected (tensor-core tiles: a warp owns 8 outputs a, columns = runs, k walks b)
            if (use_mma) {
                const int gq = lane >> 2, tq = lane & 3;
                for (int mt = warp; mt < (n >> 3); mt += RES_WARPS) {
                    const int aa0 = mt << 3;
                    double c0[4] = {0.0, 0.0, 0.0, 0.0}, c1[4] = {0.0, 0.0, 0.0, 0.0};
                    int bb = 0;
                    for (; bb + 16 <= n; bb += 16) {
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            dmma_m8n8k4(c0[u], c1[u], sm.Q[(aa0 + gq) * ldq + bb + 4 * u + tq], (gq < RES_RB) ? sm.xh[(bb + 4 * u + tq) * RES_RB + gq] : 0.0);
                    }
                    for (; bb < n; bb += 4) dmma_m8n8k4(c0[0], c1[0], sm.Q[(aa0 + gq) * ldq + bb + tq], (gq < RES_RB) ? sm.xh[(bb + tq) * RES_RB + gq] : 0.0);
                    if (tq < RES_RB / 2) {
                        sm.part[(2 * tq) * n + aa0 + gq] = (c0[0] + c0[1]) + (c0[2] + c0[3]);
                        sm.part[(2 * tq + 1) * n + aa0 + gq] = (c1[0] + c1[1]) + (c1[2] + c1[3]);
                    }
                }
            } else {
                const int slice = tid / nb32, a = tid - slice * nb32;
                double acc[RES_RB];
#pragma unroll
                for (int q = 0; q < RES_RB; q++) acc[q] = 0.0;
                if (slice < S && a < n) {
                    const int c0 = (int)((long long)slice * n / S), c1 = (int)((long long)(slice + 1) * n / S);
                    for (int b = c0; b < c1; b++) {
                        const double qv = sm.Q[a * ldq + b];
#pragma unroll
                        for (int q = 0; q < RES_RB; q++) acc[q] = fma(qv, sm.xh[q * n + b], acc[q]);
                    }
#pragma unroll
                    for (int q = 0; q < RES_RB; q++) sm.part[((size_t)slice * RES_RB + q) * n + a] = acc[q];
                }
            }
            __syncthreads();
            // E: x_i, u_i += z - x_i, publish d_i = x_i - u_i
            for (int e = tid; e < RES_RB * n; e += RES_THREADS) {
                const int q = e / n, a = e - q * n;
                if (!((act >> q) & 1)) continue;
                double xv;
                if (sm.flag[q] == -1) xv = sm.vs[e];
                else {
                    xv = 0.0;
                    for (int sl = 0; sl < (use_mma ? 1 : S); sl++) xv += sm.part[((size_t)sl * RES_RB + q) * n + a];
                }
                double* up = sm.u + (size_t)(b0 + q) * n + a;
                const double un = *up + (sm.zs[e] - xv);
                *up = un;
                res_run_view(ws, run0 + b0 + q, n, m).d[(size_t)ci * n + a] = xv - un;
            }
            __syncthreads();
        }
        if (!any_active) break;   // every run of my group has finished (same commands seen by all CTAs of the group)
        bar_target += m;
        group_barrier(bar, bar_target);

        // =========================== S2 ===========================
        // (a) every CTA of the group: the next z of every active run under the assumption that the run goes on in its phase.
        //     D = sum_i d_i in full by every CTA (a few overlapped L2 round trips, no extra barrier), then only MY rows
        //     (a = ci, ci + m, ...) of z = D/m or Zinv (2 rho D - q0), written to the other parity's buffer
        for (int b0 = 0; b0 < nrun; b0 += RES_RB) {
            const int nb = (nrun - b0 < RES_RB) ? (nrun - b0) : RES_RB;
            int act = 0;
            for (int q = 0; q < nb; q++) act |= (sm.cmds[b0 + q] != CMD_IDLE) << q;
            if (act == 0) continue;
            double* rhs = sm.part;   // [RB][n]
            {
                const int slice = tid / nb32, a = tid - slice * nb32;
                for (int q = slice; q < nb; q += S) {
                    if (a < n && ((act >> q) & 1)) {
                        const int run = run0 + b0 + q;
                        const double* dq = res_run_view(ws, run, n, m).d + a;
                        double D = 0.0;
                        int i = 0;
                        for (; i + 8 <= m; i += 8) {
                            double dv[8];
#pragma unroll
                            for (int u8 = 0; u8 < 8; u8++) dv[u8] = __ldcg(dq + (size_t)(i + u8) * n);
#pragma unroll
                            for (int u8 = 0; u8 < 8; u8++) D = D + dv[u8];
                        }
                        for (; i < m; i++) D = D + __ldcg(dq + (size_t)i * n);
                        const int cmd = sm.cmds[b0 + q];
                        const bool p2 = (cmd == CMD_HOME_P2 || cmd == CMD_SPEC_P2);
                        rhs[q * n + a] = p2 ? (2 * rhos[run / prm.R] * D - sm.q0s[a]) : D;
                    }
                }
            }
            __syncthreads();
            if (ci < n) {
                const int nrows = (n - ci + m - 1) / m;
                for (int idx = warp; idx < nb * nrows; idx += RES_WARPS) {
                    const int q = idx / nrows, a = ci + (idx - q * nrows) * m;
                    if (!((act >> q) & 1)) continue;
                    const int run = run0 + b0 + q;
                    const int cmd = sm.cmds[b0 + q];
                    double* zo = res_run_view(ws, run, n, m).zs + (size_t)(par ^ 1) * n;
                    if (cmd == CMD_HOME_P2 || cmd == CMD_SPEC_P2) {
                        const double* Za = Zinv + ((size_t)(run / prm.R) * n + a) * n;
                        double acc = 0.0;
                        for (int bb = lane; bb < n; bb += 32) acc = fma(Za[bb], rhs[q * n + bb], acc);
                        acc = warp_sum(acc);
                        if (lane == 0) zo[a] = acc;
                    } else if (lane == 0) {
                        zo[a] = rhs[q * n + a] / m;
                    }
                }
            }
            __syncthreads();
        }
        // (b) the runs I am home of: the reference's loop control on this iteration's z; a run that goes on in its phase takes the
        //     rows computed in (a), a phase start gets its z from the home CTA
        for (int lr = ci; lr < nrun; lr += m) {
            const int run = run0 + lr;
            const RunView rv = res_run_view(ws, run, n, m);
            const int cmd = sm.cmds[lr];
            if (cmd == CMD_IDLE) continue;
            const int phase = rv.ic[1], t = rv.ic[2];
            // objective and step length from S1's side warps, max violation from the m CTAs (rotated forms)
            double fz = sm.hs[2 * (lr / m)], mvz = -QCQP_INF;
            const double step2 = sm.hs[2 * (lr / m) + 1];
            if (warp == 0) {
                for (int i = lane; i < m; i += 32) { const double v = __ldcg(rv.viol + i); mvz = (v > mvz) ? v : mvz; }
                mvz = warp_max(mvz);
                if (lane == 0) h_red[1] = mvz;
            }
            __syncthreads();
            mvz = h_red[1];
            __syncthreads();
            const double* zt = (cmd == CMD_HOME_P1 || cmd == CMD_HOME_P2) ? rv.z : rv.zs + (size_t)par * n;   // this iteration's z
            if (phase == 1) {
                if (t >= prm.num_iters || mvz < prm.tol) {
                    // x1 = better(x0, z) (qcqp.py:280-281)
                    if (!res_better_first(rv.sc[1], rv.sc[0], mvz, fz)) {
                        for (int a = tid; a < n; a += RES_THREADS) rv.x1[a] = __ldcg(zt + a);
                        if (tid == 0) { rv.sc[2] = fz; rv.sc[3] = mvz; }
                    }
                    __syncthreads();
                    enter_phase2(rv, run);
                } else if (tid == 0) {
                    rv.ic[2] = t + 1; rv.ic[4] = rv.ic[4] + 1; rv.ic[0] = CMD_SPEC_P1;
                }
            } else {
                bool stop = false;
                if (rv.ic[3] && sqrt(step2) < prm.tol) stop = true;
                if (!stop) {
                    for (int a = tid; a < n; a += RES_THREADS) rv.last_z[a] = __ldcg(zt + a);
                    if (mvz > prm.viol_lim) stop = true;
                    else if (res_better_first(mvz, fz, rv.sc[5], rv.sc[4])) {
                        for (int a = tid; a < n; a += RES_THREADS) rv.bestx[a] = __ldcg(zt + a);
                        __syncthreads();
                        if (tid == 0) { rv.sc[4] = fz; rv.sc[5] = mvz; }
                    }
                    if (!stop && t >= prm.num_iters) stop = true;
                }
                __syncthreads();
                if (stop) finish_run(rv, run);
                else if (tid == 0) { rv.ic[3] = 1; rv.ic[2] = t + 1; rv.ic[5] = rv.ic[5] + 1; rv.ic[0] = CMD_SPEC_P2; }
            }
            __syncthreads();
        }
        par ^= 1;
        bar_target += m;
        group_barrier(bar, bar_target);
    }
}

// resident variant applies when every constraint's eigenbasis fits one SM's shared memory and a group of m CTAs fits the GPU
bool admm_res_plan(const qcqp_pack* p, int runs, ResK* k, size_t* smem_bytes)
{
    const PackView& v = p->v;
    const int sms = num_sms(p->device);
    if (v.m <= 0 || v.m > sms || runs <= 0) return false;
    int G = sms / v.m;
    if (G > runs) G = runs;
    if (const char* fg = getenv("QCQP_ADMM_GROUPS")) { const int f = atoi(fg); if (f >= 1 && f <= G) G = f; }   // A/B runs
    int rpg = (runs + G - 1) / G;
    G = (runs + rpg - 1) / rpg;
    const int nb32 = (v.n + 31) / 32 * 32;
    if (nb32 > RES_THREADS || v.n > 32 * RES_EPL) return false;
    const int S = RES_THREADS / nb32;
    const bool mma = (v.n & 7) == 0 && !getenv("QCQP_ADMM_NO_MMA");
    const int ldq = mma ? v.n + ((4 - (v.n & 15) + 16) & 15) : v.n + 1;
    size_t doubles = (size_t)v.n * ldq + 2 * (size_t)v.n + (size_t)rpg * v.n + 7 * (size_t)RES_RB * v.n +
                     (size_t)S * 2 * RES_RB * v.n + 2 * (size_t)v.n + RES_WARPS + 8 + 2 * (size_t)((rpg + v.m - 1) / v.m) + 2 + (size_t)v.n + (size_t)(rpg + 1) / 2 + 1;
    size_t bytes = doubles * 8 + 64;
    if (bytes > (size_t)max_smem_optin(p->device)) return false;
    k->no_mma = getenv("QCQP_ADMM_NO_MMA") ? 1 : 0;
    k->ldq = ldq;
    k->G = G; k->rpg = rpg; k->S = S; k->nb32 = nb32; k->runs = runs;
    *smem_bytes = bytes;
    return true;
}

int admm_res_launch(qcqp_pack* p, const qcqp_admm_params* prm, const ResK& plan, size_t smem, const double* drhos, const double* dZinv, int K,
                    const double* dX0, int R, double* dX, double* df0, double* dmv, qcqp_admm_stats* dstats, cudaStream_t stream)
{
    const PackView& v = p->v;
    ResK k = plan;
    k.num_iters = prm->num_iters; k.viol_lim = prm->viol_lim; k.tol = prm->tol; k.phase1 = prm->phase1; k.K = K; k.R = R;
    const size_t ws_doubles = (size_t)k.runs * res_run_doubles(v.n, v.m);
    const size_t bar_off = (ws_doubles * 8 + 255) & ~(size_t)255;
    int rc = ensure_workspace(p, bar_off + (size_t)k.G * 4 + 256);
    if (rc != QCQP_OK) return rc;
    double* ws = (double*)p->ws;
    unsigned* bars = (unsigned*)((char*)p->ws + bar_off);
    QCQP_CUDA_TRY(cudaMemsetAsync(bars, 0, (size_t)k.G * 4, stream));
    QCQP_CUDA_TRY(cudaFuncSetAttribute(admm_res_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PackView pv = v;
    void* args[] = {(void*)&pv, (void*)&k, (void*)&drhos, (void*)&dZinv, (void*)&dX0, (void*)&dX, (void*)&df0, (void*)&dmv, (void*)&dstats,
                    (void*)&ws, (void*)&bars};
    QCQP_CUDA_TRY(cudaLaunchCooperativeKernel((void*)admm_res_kernel, dim3(v.m * k.G), dim3(RES_THREADS), args, smem, stream));
    return QCQP_OK;
}

// tries the resident kernel; *used = false when the problem does not fit it (admm.cu then runs one CTA per run)
int admm_res_try(qcqp_pack* p, const qcqp_admm_params* prm, const double* drhos, const double* dZinv, int K, const double* dX0, int R, double* dX,
                 double* df0, double* dmv, qcqp_admm_stats* dstats, cudaStream_t stream, bool* used)
{
    *used = false;
    ResK plan;
    size_t smem = 0;
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, p->device);
    if (!coop || !admm_res_plan(p, K * R, &plan, &smem)) return QCQP_OK;
    // the group barriers need every CTA resident: check the grid against the occupancy of this kernel on this device
    int per_sm = 0;
    if (cudaFuncSetAttribute(admm_res_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, admm_res_kernel, RES_THREADS, smem) != cudaSuccess ||
        (long long)per_sm * num_sms(p->device) < (long long)p->v.m * plan.G) {
        cudaGetLastError();
        return QCQP_OK;
    }
    *used = true;
    return admm_res_launch(p, prm, plan, smem, drhos, dZinv, K, dX0, R, dX, df0, dmv, dstats, stream);
}

}  // namespace qcqp
