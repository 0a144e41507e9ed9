// cd_lpc.cu -- coordinate descent for SEPARABLE problems: every constraint touches exactly one coordinate and every
// coordinate has exactly one constraint (Boolean least squares, MAXCUT: x_k^2 = 1; any per-coordinate quadratic constraint).
// Same algorithm and results as cd.cu (improve_coord_descent, qcqp.py:181-192), different parallel decomposition:
// ONE WARP PER RESTART, ONE LANE PER COORDINATE, 32 consecutive coordinates per pass.
//
// Why that is legal.  For a separable problem the one-variable restriction of coordinate k's constraint is the constant
// triple (P_j[k,k], q_j[k], r_j) -- t0 = f_j(z) = r_j exactly, as get_onevar_func computes it (utilities.py:99-105).
//  * Phase 1 (qcqp.py:101-148) ignores the objective, so the steps of different coordinates do not interact at all except
//    through (i) the order in which they consume the MT19937 stream and (ii) update_counter.  Whether a bisection probe is
//    feasible does not depend on the random numbers, so each lane runs its whole bisection draw-free, counting the words
//    the reference would consume (choice: one word iff two pieces; uniform: two); a warp scan turns the counts into
//    stream offsets and every lane reads exactly the words of its LAST feasible probe -- the only draw that survives.
//  * Phase 2 (qcqp.py:152-178): a step that does not move x_k changes nothing (x, f_0(x), g = P_0 x, the RNG), so 32 steps
//    are evaluated speculatively from the same state; the steps before the first one that moves (or needs a random
//    number) are committed wholesale, that one is applied (g += delta * column k: the only time a row of P_0 is read), and
//    the pass restarts after it.  On Boolean LS ~3% of the steps move.
// The objective's row dot comes from the cached g = P_0 x: (P_0 z)_k = g_k - P_0[k,k] x_k.
#include <cstdlib>

#include "cd_shared.cuh"
#include "common.cuh"
#include "onevar.cuh"

namespace qcqp {

__device__ __forceinline__ uint32_t mt_temper(uint32_t y)
{
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// MT19937 state regeneration by the whole warp: chunks of 32 indices in order, each chunk read-then-write, which is
// exactly the sequential recurrence (an index only ever needs older values at higher indices and newer ones >= 227 below).
__device__ __forceinline__ void mt_refill_warp(uint32_t* mt, int lane)
{
    for (int base = 0; base < 624; base += 32) {
        const int i = base + lane;
        uint32_t nv = 0;
        if (i < 624) {
            const uint32_t a = mt[i], b = mt[(i + 1 == 624) ? 0 : i + 1], c = mt[(i + 397 >= 624) ? i + 397 - 624 : i + 397];
            const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
            nv = c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        __syncwarp();
        if (i < 624) mt[i] = nv;
        __syncwarp();
    }
}

struct LpcMem {
    double* x;
    double* g;
    uint32_t* mt;
    double* blk;   // [32][32] diagonal block of P_0 for the pass in flight (dense objective only)
};

// g = P_0 x and f_0(x) = x.g + q_0.x + r_0 from scratch
__device__ __forceinline__ double lpc_refresh(const PackView& P, const LpcView& V, const LpcMem& w, int lane)
{
    const int n = P.n;
    if (V.obj_dense) {
        const int ld = P.ld;
        const double* M = P.dense_P;   // the objective is dense slot 0
        const int n2 = (n + 1) >> 1;
        for (int c0 = 0; c0 < n2; c0 += 128) {
            double2 a0 = make_double2(0.0, 0.0), a1 = a0, a2 = a0, a3 = a0;
            const int c = c0 + lane;
            const bool v0 = c < n2, v1 = c + 32 < n2, v2 = c + 64 < n2, v3 = c + 96 < n2;
            const double2 zz = make_double2(0.0, 0.0);
            int r = 0;
            // 8 rows x 4 column groups = 32 sixteen-byte loads in flight per lane: the refresh is a latency-bound GEMV over the
            // whole of P_0 by one warp, and the restarts that need it (>= 16 sweeps) are the ones that end the launch
            for (; r + 8 <= n; r += 8) {
                double2 t[8][4];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const double2* row = reinterpret_cast<const double2*>(M + (size_t)(r + u) * ld);
                    t[u][0] = v0 ? __ldg(&row[c]) : zz; t[u][1] = v1 ? __ldg(&row[c + 32]) : zz;
                    t[u][2] = v2 ? __ldg(&row[c + 64]) : zz; t[u][3] = v3 ? __ldg(&row[c + 96]) : zz;
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const double xr = w.x[r + u];
                    a0.x = fma(t[u][0].x, xr, a0.x); a0.y = fma(t[u][0].y, xr, a0.y);
                    a1.x = fma(t[u][1].x, xr, a1.x); a1.y = fma(t[u][1].y, xr, a1.y);
                    a2.x = fma(t[u][2].x, xr, a2.x); a2.y = fma(t[u][2].y, xr, a2.y);
                    a3.x = fma(t[u][3].x, xr, a3.x); a3.y = fma(t[u][3].y, xr, a3.y);
                }
            }
            for (; r + 4 <= n; r += 4) {
                double2 t[4][4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const double2* row = reinterpret_cast<const double2*>(M + (size_t)(r + u) * ld);
                    t[u][0] = v0 ? __ldg(&row[c]) : zz; t[u][1] = v1 ? __ldg(&row[c + 32]) : zz;
                    t[u][2] = v2 ? __ldg(&row[c + 64]) : zz; t[u][3] = v3 ? __ldg(&row[c + 96]) : zz;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const double xr = w.x[r + u];
                    a0.x = fma(t[u][0].x, xr, a0.x); a0.y = fma(t[u][0].y, xr, a0.y);
                    a1.x = fma(t[u][1].x, xr, a1.x); a1.y = fma(t[u][1].y, xr, a1.y);
                    a2.x = fma(t[u][2].x, xr, a2.x); a2.y = fma(t[u][2].y, xr, a2.y);
                    a3.x = fma(t[u][3].x, xr, a3.x); a3.y = fma(t[u][3].y, xr, a3.y);
                }
            }
            for (; r < n; r++) {
                const double xr = w.x[r];
                const double2* row = reinterpret_cast<const double2*>(M + (size_t)r * ld);
                if (v0) { const double2 t = row[c]; a0.x = fma(t.x, xr, a0.x); a0.y = fma(t.y, xr, a0.y); }
                if (v1) { const double2 t = row[c + 32]; a1.x = fma(t.x, xr, a1.x); a1.y = fma(t.y, xr, a1.y); }
                if (v2) { const double2 t = row[c + 64]; a2.x = fma(t.x, xr, a2.x); a2.y = fma(t.y, xr, a2.y); }
                if (v3) { const double2 t = row[c + 96]; a3.x = fma(t.x, xr, a3.x); a3.y = fma(t.y, xr, a3.y); }
            }
            double2* g2 = reinterpret_cast<double2*>(w.g);
            if (v0) g2[c] = a0;
            if (v1) g2[c + 32] = a1;
            if (v2) g2[c + 64] = a2;
            if (v3) g2[c + 96] = a3;
        }
    } else {
        for (int k = lane; k < n; k += 32) {
            double s = V.o_diag[k] * w.x[k];
            const int rb = V.o_rbeg[k], rl = V.o_rlen[k];
            for (int t = rb; t < rb + rl; t++) s = fma(P.row_val[t], w.x[P.row_col[t]], s);
            w.g[k] = s;
        }
    }
    __syncwarp();
    double acc = 0.0;
    for (int k = lane; k < n; k += 32) acc = fma(w.x[k], w.g[k] + V.o_q[k], acc);
    return warp_sum(acc) + V.o_r;
}

// max(prob.violations(x)) -- each constraint is (p x_k + q) x_k + r, the reference's own operation order
__device__ __forceinline__ double lpc_max_violation(const PackView& P, const LpcView& V, const LpcMem& w, int lane)
{
    double mv = -QCQP_INF;
    for (int k = lane; k < P.n; k += 32) {
        const double v = violation_of(V.c_rel[k], onevar_eval(V.c_p[k], V.c_q[k], V.c_r[k], w.x[k]));
        mv = (v > mv) ? v : mv;
    }
    return warp_max(mv);
}

// after x_k += delta: g += delta * (column k of P_0)
__device__ __forceinline__ void lpc_axpy(const PackView& P, const LpcView& V, const LpcMem& w, int k, double delta, int lane)
{
    if (V.obj_dense) {
        const int n2 = (P.n + 1) >> 1;
        const double2* row = reinterpret_cast<const double2*>(P.dense_P + (size_t)k * P.ld);   // symmetric: row k == column k
        double2* g2 = reinterpret_cast<double2*>(w.g);
        // the row comes from L2 once per move: put 16 independent 16-byte loads in flight (a whole row for n <= 1024)
        // before the first use, so a move costs one L2 round trip
        for (int c0 = lane; c0 < n2; c0 += 512) {
            double2 rv[16];
#pragma unroll
            for (int u = 0; u < 16; u++) rv[u] = (c0 + 32 * u < n2) ? __ldg(&row[c0 + 32 * u]) : make_double2(0.0, 0.0);
#pragma unroll
            for (int u = 0; u < 16; u++) {
                const int c = c0 + 32 * u;
                if (c < n2) {
                    double2 gv = g2[c];
                    gv.x = fma(rv[u].x, delta, gv.x);
                    gv.y = fma(rv[u].y, delta, gv.y);
                    g2[c] = gv;
                }
            }
        }
    } else {
        if (lane == 0) w.g[k] = fma(V.o_diag[k], delta, w.g[k]);
        const int rb = V.o_rbeg[k], rl = V.o_rlen[k];
        for (int t = rb + lane; t < rb + rl; t += 32) {
            const int c = P.row_col[t];
            w.g[c] = fma(P.row_val[t], delta, w.g[c]);   // distinct columns within a row: no conflicts
        }
    }
    __syncwarp();
}

// Dense objective, end of a 32-coordinate pass: g += sum over the coordinates that moved (bit m of `moved`: x_{k0+m} += the
// delta held by lane m) of delta * row of P_0.  Two whole rows (2 x 16 sixteen-byte loads per lane for n <= 1024) are in
// flight at a time, so a pass pays one L2 round trip per pair of moves instead of one per move and column chunk.
__device__ __forceinline__ void lpc_axpy_multi(const PackView& P, const LpcMem& w, int k0, unsigned moved, double mydelta, int lane)
{
    const int n2 = (P.n + 1) >> 1;
    double2* g2 = reinterpret_cast<double2*>(w.g);
    const double2 zz = make_double2(0.0, 0.0);
    unsigned mm = moved;
    while (mm) {
        const int m0 = __ffs(mm) - 1;
        mm &= mm - 1;
        const int m1 = mm ? (__ffs(mm) - 1) : -1;
        if (mm) mm &= mm - 1;
        const double d0 = __shfl_sync(FULL, mydelta, m0);
        const double d1 = __shfl_sync(FULL, mydelta, m1 < 0 ? 0 : m1);
        const double2* r0 = reinterpret_cast<const double2*>(P.dense_P + (size_t)(k0 + m0) * P.ld);
        const double2* r1 = reinterpret_cast<const double2*>(P.dense_P + (size_t)(k0 + (m1 < 0 ? m0 : m1)) * P.ld);
        for (int cb = 0; cb < n2; cb += 512) {
            double2 a[16], b[16];
#pragma unroll
            for (int u = 0; u < 16; u++) {
                const int c = cb + lane + 32 * u;
                a[u] = (c < n2) ? __ldg(&r0[c]) : zz;
                b[u] = (c < n2 && m1 >= 0) ? __ldg(&r1[c]) : zz;
            }
#pragma unroll
            for (int u = 0; u < 16; u++) {
                const int c = cb + lane + 32 * u;
                if (c < n2) {
                    double2 gv = g2[c];
                    gv.x = fma(a[u].x, d0, gv.x); gv.y = fma(a[u].y, d0, gv.y);
                    gv.x = fma(b[u].x, d1, gv.x); gv.y = fma(b[u].y, d1, gv.y);
                    g2[c] = gv;
                }
            }
        }
    }
    __syncwarp();
}

// Draw-free bisection of ONE coordinate of phase 1 (qcqp.py:113-131), computed ahead of the sweep by lpc_p1_pre_kernel: whether a
// bisection probe is feasible depends on x_k alone (separable constraints, no objective in phase 1), and in the FIRST sweep x_k is
// still the start value when its step comes, so the whole sweep's bisections are independent of each other: one thread per
// (restart, coordinate) at full occupancy instead of one warp per restart walking them 32 at a time.
struct LpcPre {
    double l0, h0, l1, h1;   // pieces of the last feasible probe
    int words;               // MT19937 words the reference consumes on this coordinate
    int flags;               // lastn | moved << 4 | hard << 5
};

__global__ void __launch_bounds__(256) lpc_p1_pre_kernel(PackView P, LpcView V, CdK prm, const double* __restrict__ X0, int R, LpcPre* __restrict__ pre)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int n = P.n;
    if (idx >= (size_t)R * n) return;
    const int k = (int)(idx % n);
    const double p = V.c_p[k], q = V.c_q[k], r = V.c_r[k], xk = X0[idx];
    const int rel = V.c_rel[k];
    const double tol = prm.tol, viol_tol = prm.viol_tol;
    const double viol = violation_of(rel, onevar_eval(p, q, r, xk));
    double ss = -tol, es = viol - viol_tol;
    LpcPre e;
    e.l0 = e.h0 = e.l1 = e.h1 = 0.0; e.words = 0;
    int lastn = 0;
    bool moved = false, hard = false;
    while (es - ss > tol) {
        const double s = (ss + es) / 2;
        double a0, b0, a1, b1;
        const int nC = single_constraint_pieces(p, q, r, rel, s, &a0, &b0, &a1, &b1);
        if (nC > 0) {
            e.words += (nC == 2) ? 3 : 2;
            lastn = nC; e.l0 = a0; e.h0 = b0; e.l1 = a1; e.h1 = b1;
            moved = true;
            es = s;
            if (is_inf(a0) || is_inf(b0) || (nC == 2 && (is_inf(a1) || is_inf(b1))) || nC > 2) hard = true;
        } else ss = s;
    }
    e.flags = (lastn & 15) | (moved ? 16 : 0) | (hard ? 32 : 0);
    pre[idx] = e;
}

// stage 0: the whole of improve_coord_descent in one launch.
// stage 1: phase 1 only (x, stream and stats are written back).   stage 2: phase 2 only, starting from stage 1's output with
// g = P_0 x supplied in G (one tiled GEMM for all restarts instead of one latency-bound GEMV per warp); the returned
// (f0, maxviol) then come from the batched eval kernels.
__global__ void __launch_bounds__(32, 7) cd_lpc_kernel(PackView P, LpcView V, CdK prm, int stage, const double* __restrict__ X0, int R,
                                                     qcqp_rng_state* rngs, double* X, const double* __restrict__ G,
                                                     double* __restrict__ f0_out, double* __restrict__ mv_out, qcqp_cd_stats* stats_out,
                                                     const LpcPre* __restrict__ pre)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x;
    const int n = P.n;
    const int npad = (n + 1) & ~1;
    const size_t rr = blockIdx.x;
    if ((int)rr >= R) return;
    LpcMem w;
    w.x = reinterpret_cast<double*>(smem);
    w.g = w.x + npad;
    w.mt = reinterpret_cast<uint32_t*>(w.g + npad);
    w.blk = reinterpret_cast<double*>(w.mt + 624);

    {
        const double* src = (stage == 2) ? X : X0;
        for (int i = lane; i < n; i += 32) w.x[i] = src[rr * n + i];
    }
    if (lane == 0 && npad > n) { w.x[n] = 0.0; w.g[n] = 0.0; }
    for (int i = lane; i < 624; i += 32) w.mt[i] = rngs[rr].key[i];
    int pos = rngs[rr].pos;     // warp-uniform stream position
    __syncwarp();

    qcqp_cd_stats st;
    st.steps_p1 = st.steps_p2 = st.updates_p1 = st.updates_p2 = st.steps_skipped = 0;
    st.sweeps_p1 = st.sweeps_p2 = 0; st.status = QCQP_RUN_OK; st.ran_phase2 = 0;
    const double tol = prm.tol, viol_tol = prm.viol_tol;
    bool dead = false;
    if (stage == 2) {
        st = stats_out[rr];
        dead = st.status != QCQP_RUN_OK;
    }

    // =========================================== phase 1 ===========================================
    if (prm.phase1 && stage != 2) {
        long long uc = 0;
        double viol_last = QCQP_INF;
        for (int t = 0; t < prm.num_iters && !dead; t++) {
            if (viol_last < viol_tol) break;
            st.sweeps_p1++;
            bool skip = false;
            const long long upd_before = st.updates_p1;
            for (int k0 = 0; k0 < n && !skip && !dead;) {
                const int B = (n - k0 < 32) ? (n - k0) : 32;
                const bool act = lane < B;
                const int k = k0 + lane;
                double p = 0.0, q = 0.0, r = 0.0, xk = 0.0;
                int rel = QCQP_RELOP_LE;
                if (act) { p = V.c_p[k]; q = V.c_q[k]; r = V.c_r[k]; rel = V.c_rel[k]; xk = w.x[k]; }
                // ---- draw-free bisection of this lane's coordinate (qcqp.py:113-131) ----
                const double viol = violation_of(rel, onevar_eval(p, q, r, xk));
                double ss = -tol, es = viol - viol_tol;
                int words = 0, lastn = 0;
                double l0 = 0.0, h0 = 0.0, l1 = 0.0, h1 = 0.0;
                bool moved = false, hard = false;
                if (pre && t == 0) {
                    // first sweep: the bisections were computed ahead, one thread per coordinate (lpc_p1_pre_kernel)
                    if (act) {
                        const LpcPre e = pre[rr * n + k];
                        words = e.words; lastn = e.flags & 15; moved = (e.flags & 16) != 0; hard = (e.flags & 32) != 0;
                        l0 = e.l0; h0 = e.h0; l1 = e.l1; h1 = e.h1;
                    }
                } else if (act) {
                    while (es - ss > tol) {
                        const double s = (ss + es) / 2;
                        double a0, b0, a1, b1;
                        const int nC = single_constraint_pieces(p, q, r, rel, s, &a0, &b0, &a1, &b1);
                        if (nC > 0) {
                            words += (nC == 2) ? 3 : 2;     // choice(2) is exactly one masked word; uniform is two
                            lastn = nC; l0 = a0; h0 = b0; l1 = a1; h1 = b1;
                            moved = true;
                            es = s;
                            // an infinite bound makes np.random.uniform raise, possibly depending on the drawn piece:
                            // such a coordinate is replayed with the real stream below
                            if (is_inf(a0) || is_inf(b0) || (nC == 2 && (is_inf(a1) || is_inf(b1))) || nC > 2) hard = true;
                        } else ss = s;
                    }
                }
                // ---- update_counter in sequence order (qcqp.py:132-141) ----
                const unsigned movedmask = __ballot_sync(FULL, act && moved);
                const unsigned hardmask = __ballot_sync(FULL, act && hard);
                const unsigned le = (lane == 31) ? FULL : ((2u << lane) - 1u);
                const unsigned mm = movedmask & le;
                const long long cval = moved ? 0 : (mm ? (long long)(lane - (31 - __clz(mm))) : uc + lane + 1);
                const unsigned hit = __ballot_sync(FULL, act && !moved && cval == n);
                int cut = B - 1;                                  // last lane committed in this pass
                if (hit) cut = __ffs(hit) - 1;
                const int firsthard = hardmask ? (__ffs(hardmask) - 1) : 32;
                const bool do_hard = firsthard <= cut;
                if (do_hard) cut = firsthard - 1;
                const bool mine = act && lane <= cut;
                // ---- stream offsets: words consumed by the committed coordinates before me ----
                const int wmine = (mine && moved) ? words : 0;
                int incl = wmine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
                int rem_total = __shfl_sync(FULL, incl, 31);
                const int cntlast = (lastn == 2) ? 3 : 2;
                int need = (incl - wmine) + wmine - cntlast;      // relative position of my last probe's first word
                uint32_t wd0 = 0, wd1 = 0, wd2 = 0;
                int got = (wmine > 0) ? 0 : cntlast;              // words already fetched
                for (;;) {
                    const int avail = 624 - pos;
                    while (got < cntlast && need + got < avail) {
                        const uint32_t y = mt_temper(w.mt[pos + need + got]);
                        if (got == 0) wd0 = y; else if (got == 1) wd1 = y; else wd2 = y;
                        got++;
                    }
                    if (rem_total <= avail) { pos += rem_total; break; }
                    __syncwarp();
                    mt_refill_warp(w.mt, lane);
                    need -= avail; rem_total -= avail; pos = 0;
                }
                if (wmine > 0) {
                    // np.random.uniform(*C[np.random.choice(len(C))])   (utilities.py:266-267)
                    const int idx = (lastn == 2) ? (int)(wd0 & 1u) : 0;
                    const uint32_t ua = ((lastn == 2) ? wd1 : wd0) >> 5, ub = ((lastn == 2) ? wd2 : wd1) >> 6;
                    const double dd = ((double)ua * 67108864.0 + (double)ub) / 9007199254740992.0;
                    const double lo = idx ? l1 : l0, hi = idx ? h1 : h0;
                    const double range = hi - lo;
                    w.x[k] = lo + range * dd;
                }
                __syncwarp();
                if (cut >= 0) {
                    uc = __shfl_sync(FULL, cval, cut);
                    st.steps_p1 += cut + 1;
                    const unsigned cm = (cut == 31) ? FULL : ((2u << cut) - 1u);
                    st.updates_p1 += __popc(movedmask & cm);
                    if (hit && cut == __ffs(hit) - 1) skip = true;      // 'failed': the rest of the sweep is not executed
                }
                k0 += cut + 1;
                if (do_hard && !skip) {
                    // ---- one coordinate with the real stream (lane 0 draws), exactly the reference's loop ----
                    const double hp = bcast(p, firsthard), hq = bcast(q, firsthard), hr = bcast(r, firsthard), hx = bcast(xk, firsthard);
                    const int hrel = bcast_i(rel, firsthard);
                    const double hv = violation_of(hrel, onevar_eval(hp, hq, hr, hx));
                    double s2 = -tol, e2 = hv - viol_tol, nx = hx;
                    bool mv2 = false;
                    int err = 0;
                    st.steps_p1++;
                    while (e2 - s2 > tol && !err) {
                        const double s = (s2 + e2) / 2;
                        double cl[2], ch[2];
                        const int nC = single_constraint_pieces(hp, hq, hr, hrel, s, &cl[0], &ch[0], &cl[1], &ch[1]);
                        if (nC > 0) {
                            double xv = 0.0;
                            if (lane == 0) {
                                MtRng rng;
                                rng.key = w.mt; rng.pos = pos;
                                choose_point(0.0, 0.0, 0.0, cl, ch, nC, rng, &xv, &err);
                                pos = rng.pos;
                            }
                            pos = bcast_i(pos, 0); err = bcast_i(err, 0); nx = bcast(xv, 0);
                            mv2 = true; e2 = s;
                        } else s2 = s;
                    }
                    if (err) { st.status = err; dead = true; }
                    else if (mv2) { if (lane == 0) w.x[k0] = nx; uc = 0; st.updates_p1++; }
                    else { uc++; if (uc == n) skip = true; }
                    __syncwarp();
                    k0 += 1;
                }
            }
            if (dead) break;
            viol_last = lpc_max_violation(P, V, w, lane);      // qcqp.py:142
            if (!skip && st.updates_p1 == upd_before && !(viol_last < viol_tol) && t + 1 < prm.num_iters) {
                // fixed point: every remaining iteration of qcqp.py:110 repeats this no-op sweep (see cd.cu)
                st.steps_skipped += (long long)(prm.num_iters - (t + 1)) * n;
                break;
            }
        }
    }

    if (stage == 1) {
        for (int i = lane; i < n; i += 32) X[rr * n + i] = w.x[i];
        for (int i = lane; i < 624; i += 32) rngs[rr].key[i] = w.mt[i];
        if (lane == 0) { rngs[rr].pos = pos; stats_out[rr] = st; }
        return;
    }

    // =========================================== phase 2 ===========================================
    double mv = lpc_max_violation(P, V, w, lane);              // improve_coord_descent's gate (qcqp.py:189)
    if (!dead && mv < viol_tol) {
        st.ran_phase2 = 1;
        const double viol_p2 = mv;                               // frozen (qcqp.py:157)
        double f0val;
        if (stage == 2) {
            // g = P_0 x of this restart, computed for all restarts by one GEMM; f_0(x) = x.g + q_0.x + r_0
            const double* gr = G + rr * (size_t)npad;
            double acc = 0.0;
            for (int k = lane; k < n; k += 32) { const double gk = gr[k]; w.g[k] = gk; acc = fma(w.x[k], gk + V.o_q[k], acc); }
            f0val = warp_sum(acc) + V.o_r;
            __syncwarp();
        } else {
            f0val = lpc_refresh(P, V, w, lane);
        }
        int uc = 0;                 // update_counter of phase 2 (qcqp.py:160): never exceeds n
        bool done = false;
        // per-lane memo of the constraint's pieces at the frozen level
        double mp = 0.0, mq = 0.0, mr = 0.0, ml0 = 0.0, mh0 = 0.0, ml1 = 0.0, mh1 = 0.0;
        int mrel = -1, mnC = 0;
        bool mfin = false;      // every endpoint of the memoised pieces is finite
        for (int t = 0; t < prm.num_iters && !done; t++) {
            st.sweeps_p2++;
            // ---- dense objective: blocked Gauss-Seidel.  All moves inside a 32-coordinate pass are resolved in registers with
            // the 32 x 32 diagonal block of P_0 (one coalesced fetch per pass); the rest of g is brought up to date once, after
            // the pass, with the rows of every coordinate that moved in flight together. ----
            // the per-coordinate constants of a pass (constraint triple, relop, P_0[k,k], q_0[k]) are loaded one pass ahead: they
            // come from L2 (the rows of P_0 stream through L1) and would otherwise open every pass with a full round trip
            double n_p = 0.0, n_q = 0.0, n_r = 0.0, n_od = 0.0, n_oq = 0.0;
            int n_rel = 0;
            if (V.obj_dense && lane < n) { n_p = V.c_p[lane]; n_q = V.c_q[lane]; n_r = V.c_r[lane]; n_rel = V.c_rel[lane]; n_od = V.o_diag[lane]; n_oq = V.o_q[lane]; }
            for (int k0 = 0; V.obj_dense && k0 < n && !done; k0 += 32) {
                const int B = (n - k0 < 32) ? (n - k0) : 32;
                const bool act = lane < B;
                const int k = k0 + lane;
                double xk = 0.0, p0 = 0.0, oq = 0.0, gl = 0.0, q0 = 0.0, r0 = f0val, xi = 0.0, mydelta = 0.0;
                int rc = 0;
                const double c_pk = n_p, c_qk = n_q, c_rk = n_r, c_odk = n_od, c_oqk = n_oq;
                const int c_relk = n_rel;
                {
                    int kn = k + 32;
                    if (kn >= n) kn = lane;                      // the next sweep starts over at coordinate `lane`
                    if (kn < n) { n_p = V.c_p[kn]; n_q = V.c_q[kn]; n_r = V.c_r[kn]; n_rel = V.c_rel[kn]; n_od = V.o_diag[kn]; n_oq = V.o_q[kn]; }
                }
                // diagonal block D[i][j] = P_0[k0+i][k0+j]: iteration i loads row i coalesced (lane = column).  The loads are issued
                // before the lanes evaluate their coordinates (85 % of the passes have a candidate mover and need the block), so
                // the round trip overlaps that evaluation; the block is parked in shared memory only if a candidate exists.
                double dv[32];
                {
                    const double* pp = P.dense_P + (size_t)k0 * P.ld + k0 + lane;
                    if (B == 32) {
                        // full block (all passes but the last of a sweep): no per-element predicates, one pointer bump per row
#pragma unroll
                        for (int i = 0; i < 32; i++) { dv[i] = __ldg(pp); pp += P.ld; }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; i++) dv[i] = (i < B && act) ? __ldg(pp + (size_t)i * P.ld) : 0.0;
                    }
                }
                if (act) {
                    const double p = c_pk, q = c_qk, r = c_rk;
                    const int rel = c_relk;
                    xk = w.x[k];
                    if (!(mrel == rel && mp == p && mq == q && mr == r)) {
                        mnC = single_constraint_pieces(p, q, r, rel, viol_p2, &ml0, &mh0, &ml1, &mh1);
                        mfin = mnC > 0 && mnC <= 2 && !is_inf(ml0) && !is_inf(mh0) && (mnC < 2 || (!is_inf(ml1) && !is_inf(mh1)));
                        mp = p; mq = q; mr = r; mrel = rel;
                    }
                    p0 = c_odk; oq = c_oqk; gl = w.g[k];
                    q0 = 2 * (gl - p0 * xk) + oq;
                    r0 = f0val - xk * (p0 * xk + q0);
                    rc = mfin ? choose_point_det_t<true>(p0, q0, r0, ml0, mh0, ml1, mh1, mnC, &xi) : choose_point_det(p0, q0, r0, ml0, mh0, ml1, mh1, mnC, &xi);
                }
                if (__ballot_sync(FULL, act && (rc == 2 || (rc == 1 && fabs(xi - xk) > tol)))) {
#pragma unroll
                    for (int i = 0; i < 32; i++) w.blk[i * 32 + lane] = dv[i];
                }
                __syncwarp();
                unsigned moved = 0;
                int cur = 0;            // next coordinate of the pass to resolve
                while (cur < B) {
                    const bool pending = act && lane >= cur;
                    const bool wants_move = pending && rc == 1 && fabs(xi - xk) > tol;
                    const unsigned stop = __ballot_sync(FULL, wants_move || (pending && rc == 2));
                    const int first = stop ? (__ffs(stop) - 1) : B;
                    const int quiet = first - cur;      // steps that change nothing (qcqp.py:172-176)
                    if (n - uc <= quiet) { st.steps_p2 += (n - uc); done = true; break; }
                    uc += quiet;
                    st.steps_p2 += quiet;
                    if (first == B) break;
                    const int frc = bcast_i(rc, first);
                    st.steps_p2++;
                    bool found = true;
                    double fx = 0.0, fxi = 0.0, fp0 = 0.0, fq0 = 0.0, fr0 = 0.0;
                    if (frc == 2) {
                        fx = bcast(xk, first); fp0 = bcast(p0, first); fq0 = bcast(q0, first); fr0 = bcast(r0, first);
                        double cl[2], ch[2];
                        cl[0] = bcast(ml0, first); ch[0] = bcast(mh0, first); cl[1] = bcast(ml1, first); ch[1] = bcast(mh1, first);
                        const int nC = bcast_i(mnC, first);
                        int err = 0, fnd = 0;
                        double xv = 0.0;
                        if (lane == 0) {
                            MtRng rng;
                            rng.key = w.mt; rng.pos = pos;
                            fnd = choose_point(fp0, fq0, fr0, cl, ch, nC, rng, &xv, &err);
                            pos = rng.pos;
                        }
                        pos = bcast_i(pos, 0); err = bcast_i(err, 0); fnd = bcast_i(fnd, 0); fxi = bcast(xv, 0);
                        if (err) { st.status = err; dead = true; done = true; break; }
                        found = fnd != 0;
                    }
                    // frc == 1: the lane was picked because |xi - xk| > tol; it forms delta and f_0(x) = t0 + b (t2 b + t1) from its own
                    // registers and only those two values travel.  frc == 2 (random draw, rare): everything was broadcast above.
                    bool moves = true;
                    double delta, f0new;
                    if (frc == 2) {
                        moves = found && fabs(fxi - fx) > tol;
                        delta = fxi - fx;
                        f0new = fr0 + fxi * (fp0 * fxi + fq0);
                    } else {
                        delta = bcast(xi - xk, first);
                        f0new = bcast(r0 + xi * (p0 * xi + q0), first);
                        fxi = xi;                                   // meaningful in lane `first` only
                    }
                    if (moves) {
                        if (lane == first) { w.x[k] = fxi; mydelta = delta; }
                        moved |= 1u << first;
                        f0val = f0new;
                        uc = 0;
                        st.updates_p2++;
                        // the coordinates still to come see the move through their own entry of column `first`
                        if (act && lane > first) {
                            gl = fma(w.blk[first * 32 + lane], delta, gl);     // D[lane][first] = D[first][lane]
                            q0 = 2 * (gl - p0 * xk) + oq;
                            r0 = f0val - xk * (p0 * xk + q0);
                            rc = mfin ? choose_point_det_t<true>(p0, q0, r0, ml0, mh0, ml1, mh1, mnC, &xi) : choose_point_det(p0, q0, r0, ml0, mh0, ml1, mh1, mnC, &xi);
                        }
                    } else {
                        uc++;
                        if (uc == n) { done = true; break; }
                    }
                    cur = first + 1;
                }
                if (done) break;
                if (moved) lpc_axpy_multi(P, w, k0, moved, mydelta, lane);
            }
            // sparse objective: the per-coordinate constants of the NEXT aligned window are loaded while this one is decided (they come
            // from L2; in the late sweeps of a long run nearly every window is quiet and the round trip would be all it costs)
            double pf_p = 0.0, pf_q = 0.0, pf_r = 0.0, pf_od = 0.0, pf_oq = 0.0;
            int pf_rel = 0, pf_inc = 0, pf_k0 = -1;
            for (int k0 = 0; !V.obj_dense && k0 < n && !done;) {
                const int B = (n - k0 < 32) ? (n - k0) : 32;
                const bool act = lane < B;
                const int k = k0 + lane;
                double xk = 0.0, p0 = 0.0, q0 = 0.0, r0 = f0val, xi = 0.0;
                int rc = 0;
                double c_p = 0.0, c_q = 0.0, c_r = 0.0, c_od = 0.0, c_oq = 0.0;
                int c_rel = 0, c_inc = 0;
                if (pf_k0 == k0) { c_p = pf_p; c_q = pf_q; c_r = pf_r; c_rel = pf_rel; c_od = pf_od; c_oq = pf_oq; c_inc = pf_inc; }
                else if (act) { c_p = V.c_p[k]; c_q = V.c_q[k]; c_r = V.c_r[k]; c_rel = V.c_rel[k]; c_od = V.o_diag[k]; c_oq = V.o_q[k]; c_inc = V.o_inc[k]; }
                {
                    pf_k0 = (k0 + 32 < n) ? k0 + 32 : 0;
                    const int kp = pf_k0 + lane;
                    if (kp < n) { pf_p = V.c_p[kp]; pf_q = V.c_q[kp]; pf_r = V.c_r[kp]; pf_rel = V.c_rel[kp]; pf_od = V.o_diag[kp]; pf_oq = V.o_q[kp]; pf_inc = V.o_inc[kp]; }
                }
                bool certq = false;      // this step certainly changes nothing
                const bool use_filter = !(prm.mode & 0x100);
                if (act) {
                    const double p = c_p, q = c_q, r = c_r;
                    const int rel = c_rel;
                    xk = w.x[k];
                    if (!(mrel == rel && mp == p && mq == q && mr == r)) {
                        mnC = single_constraint_pieces(p, q, r, rel, viol_p2, &ml0, &mh0, &ml1, &mh1);
                        mfin = mnC > 0 && mnC <= 2 && !is_inf(ml0) && !is_inf(mh0) && (mnC < 2 || (!is_inf(ml1) && !is_inf(mh1)));
                        mp = p; mq = q; mr = r; mrel = rel;
                    }
                    if (c_inc) {
                        // obj = f0.get_onevar_func(x, k): t2 = P0[k,k], t1 = 2 (P0 z)_k + q0[k], t0 = f0(x) - x_k (t2 x_k + t1)
                        p0 = c_od;
                        q0 = 2 * (w.g[k] - p0 * xk) + c_oq;
                        r0 = f0val - xk * (p0 * xk + q0);
                        if (use_filter && mfin && p0 == 0.0) {
                            // linear restriction (MAXCUT: P_0 has a zero diagonal): the reference compares e q0 + r0 over the endpoints
                            // e of the pieces, so the minimiser is the smallest endpoint for q0 > 0 and the largest for q0 < 0 -- certain
                            // as soon as |q0| times the smallest gap between endpoints stands clear (relative 1e-9; rounding is 1e-16) of
                            // the magnitudes being added.  Exact zeros and near-ties are NOT certain and take the reference arithmetic.
                            const bool two = (mnC == 2);
                            const double aq = fabs(q0);
                            const double emax = two ? mh1 : mh0;
                            const double gap = two ? fmin(fmin(mh0 - ml0, ml1 - mh0), mh1 - ml1) : (mh0 - ml0);
                            const double est = (q0 > 0.0) ? ml0 : emax;
                            certq = gap > 0.0 && aq * gap > 1e-9 * (fabs(f0val) + aq * fmax(fabs(ml0), fabs(emax))) && fabs(est - xk) <= tol;
                        }
                    }
                }
                if (!__ballot_sync(FULL, act && !certq)) {
                    // a quiet window: update_counter += 1 per step (qcqp.py:172-176)
                    if (n - uc <= B) { st.steps_p2 += (n - uc); done = true; break; }
                    uc += B;
                    st.steps_p2 += B;
                    k0 += B;
                    continue;
                }
                if (act) {
                    rc = mfin ? choose_point_det_t<true>(p0, q0, r0, ml0, mh0, ml1, mh1, mnC, &xi) : choose_point_det(p0, q0, r0, ml0, mh0, ml1, mh1, mnC, &xi);
                }
                const bool wants_move = act && rc == 1 && fabs(xi - xk) > tol;
                const unsigned stop = __ballot_sync(FULL, wants_move || (act && rc == 2));
                const int first = stop ? (__ffs(stop) - 1) : B;
                // the `first` steps before it change nothing: update_counter += 1 each (qcqp.py:172-176)
                if (n - uc <= first) { st.steps_p2 += (n - uc); done = true; break; }   // converged inside the run
                uc += first;
                st.steps_p2 += first;
                if (first == B) { k0 += B; continue; }
                // ---- the coordinate that moves or needs a random number ----
                const int kf = k0 + first;
                const double fx = bcast(xk, first), fp0 = bcast(p0, first), fq0 = bcast(q0, first), fr0 = bcast(r0, first);
                double fxi = bcast(xi, first);
                const int frc = bcast_i(rc, first);
                st.steps_p2++;
                bool found = true;
                if (frc == 2) {
                    double cl[2], ch[2];
                    cl[0] = bcast(ml0, first); ch[0] = bcast(mh0, first); cl[1] = bcast(ml1, first); ch[1] = bcast(mh1, first);
                    const int nC = bcast_i(mnC, first);
                    int err = 0, fnd = 0;
                    double xv = 0.0;
                    if (lane == 0) {
                        MtRng rng;
                        rng.key = w.mt; rng.pos = pos;
                        fnd = choose_point(fp0, fq0, fr0, cl, ch, nC, rng, &xv, &err);
                        pos = rng.pos;
                    }
                    pos = bcast_i(pos, 0); err = bcast_i(err, 0); fnd = bcast_i(fnd, 0); fxi = bcast(xv, 0);
                    if (err) { st.status = err; dead = true; done = true; break; }
                    found = fnd != 0;
                }
                if (found && fabs(fxi - fx) > tol) {
                    if (lane == 0) w.x[kf] = fxi;
                    f0val = fr0 + fxi * (fp0 * fxi + fq0);        // f_0(x) = t0 + b (t2 b + t1)
                    lpc_axpy(P, V, w, kf, fxi - fx, lane);
                    uc = 0;
                    st.updates_p2++;
                } else {
                    uc++;
                    if (uc == n) { done = true; break; }
                }
                k0 = kf + 1;
            }
            if (!done && prm.refresh_every > 0 && ((t + 1) % prm.refresh_every) == 0) f0val = lpc_refresh(P, V, w, lane);
        }
    }

    // ---------------- results (qcqp.py:415-417) ----------------
    for (int i = lane; i < n; i += 32) X[rr * n + i] = w.x[i];
    for (int i = lane; i < 624; i += 32) rngs[rr].key[i] = w.mt[i];
    if (stage == 2) {
        if (lane == 0) { rngs[rr].pos = pos; stats_out[rr] = st; }
        return;                                                  // (f0, maxviol): batched eval kernels, launched next
    }
    const double f0fin = lpc_refresh(P, V, w, lane);
    mv = lpc_max_violation(P, V, w, lane);
    if (lane == 0) {
        rngs[rr].pos = pos;
        f0_out[rr] = f0fin;
        mv_out[rr] = mv;
        if (stats_out) stats_out[rr] = st;
    }
}

int eval_launch(qcqp_pack* p, const double* dX, int R, double* df0, double* dmv, double* dviol, cudaStream_t stream);
int gemm_plain_launch(int M, int N, int K, const double* dA, int lda, const double* dB, int ldb, double* dC, int ldc, cudaStream_t stream);

int lpc_launch(qcqp_pack* p, const CdK& k_in, const double* dX0, int R, qcqp_rng_state* drng, double* dX, double* df0, double* dmv,
               qcqp_cd_stats* dstats, cudaStream_t stream)
{
    CdK k = k_in;
    {
        const char* eq = getenv("QCQP_LPC_QUIET");     // QCQP_LPC_QUIET=0: no certified quiet-window filter in the sparse loop (A/B; same results)
        if (eq && eq[0] == '0') k.mode |= 0x100;
    }
    const int n = p->v.n;
    const int npad = (n + 1) & ~1;
    size_t smem = (size_t)2 * npad * 8 + 624 * 4 + (p->lpc.obj_dense ? 32 * 32 * 8 : 0);
    if (smem > (size_t)max_smem_optin(p->device)) return fail(QCQP_ERR_CAPACITY, "qcqp_cd_improve: n too large for the separable kernel");
    QCQP_CUDA_TRY(cudaFuncSetAttribute(cd_lpc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        // one warp per CTA: only ceil(R / #SM) CTAs are ever resident on an SM, so carve out just their shared memory and
        // leave the rest of the 228 KB to L1 -- the per-coordinate constants and the rows of P_0 are re-read through it
        const int per_sm = (R + num_sms(p->device) - 1) / num_sms(p->device);
        const double need = (double)per_sm * (double)(smem + 1024);
        int pct = (int)(100.0 * need / (228.0 * 1024.0)) + 1;
        if (pct < 10) pct = 10;
        if (pct > 100) pct = 100;
        QCQP_CUDA_TRY(cudaFuncSetAttribute(cd_lpc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    }
    if (p->lpc.obj_dense && R >= 64) {
        // phase 1 -> G = X P_0 (tiled GEMM) -> phase 2 -> batched (f0, maxviol)
        const size_t gbytes = ((size_t)R * npad * 8 + 255) & ~(size_t)255;
        const size_t sbytes = ((size_t)R * sizeof(qcqp_cd_stats) + 255) & ~(size_t)255;
        const char* ep = getenv("QCQP_LPC_PRE");                      // QCQP_LPC_PRE=0: bisect inside the sweep kernel (A/B; same results)
        const bool use_pre = k.phase1 && !(ep && ep[0] == '0');
        int rc = ensure_workspace(p, gbytes + sbytes + (use_pre ? (size_t)R * n * sizeof(LpcPre) : 0));
        if (rc != QCQP_OK) return rc;
        double* G = (double*)p->ws;
        qcqp_cd_stats* stats = dstats ? dstats : (qcqp_cd_stats*)((char*)p->ws + gbytes);
        LpcPre* pre = use_pre ? (LpcPre*)((char*)p->ws + gbytes + sbytes) : nullptr;
        if (!p->ev_ok) {
            for (int i = 0; i < 6; i++) QCQP_CUDA_TRY(cudaEventCreate(&p->ev[i]));
            p->ev_ok = true;
        }
        p->ev_count = 0;
        QCQP_CUDA_TRY(cudaEventRecord(p->ev[0], stream));
        if (pre) {
            const size_t tot = (size_t)R * n;
            lpc_p1_pre_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(p->v, p->lpc, k, dX0, R, pre);
            QCQP_CUDA_TRY(cudaGetLastError());
        }
        cd_lpc_kernel<<<R, 32, smem, stream>>>(p->v, p->lpc, k, 1, dX0, R, drng, dX, nullptr, df0, dmv, stats, pre);
        QCQP_CUDA_TRY(cudaGetLastError());
        QCQP_CUDA_TRY(cudaEventRecord(p->ev[1], stream));
        rc = gemm_plain_launch(R, n, n, dX, n, p->v.dense_P, p->v.ld, G, npad, stream);
        if (rc != QCQP_OK) return rc;
        QCQP_CUDA_TRY(cudaEventRecord(p->ev[2], stream));
        // phase 2: the resolver / helper kernel with TMA-staged diagonal blocks (cd_lpc2.cu); QCQP_LPC2=0 keeps the one-warp kernel
        // of round 1 for A/B runs (bit-identical results)
        const char* e2 = getenv("QCQP_LPC2");
        if (lpc2_supported(n) && !(e2 && e2[0] == '0')) {
            rc = lpc2_launch(p, k, R, drng, dX, G, stats, stream);
            if (rc != QCQP_OK) return rc;
        } else {
            cd_lpc_kernel<<<R, 32, smem, stream>>>(p->v, p->lpc, k, 2, dX0, R, drng, dX, G, df0, dmv, stats, nullptr);
            QCQP_CUDA_TRY(cudaGetLastError());
        }
        QCQP_CUDA_TRY(cudaEventRecord(p->ev[3], stream));
        rc = eval_launch(p, dX, R, df0, dmv, nullptr, stream);
        if (rc != QCQP_OK) return rc;
        QCQP_CUDA_TRY(cudaEventRecord(p->ev[4], stream));
        p->ev_count = 5;
        return QCQP_OK;
    }
    cd_lpc_kernel<<<R, 32, smem, stream>>>(p->v, p->lpc, k, 0, dX0, R, drng, dX, nullptr, df0, dmv, dstats, nullptr);
    QCQP_CUDA_TRY(cudaGetLastError());
    return QCQP_OK;
}

}  // namespace qcqp
