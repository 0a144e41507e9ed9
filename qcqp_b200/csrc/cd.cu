// cd.cu -- two-phase coordinate descent for R independent restarts (improve_coord_descent, qcqp.py:181-192;
// coord_descent_phase1 :101-148; coord_descent_phase2 :152-178; get_onevar_func utilities.py:99-105;
// onevar_qcqp utilities.py:241-288).
//
// Mapping (DESIGN.md "CD sweep kernel"): ONE WARP PER RESTART.  A CTA holds W restart warps that walk the coordinates
// k = 0..n-1 together, plus (when the problem has dense forms) one producer warp that streams row k of every dense
// P_j from HBM/L2 into a shared-memory ring with cp.async.bulk (TMA unit) + mbarrier full/empty pairs, so a row is
// fetched once per CTA and consumed by all W restarts.  Per restart, in shared memory: x[n], the cached f_j(x),
// the MT19937 state, the one-variable coefficients of the forms incident to x_k, and the sweep-line events.
// Per coordinate step a warp does: row dots (lanes over columns, shuffle reduction) -> (t2, t1, t0) per incident form
// -> feasible set at level s (lanes over forms, fold + events) -> lane 0: sweep line + minimiser + RNG -> move and
// incremental update f_j += ... of the incident forms.
#include <cstdio>
#include <cstdlib>

#include "cd_holes.cuh"
#include "cd_shared.cuh"
#include "common.cuh"
#include "forms_eval.cuh"
#include "onevar.cuh"

namespace qcqp {

constexpr int CD_LONG_ROW = 48;      // sparse rows longer than this are dotted by the whole warp (fast mode)

struct CdLayout {
    int W;            // restart warps per CTA
    int S;            // ring stages (0: no dense forms)
    int sc_cap;       // coefficient scratch entries in smem (0: global scratch)
    int fval_smem;    // cached f_j in smem?
    int grad;         // cached dense row dots g_d = P_d x in smem (gradient mode)
    int hcap;         // hole capacity (power of two, >= 32)
    // byte offsets inside the dynamic smem block
    unsigned off_bar, off_ring, off_warp0, warp_stride;
    unsigned o_x, o_fval, o_mt, o_scp, o_scq, o_scr, o_screl, o_hx, o_clo, o_chi, o_dd, o_misc, o_g;
    unsigned total;
};




enum { PH_P1 = 0, PH_P2 = 1, PH_DONE = 2 };

// ---------------------------------------------------------------------------------------------------------
// row dots
// ---------------------------------------------------------------------------------------------------------
// dense row (length n, staged in the smem ring) against z: the caller has stored 0 at x[k], so z == x here.
// Both arrays are 16-byte aligned in shared memory; lanes own consecutive pairs (LDS.128).
__device__ __forceinline__ double dense_row_dot_warp(const double* row, const double* x, int n, int lane)
{
    __builtin_assume(__isShared(row));
    __builtin_assume(__isShared(x));
    const double2* r2 = reinterpret_cast<const double2*>(row);
    const double2* x2 = reinterpret_cast<const double2*>(x);
    const int n2 = n >> 1;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int i = lane;
    for (; i + 32 < n2; i += 64) {
        const double2 a = r2[i], xa = x2[i], c = r2[i + 32], xc = x2[i + 32];
        s0 = fma(a.x, xa.x, s0); s1 = fma(a.y, xa.y, s1);
        s2 = fma(c.x, xc.x, s2); s3 = fma(c.y, xc.y, s3);
    }
    if (i < n2) {
        const double2 a = r2[i], xa = x2[i];
        s0 = fma(a.x, xa.x, s0); s1 = fma(a.y, xa.y, s1);
    }
    if ((n & 1) && lane == 0) s2 = fma(row[n - 1], x[n - 1], s2);
    return warp_sum((s0 + s1) + (s2 + s3));
}
__device__ __forceinline__ double dense_row_dot_seq(const double* row, const double* x, int n, int k)
{
    double s = 0.0;
    for (int c = 0; c < n; c++) {
        double v = row[c];
        if (v != 0.0 && c != k) s = s + v * x[c];
    }
    return s;
}

// ---------------------------------------------------------------------------------------------------------
// onevar_qcqp, register path: at most one constraint per lane (coordinates with <= 32 incident forms).
// The feasible set of a lane's constraint is memoised on (p, q, r, s, relop): in phase 2 of a Boolean problem every
// coordinate presents the same x_k^2 = 1 at the same frozen level, so the sqrt/div chain runs once per restart.
// ---------------------------------------------------------------------------------------------------------
struct IvMemo {
    double p, q, r, s;
    int rel, c;
    Ival I0, I1;
};
// warp-uniform: the pieces in w.clo/w.chi are those of lane `src`'s memoised constraint standing alone
struct PieceCache {
    bool valid;
    int src, nC;
};

__device__ __forceinline__ int solve_small(const WarpMem& w, bool active, double p, double q, double r, int rel, double s, double p0,
                                           double q0, double r0, IvMemo& memo, PieceCache& pc, MtRng& rng, int lane, double* xout, int* err)
{
    *err = 0;
    *xout = 0.0;
    Fold f;
    f.init();
    int c = 0;
    Ival I[2];
    I[0].lo = I[0].hi = I[1].lo = I[1].hi = 0.0;
    const bool counted = active && !(p == 0.0 && q == 0.0);   // nfs filter of qcqp.py:116,166
    bool hit = false;
    if (counted) {
        if (memo.c >= 0 && memo.p == p && memo.q == q && memo.r == r && memo.s == s && memo.rel == rel) {
            c = memo.c; I[0] = memo.I0; I[1] = memo.I1;
            hit = true;
        } else {
            c = feasible_intervals(p, q, r, rel, s, I);
            memo.p = p; memo.q = q; memo.r = r; memo.s = s; memo.rel = rel; memo.c = c; memo.I0 = I[0]; memo.I1 = I[1];
        }
        f.mcnt = 1;
        if (c == 0) f.nempty = 1;
        else if (c == 1) f.add_single(I[0].lo, I[0].hi);
    }
    const unsigned act = __ballot_sync(FULL, counted);
    const unsigned two = __ballot_sync(FULL, counted && c == 2);
    const int nact = __popc(act);
    if (__popc(two) >= 2) {
        // several two-interval constraints: hull into the fold, hole (if any) stays in this lane's registers
        pc.valid = false;
        double ha = QCQP_INF, hb = QCQP_INF;
        if (counted && c == 2) {
            f.add_single(I[0].lo, I[1].hi);
            if (I[0].hi < I[1].lo) { ha = I[0].hi; hb = I[1].lo; }
        }
        fold_allreduce(f);
        return holes_finish(w, f, 32, true, ha, hb, p0, q0, r0, rng, lane, xout, err);
    }
    const int only = __ffs(act) - 1;
    // a single counted constraint whose feasible set is memoised: the pieces of the previous call are still in place
    const bool reuse = (nact == 1) && pc.valid && pc.src == only && (__ballot_sync(FULL, hit) & act) != 0;
    if (!reuse) {
        if (nact > 1) fold_allreduce(f);
        else if (nact == 1) fold_bcast(f, only);
        if (f.nempty > 0) { pc.valid = false; return 0; }
    }
    const int ntwo = __popc(two);
    if (reuse || ntwo <= 1) {
        if (!reuse) {
            Ival T0 = I[0], T1 = I[1];
            if (ntwo == 1) {
                const int src = __ffs(two) - 1;
                T0.lo = bcast(I[0].lo, src); T0.hi = bcast(I[0].hi, src);
                T1.lo = bcast(I[1].lo, src); T1.hi = bcast(I[1].hi, src);
            }
            int nC = 0;
            if (lane == 0) nC = sweep_small8(f, ntwo == 1, T0, T1, w.clo, w.chi);
            pc.nC = bcast_i(nC, 0);
            pc.valid = (nact == 1);
            pc.src = only;
        }
        int found = 0, e = 0;
        double xv = 0.0;
        if (lane == 0) found = choose_point(p0, q0, r0, w.clo, w.chi, pc.nC, rng, &xv, &e);
        *xout = bcast(xv, 0);
        *err = bcast_i(e, 0);
        return bcast_i(found, 0);
    }
    return 0;   // not reached: ntwo >= 2 is handled above
}

// ---------------------------------------------------------------------------------------------------------
// Phase-1 bisection for a coordinate with ONE counted constraint (every Boolean-type problem): the reference probes
// s = (ss+es)/2 one level at a time (qcqp.py:122-131).  Here 31 lanes evaluate the 31 nodes of the next five levels of
// that decision tree at once -- each lane replays the same (ss+es)/2 arithmetic along its own path, so the s values are
// the reference's bit for bit -- and then the warp walks the actual path, lane 0 drawing the random numbers of the
// feasible probes in the reference's order.  Returns false when the reference would have raised (err set).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool spec_bisect(double p, double q, double r, int rel, double tol, double& ss, double& es, double& new_xi,
                                            double& new_viol, MtRng& rng, int lane, int* err)
{
    *err = 0;
    while (es - ss > tol) {
        // ---- my node: lane L <-> node v = L + 1 (root = 1; child 2v = infeasible branch, 2v+1 = feasible branch) ----
        const int v = lane + 1;
        const int depth = 31 - __clz(v);
        double lss = ss, les = es;
        bool nodeok = (lane < 31);
        for (int i = depth - 1; i >= 0 && nodeok; i--) {
            if (!(les - lss > tol)) { nodeok = false; break; }
            const double sm = (lss + les) / 2;
            if ((v >> i) & 1) les = sm; else lss = sm;
        }
        nodeok = nodeok && (les - lss > tol);
        const double sv = (lss + les) / 2;
        int nC = 0;
        double cl[4], ch[4];
        cl[0] = cl[1] = ch[0] = ch[1] = 0.0;
        if (nodeok) {
            Fold f;
            f.init();
            f.mcnt = 1;
            Ival I[2];
            const int c = feasible_intervals(p, q, r, rel, sv, I);
            if (c == 1) f.add_single(I[0].lo, I[0].hi);
            if (c > 0) nC = sweep_small8(f, c == 2, I[0], I[1], cl, ch);
        }
        const unsigned okmask = __ballot_sync(FULL, nodeok);
        const unsigned feas = __ballot_sync(FULL, nodeok && nC > 0);
        // ---- walk the path actually taken ----
        int node = 1;
#pragma unroll 1
        for (int lvl = 0; lvl < 5; lvl++) {
            const int src = node - 1;
            if (!((okmask >> src) & 1u)) break;          // es - ss <= tol here: the while loop of the reference ends
            const double s = __shfl_sync(FULL, sv, src);
            if ((feas >> src) & 1u) {
                const int k = __shfl_sync(FULL, nC, src);
                const double a0 = __shfl_sync(FULL, cl[0], src), b0 = __shfl_sync(FULL, ch[0], src);
                const double a1 = __shfl_sync(FULL, cl[1], src), b1 = __shfl_sync(FULL, ch[1], src);
                double xv = 0.0;
                int e = 0;
                if (lane == 0) {
                    // onevar_qcqp with the zero objective: np.random.uniform(*C[np.random.choice(len(C))])  (utilities.py:266-267)
                    const int idx = rng.choice(k);
                    const double lo = idx ? a1 : a0, hi = idx ? b1 : b0;
                    if (is_inf(lo) || is_inf(hi)) e = QCQP_RUN_UNBOUNDED_UNIFORM;
                    else xv = rng.uniform(lo, hi);
                }
                e = bcast_i(e, 0);
                if (e) { *err = e; return false; }
                new_xi = bcast(xv, 0);
                new_viol = s;
                es = s;
                node = 2 * node + 1;
            } else {
                ss = s;
                node = 2 * node;
            }
        }
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------------
// gradient mode: g_d = P_d x for dense slot d from scratch (lanes own column pairs, rows stream coalesced from L2), and
// f_d(x) = x.g_d + q_d.x + r_d from it.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void refresh_dense_slot(const PackView& P, const WarpMem& w, int d, int npad, int lane)
{
    const int n = P.n, ld = P.ld, j = P.dense_form[d];
    const double* M = P.dense_P + (size_t)d * n * ld;
    double* g = w.g + (size_t)d * npad;
    const int n2 = (n + 1) >> 1;                  // ld is even and the pad column is zero, so pairs never run off a row
    for (int c0 = 0; c0 < n2; c0 += 128) {
        double2 a0 = make_double2(0.0, 0.0), a1 = a0, a2 = a0, a3 = a0;
        const int c = c0 + lane;
        const bool v0 = c < n2, v1 = c + 32 < n2, v2 = c + 64 < n2, v3 = c + 96 < n2;
        for (int r = 0; r < n; r++) {
            const double xr = w.x[r];
            const double2* row = reinterpret_cast<const double2*>(M + (size_t)r * ld);
            if (v0) { const double2 t = row[c]; a0.x = fma(t.x, xr, a0.x); a0.y = fma(t.y, xr, a0.y); }
            if (v1) { const double2 t = row[c + 32]; a1.x = fma(t.x, xr, a1.x); a1.y = fma(t.y, xr, a1.y); }
            if (v2) { const double2 t = row[c + 64]; a2.x = fma(t.x, xr, a2.x); a2.y = fma(t.y, xr, a2.y); }
            if (v3) { const double2 t = row[c + 96]; a3.x = fma(t.x, xr, a3.x); a3.y = fma(t.y, xr, a3.y); }
        }
        double2* g2 = reinterpret_cast<double2*>(g);
        if (v0) g2[c] = a0;
        if (v1) g2[c + 32] = a1;
        if (v2) g2[c + 64] = a2;
        if (v3) g2[c + 96] = a3;
    }
    __syncwarp();
    double acc = 0.0;
    for (int c = lane; c < n; c += 32) acc = fma(w.x[c], g[c], acc);
    for (long long e = P.q_ptr[j] + lane; e < P.q_ptr[j + 1]; e += 32) acc = fma(P.q_val[e], w.x[P.q_idx[e]], acc);
    acc = warp_sum(acc) + P.r[j];
    if (lane == 0) w.fval[j] = acc;
}

// ---------------------------------------------------------------------------------------------------------
// cached f_j(x) (and, in gradient mode, g_d) from scratch for forms j >= j0; returns max constraint violation
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double refresh_fvals(const PackView& P, const WarpMem& w, int j0, int mode, int npad, int lane)
{
    double* fv = w.fval;
    if (mode == MODE_GRAD) {
        eval_forms(P, w.x, j0, P.m, false, lane, [&](int j, double v) { fv[j] = v; }, /*skip_dense=*/true);
        for (int d = 0; d < P.n_dense; d++)
            if (P.dense_form[d] >= j0) refresh_dense_slot(P, w, d, npad, lane);
    } else {
        eval_forms(P, w.x, j0, P.m, mode == MODE_STRICT, lane, [&](int j, double v) { fv[j] = v; }, false);
    }
    __syncwarp();
    double mv = -QCQP_INF;
    for (int j = 1 + lane; j <= P.m; j += 32) {
        double v = violation_of(P.relop[j], fv[j]);
        mv = (v > mv) ? v : mv;
    }
    return warp_max(mv);
}

// per-lane metadata of one incidence, prefetched one coordinate ahead
struct Meta {
    uint32_t fw;
    int rbeg, rlen;
    double t2, qk;
};
__device__ __forceinline__ Meta load_meta(const PackView& P, int e, bool valid)
{
    Meta mt;
    mt.fw = 0xffffffffu; mt.rbeg = 0; mt.rlen = 0; mt.t2 = 0.0; mt.qk = 0.0;
    if (valid) {
        mt.fw = P.inc_form[e]; mt.rbeg = P.inc_rbeg[e]; mt.rlen = P.inc_rlen[e];
        mt.t2 = P.inc_t2[e]; mt.qk = P.inc_qk[e];
    }
    return mt;
}

__device__ __forceinline__ double sparse_dot_seq(const PackView& P, int rbeg, int rlen, const double* x)
{
    double s = 0.0;
    for (int t = rbeg; t < rbeg + rlen; t++) s = s + P.row_val[t] * x[P.row_col[t]];
    return s;
}
__device__ __forceinline__ double sparse_dot_warp(const PackView& P, int rbeg, int rlen, const double* x, int lane)
{
    double s = 0.0;
    for (int t = rbeg + lane; t < rbeg + rlen; t += 32) s = fma(P.row_val[t], x[P.row_col[t]], s);
    return warp_sum(s);
}

// gradient mode, after x_k += delta: g_d += delta * (row k of P_d) for the maintained dense slots (P_d symmetric)
__device__ __forceinline__ void grad_axpy(const WarpMem& w, const double* rows, int d0, int nd, int n, int ld, int npad, double delta,
                                          int lane)
{
    __builtin_assume(__isShared(rows));
    const int n2 = (n + 1) >> 1;
    for (int d = d0; d < nd; d++) {
        const double2* r2 = reinterpret_cast<const double2*>(rows + (size_t)d * ld);
        double2* g2 = reinterpret_cast<double2*>(w.g + (size_t)d * npad);
        for (int c = lane; c < n2; c += 32) {
            const double2 rv = r2[c];
            double2 gv = g2[c];
            gv.x = fma(rv.x, delta, gv.x);
            gv.y = fma(rv.y, delta, gv.y);
            g2[c] = gv;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) cd_kernel(PackView P, CdK prm, CdLayout lay, const double* __restrict__ X0, int R, qcqp_rng_state* rngs,
                          double* __restrict__ X, double* __restrict__ f0_out, double* __restrict__ mv_out, qcqp_cd_stats* stats_out,
                          double* ws_fval, double* ws_scr, int* ws_screl)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = P.n, m = P.m, nd = P.n_dense, ld = P.ld;
    const int W = lay.W, S = lay.S;
    const bool ring = (nd > 0);
    const int mode = prm.mode;
    const int npad = (n + 1) & ~1;
    const bool obj_dense = ring && P.dense_form[0] == 0;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + lay.off_bar);
    uint64_t* bar_empty = bar_full + (S > 0 ? S : 1);
    double* ringbuf = reinterpret_cast<double*>(smem + lay.off_ring);
    const unsigned row_bytes = (unsigned)ld * 8u;

    if (ring && threadIdx.x == 0) {
        for (int s = 0; s < S; s++) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], W); }
        mbar_fence_init();
    }
    __syncthreads();

    // ================================= producer warp: stream dense rows =========================================
    if (ring && warp == W) {
        unsigned slot = 0;
        for (;;) {
            if (!__syncthreads_or(0)) break;   // pairs with the restart warps' votes at every sweep boundary
            // phase 1 never reads the objective: its rows are streamed only while some restart of this CTA is in phase 2
            const int d0 = (obj_dense && !__syncthreads_or(0)) ? 1 : 0;
            for (int k = 0; k < n; k++, slot++) {
                int st = slot % S;
                if (slot >= (unsigned)S) mbar_wait(&bar_empty[st], ((slot / S) + 1) & 1);
                if (lane == 0) {
                    if (nd - d0 > 0) {
                        mbar_arrive_expect_tx(&bar_full[st], row_bytes * (nd - d0));
                        for (int d = d0; d < nd; d++)
                            bulk_g2s(ringbuf + ((size_t)st * nd + d) * ld, P.dense_P + ((size_t)d * n + k) * ld, row_bytes, &bar_full[st]);
                    } else {
                        mbar_arrive(&bar_full[st]);
                    }
                }
                __syncwarp();
            }
        }
        return;
    }

    // ================================= restart warps ==============================================================
    const int restart = blockIdx.x * W + warp;
    const bool live = restart < R;
    unsigned char* wb = smem + lay.off_warp0 + (size_t)warp * lay.warp_stride;
    WarpMem w;
    w.x = reinterpret_cast<double*>(wb + lay.o_x);
    w.mt = reinterpret_cast<uint32_t*>(wb + lay.o_mt);
    w.hx = reinterpret_cast<double2*>(wb + lay.o_hx);
    w.clo = reinterpret_cast<double*>(wb + lay.o_clo);
    w.chi = reinterpret_cast<double*>(wb + lay.o_chi);
    w.dd = reinterpret_cast<double*>(wb + lay.o_dd);
    w.misc = reinterpret_cast<int*>(wb + lay.o_misc);
    w.g = reinterpret_cast<double*>(wb + lay.o_g);
    const size_t rr = live ? (size_t)restart : 0;
    w.fval = lay.fval_smem ? reinterpret_cast<double*>(wb + lay.o_fval) : ws_fval + rr * (size_t)(m + 1);
    // coefficient scratch: shared memory for coordinates with <= sc_cap incident forms, a per-restart HBM block beyond
    Scratch sc_s, sc_g;
    sc_s.p = reinterpret_cast<double*>(wb + lay.o_scp);
    sc_s.q = reinterpret_cast<double*>(wb + lay.o_scq);
    sc_s.r = reinterpret_cast<double*>(wb + lay.o_scr);
    sc_s.rel = reinterpret_cast<int*>(wb + lay.o_screl);
    sc_g = sc_s;
    if (P.max_inc > lay.sc_cap) {
        sc_g.p = ws_scr + rr * 3 * (size_t)P.max_inc;
        sc_g.q = sc_g.p + P.max_inc;
        sc_g.r = sc_g.q + P.max_inc;
        sc_g.rel = ws_screl + rr * (size_t)P.max_inc;
    }

    MtRng rng;
    rng.key = w.mt;
    rng.pos = 624;
    int phase = PH_DONE;
    qcqp_cd_stats st;
    st.steps_p1 = st.steps_p2 = st.updates_p1 = st.updates_p2 = st.steps_skipped = 0;
    st.sweeps_p1 = st.sweeps_p2 = 0; st.status = QCQP_RUN_OK; st.ran_phase2 = 0;

    if (live) {
        for (int i = lane; i < n; i += 32) w.x[i] = X0[rr * n + i];
        for (int i = lane; i < 624; i += 32) w.mt[i] = rngs[rr].key[i];
        rng.pos = rngs[rr].pos;
        phase = PH_P1;
    }
    __syncwarp();

    const bool strict = (mode == MODE_STRICT);
    const bool grad = (mode == MODE_GRAD) && ring;
    const double tol = prm.tol, viol_tol = prm.viol_tol;
    int t = 0;                    // sweeps done in the current phase
    long long update_counter = 0;
    double viol_last = QCQP_INF;  // phase 1
    double viol_p2 = 0.0;         // phase 2: frozen at entry (qcqp.py:157)
    bool p1_over = live && !prm.phase1;
    unsigned slot = 0;
    PieceCache pc;
    pc.valid = false; pc.src = 0; pc.nC = 0;
    IvMemo memo;
    memo.c = -1; memo.p = memo.q = memo.r = memo.s = 0.0; memo.rel = 0;
    memo.I0.lo = memo.I0.hi = memo.I1.lo = memo.I1.hi = 0.0;

    for (;;) {
        // ---------------- sweep boundary: phase transitions (warp-uniform) ----------------
        if (phase == PH_P1) {
            if (p1_over || t >= prm.num_iters || viol_last < viol_tol) {
                // improve_coord_descent: if max(prob.violations(x)) < viol_tol: phase 2   (qcqp.py:189-190)
                double mv = refresh_fvals(P, w, 0, mode, npad, lane);
                if (m == 0) { st.status = QCQP_RUN_EMPTY_MAX; phase = PH_DONE; }
                else if (mv < viol_tol) { phase = PH_P2; viol_p2 = mv; t = 0; update_counter = 0; st.ran_phase2 = 1; }
                else phase = PH_DONE;
            } else if (t == 0) {
                refresh_fvals(P, w, 1, mode, npad, lane);
            }
        }
        if (phase == PH_P2 && t >= prm.num_iters) phase = PH_DONE;
        if (ring) {
            if (!__syncthreads_or(phase != PH_DONE)) break;
            if (obj_dense) __syncthreads_or(phase == PH_P2);   // tells the producer whether objective rows are needed
        } else if (phase == PH_DONE) break;
        if (phase == PH_P1) st.sweeps_p1++;
        if (phase == PH_P2) st.sweeps_p2++;
        bool skip = false;   // phase 1 'failed' break: the rest of this sweep is not executed (qcqp.py:138-141)
        const long long upd_before = st.updates_p1;

        // incidence ranges and per-lane metadata run one coordinate ahead of the work
        int pb0 = P.inc_ptr[0], pb1 = P.inc_ptr[1], pb2 = P.inc_ptr[n >= 2 ? 2 : 1];
        Meta pf = load_meta(P, pb0 + lane, pb0 + lane < pb1);
        for (int k = 0; k < n; k++, slot++) {
            const Meta cur = pf;
            const int beg = pb0, end = pb1;
            pb0 = pb1; pb1 = pb2;
            if (k + 3 <= n) pb2 = P.inc_ptr[k + 3];
            pf = load_meta(P, pb0 + lane, (k + 1 < n) && (pb0 + lane < pb1));

            const bool work = (phase != PH_DONE) && !skip;
            // z = x with x_k := 0 (utilities.py:100-101): park x_k in a register and zero it in place for this step
            double xk = 0.0;
            if (work) {
                xk = w.x[k];
                if (!grad) {
                    __syncwarp();
                    if (lane == 0) w.x[k] = 0.0;
                    __syncwarp();
                }
            }
            // ---- dense rows of this coordinate: (P_d z)_k for every dense slot d ----
            const int sg = ring ? (int)(slot % S) : 0;
            const double* rows = ringbuf + (size_t)sg * nd * ld;
            const int dfirst = (phase == PH_P1 && obj_dense) ? 1 : 0;     // phase 1 never touches the objective
            if (ring) {
                mbar_wait(&bar_full[sg], (slot / S) & 1);
                if (work) {
                    if (grad) {
                        // cached g_d = P_d x: (P_d z)_k = g_d[k] - P_d[k,k] x_k; the row itself is only needed if x_k moves
                        for (int d = dfirst + lane; d < nd; d += 32)
                            w.dd[d] = w.g[(size_t)d * npad + k] - rows[(size_t)d * ld + k] * xk;
                    } else if (!strict) {
                        for (int d = dfirst; d < nd; d++) {
                            double v = dense_row_dot_warp(rows + (size_t)d * ld, w.x, n, lane);
                            if (lane == 0) w.dd[d] = v;
                        }
                    } else {
                        for (int d = lane; d < nd; d += 32) w.dd[d] = dense_row_dot_seq(rows + (size_t)d * ld, w.x, n, k);
                    }
                }
                __syncwarp();
                if (!grad || !work) { if (lane == 0) mbar_arrive(&bar_empty[sg]); }
            }
            if (!work) {
                if (!ring) break;
                continue;
            }
            const int cnt = end - beg;
            double new_xi = xk;
            bool move = false;
            bool dead = false;   // the reference would have raised: stop this restart, leave x as it was

            if (cnt <= 32) {
                // ======================= register path: one incidence per lane =======================
                const bool valid = lane < cnt;
                const int form = (int)(cur.fw & INC_FORM_MASK);
                const bool is_obj = valid && form == 0;
                const bool has_obj = __ballot_sync(FULL, is_obj) != 0;   // the objective, when incident, is lane 0
                const bool need = valid && !(phase == PH_P1 && is_obj);
                double dot = 0.0;
                bool deferred = false;
                if (need) {
                    if (cur.rlen < 0) dot = w.dd[cur.rbeg];
                    else if (strict || cur.rlen <= CD_LONG_ROW) dot = sparse_dot_seq(P, cur.rbeg, cur.rlen, w.x);
                    else deferred = true;
                }
                unsigned coop = __ballot_sync(FULL, deferred);
                while (coop) {
                    const int src = __ffs(coop) - 1;
                    coop &= coop - 1;
                    double v = sparse_dot_warp(P, bcast_i(cur.rbeg, src), bcast_i(cur.rlen, src), w.x, lane);
                    if (lane == src) dot = v;
                }
                // (t2, t1, t0) of get_onevar_func (utilities.py:99-105); t0 from the cached f_j(x)
                const double cp = cur.t2;
                const double cq = 2 * dot + cur.qk;
                const double cr = need ? (w.fval[form] - xk * (cp * xk + cq)) : 0.0;
                const int crel = (int)((cur.fw >> INC_RELOP_SHIFT) & 3);
                const bool active = valid && !is_obj;
                double p0 = 0.0, q0 = 0.0, r0 = 0.0;
                if (phase == PH_P2) {
                    if (has_obj) { p0 = bcast(cp, 0); q0 = bcast(cq, 0); r0 = bcast(cr, 0); }
                    else r0 = w.fval[0];
                }
                if (phase == PH_P1) {
                    // ---- phase 1: bisect the violation level (qcqp.py:113-141) ----
                    st.steps_p1++;
                    const bool nz = active && !(cp == 0.0 && cq == 0.0);
                    if (__ballot_sync(FULL, nz) == 0) { st.status = QCQP_RUN_EMPTY_MAX; dead = true; }
                    else {
                        const double viol = warp_max(nz ? violation_of(crel, onevar_eval(cp, cq, cr, xk)) : -QCQP_INF);
                        double new_viol = viol;
                        double ss = -tol, es = viol - viol_tol;
                        const unsigned nzmask = __ballot_sync(FULL, nz);
                        if (__popc(nzmask) == 1) {
                            const int src = __ffs(nzmask) - 1;
                            int err;
                            if (!spec_bisect(bcast(cp, src), bcast(cq, src), bcast(cr, src), bcast_i(crel, src), tol, ss, es, new_xi,
                                             new_viol, rng, lane, &err)) { st.status = err; dead = true; }
                        }
                        while (!dead && es - ss > tol) {
                            const double s = (ss + es) / 2;
                            double xi;
                            int err;
                            int ok = solve_small(w, active, cp, cq, cr, crel, s, 0.0, 0.0, 0.0, memo, pc, rng, lane, &xi, &err);
                            if (err) { st.status = err; dead = true; break; }
                            if (!ok) ss = s;
                            else { new_xi = xi; new_viol = s; es = s; }
                        }
                        if (!dead) {
                            if (new_viol < viol) { move = true; update_counter = 0; st.updates_p1++; }
                            else {
                                update_counter++;
                                if (update_counter == n) skip = true;
                            }
                        }
                    }
                } else {
                    // ---- phase 2: minimise the objective at the frozen level (qcqp.py:162-176) ----
                    st.steps_p2++;
                    double xi;
                    int err;
                    int ok = solve_small(w, active, cp, cq, cr, crel, viol_p2, p0, q0, r0, memo, pc, rng, lane, &xi, &err);
                    if (err) { st.status = err; dead = true; }
                    else if (ok && fabs(xi - xk) > tol) { move = true; new_xi = xi; update_counter = 0; st.updates_p2++; }
                    else {
                        update_counter++;
                        if (update_counter == n) phase = PH_DONE;   // converged
                    }
                }
                // f_j(x) = t0 + b (t2 b + t1) for every incident form; x_k back in place (moved or not)
                if (move && need) w.fval[form] = cr + new_xi * (cp * new_xi + cq);
                if (lane == 0) w.x[k] = new_xi;
                if (grad) {
                    if (move) grad_axpy(w, rows, dfirst, nd, n, ld, npad, new_xi - xk, lane);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_empty[sg]);
                }
                if (dead) phase = PH_DONE;
                __syncwarp();
                continue;
            }
            pc.valid = false;   // the general path reuses the piece buffers

            // ======================= general path: coefficients through scratch (any count) =======================
            const bool has_obj = (cur.fw & INC_FORM_MASK) == 0 && lane == 0;
            const bool obj_inc = __ballot_sync(FULL, has_obj) != 0;
            const int cbeg = beg + (obj_inc ? 1 : 0);
            const int mk = end - cbeg;
            const Scratch sc = (mk <= lay.sc_cap) ? sc_s : sc_g;
            double p0 = 0.0, q0 = 0.0, r0 = 0.0;
            if (phase == PH_P2) {
                r0 = w.fval[0];
                if (obj_inc) {
                    const int orb = bcast_i(cur.rbeg, 0), orl = bcast_i(cur.rlen, 0);
                    double dot;
                    if (orl < 0) dot = w.dd[orb];
                    else if (strict) dot = sparse_dot_seq(P, orb, orl, w.x);
                    else dot = sparse_dot_warp(P, orb, orl, w.x, lane);
                    p0 = bcast(cur.t2, 0);
                    q0 = 2 * dot + bcast(cur.qk, 0);
                    r0 = w.fval[0] - xk * (p0 * xk + q0);
                }
            }
            for (int base = 0; base < mk; base += 32) {
                const int i = base + lane;
                const bool mine = i < mk;
                const Meta mt = load_meta(P, cbeg + i, mine);
                double dot = 0.0;
                bool deferred = false;
                if (mine) {
                    if (mt.rlen < 0) dot = w.dd[mt.rbeg];
                    else if (strict || mt.rlen <= CD_LONG_ROW) dot = sparse_dot_seq(P, mt.rbeg, mt.rlen, w.x);
                    else deferred = true;
                }
                unsigned coop = __ballot_sync(FULL, deferred);
                while (coop) {
                    const int src = __ffs(coop) - 1;
                    coop &= coop - 1;
                    double v = sparse_dot_warp(P, bcast_i(mt.rbeg, src), bcast_i(mt.rlen, src), w.x, lane);
                    if (lane == src) dot = v;
                }
                if (mine) {
                    const int j = (int)(mt.fw & INC_FORM_MASK);
                    const double t1 = 2 * dot + mt.qk;
                    sc.p[i] = mt.t2;
                    sc.q[i] = t1;
                    sc.r[i] = w.fval[j] - xk * (mt.t2 * xk + t1);
                    sc.rel[i] = (int)((mt.fw >> INC_RELOP_SHIFT) & 3);
                }
            }
            __syncwarp();
            if (phase == PH_P1) {
                st.steps_p1++;
                double vmax = -QCQP_INF;
                int cz = 0;
                for (int i = lane; i < mk; i += 32) {
                    double p = sc.p[i], q = sc.q[i];
                    if (p == 0.0 && q == 0.0) continue;
                    cz++;
                    double v = violation_of(sc.rel[i], onevar_eval(p, q, sc.r[i], xk));
                    vmax = (v > vmax) ? v : vmax;
                }
                cz = warp_sum_i(cz);
                if (cz == 0) { st.status = QCQP_RUN_EMPTY_MAX; dead = true; }
                else {
                    const double viol = warp_max(vmax);
                    double new_viol = viol;
                    double ss = -tol, es = viol - viol_tol;
                    while (es - ss > tol) {
                        const double s = (ss + es) / 2;
                        double xi;
                        int err;
                        int ok = solve_level(w, sc, mk, s, 0.0, 0.0, 0.0, rng, lane, &xi, &err);
                        if (err) { st.status = err; dead = true; break; }
                        if (!ok) ss = s;
                        else { new_xi = xi; new_viol = s; es = s; }
                    }
                    if (!dead) {
                        if (new_viol < viol) { move = true; update_counter = 0; st.updates_p1++; }
                        else {
                            update_counter++;
                            if (update_counter == n) skip = true;
                        }
                    }
                }
            } else {
                st.steps_p2++;
                double xi;
                int err;
                int ok = solve_level(w, sc, mk, viol_p2, p0, q0, r0, rng, lane, &xi, &err);
                if (err) { st.status = err; dead = true; }
                else if (ok && fabs(xi - xk) > tol) { move = true; new_xi = xi; update_counter = 0; st.updates_p2++; }
                else {
                    update_counter++;
                    if (update_counter == n) phase = PH_DONE;   // converged
                }
            }
            if (move) {
                const double b = new_xi;
                for (int i = lane; i < mk; i += 32) {
                    int j = (int)(P.inc_form[cbeg + i] & INC_FORM_MASK);
                    w.fval[j] = sc.r[i] + b * (sc.p[i] * b + sc.q[i]);
                }
                if (lane == 0 && phase != PH_P1 && obj_inc) w.fval[0] = r0 + b * (p0 * b + q0);
            }
            if (lane == 0) w.x[k] = new_xi;
            if (grad) {
                if (move) grad_axpy(w, rows, dfirst, nd, n, ld, npad, new_xi - xk, lane);
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[sg]);
            }
            if (dead) phase = PH_DONE;
            __syncwarp();
        }
        // ---------------- end of sweep ----------------
        if (phase == PH_P1) {
            double mv = refresh_fvals(P, w, 1, mode, npad, lane);   // viol = max(prob.violations(x))  (qcqp.py:142)
            viol_last = mv;
            t++;
            if (!skip && st.updates_p1 == upd_before && !(viol_last < viol_tol) && t < prm.num_iters) {
                // A full sweep that moved nothing drew no random number either (a feasible probe always moves) and
                // update_counter is past n, so every remaining iteration of qcqp.py:110 would repeat it exactly.
                st.steps_skipped += (long long)(prm.num_iters - t) * n;
                t = prm.num_iters;
            }
        } else if (phase == PH_P2) {
            t++;
            if (prm.refresh_every > 0 && (t % prm.refresh_every) == 0) refresh_fvals(P, w, 0, mode, npad, lane);
        }
    }

    // ---------------- results: x, (f0.eval(x), max(violations(x))) as QCQP._improve returns them (qcqp.py:415-417) -------
    if (live) {
        double mv = refresh_fvals(P, w, 0, mode == MODE_GRAD ? MODE_FRESH : mode, npad, lane);   // final values: full evaluation
        for (int i = lane; i < n; i += 32) X[rr * n + i] = w.x[i];
        for (int i = lane; i < 624; i += 32) rngs[rr].key[i] = w.mt[i];
        if (lane == 0) {
            rngs[rr].pos = rng.pos;
            f0_out[rr] = w.fval[0];
            mv_out[rr] = (m > 0) ? mv : 0.0;
            if (stats_out) stats_out[rr] = st;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side: shared-memory plan and launch
// ---------------------------------------------------------------------------------------------------------
static unsigned align_up(unsigned v, unsigned a) { return (v + a - 1) / a * a; }

static int plan_layout(const qcqp_pack* p, int R, bool want_grad, CdLayout* L)
{
    const PackView& v = p->v;
    const int smem_max = max_smem_optin(p->device);
    const int sms = num_sms(p->device);
    CdLayout l;
    memset(&l, 0, sizeof(l));
    int hcap = 32;
    while (hcap < v.max_two) hcap <<= 1;
    l.hcap = hcap;
    // coefficient scratch in shared memory for every coordinate with <= 1024 incident forms; a coordinate beyond that
    // (circle packing's r: 20 702) uses the per-restart HBM block
    l.sc_cap = (v.max_inc <= 1024) ? (v.max_inc > 0 ? v.max_inc : 1) : (v.max_inc_small > 0 ? v.max_inc_small : 1);
    l.fval_smem = ((size_t)(v.m + 1) * 8 <= 24 * 1024) ? 1 : 0;
    unsigned o = 0;
    l.o_x = o; o += align_up((unsigned)v.n * 8, 16);
    l.o_fval = o; if (l.fval_smem) o += align_up((unsigned)(v.m + 1) * 8, 16);
    l.o_mt = o; o += 624 * 4;
    l.o_scp = o; o += align_up((unsigned)l.sc_cap * 8, 16);
    l.o_scq = o; o += align_up((unsigned)l.sc_cap * 8, 16);
    l.o_scr = o; o += align_up((unsigned)l.sc_cap * 8, 16);
    l.o_screl = o; o += align_up((unsigned)l.sc_cap * 4, 16);
    l.o_hx = o; o += (unsigned)hcap * 16;
    l.o_clo = o; o += align_up((unsigned)(hcap + 2) * 8, 16);
    l.o_chi = o; o += align_up((unsigned)(hcap + 2) * 8, 16);
    l.o_dd = o; o += align_up((unsigned)(v.n_dense > 0 ? v.n_dense : 1) * 8, 16);
    l.o_misc = o; o += 16;
    l.o_g = o;
    const unsigned g_bytes = (unsigned)v.n_dense * (unsigned)((v.n + 1) & ~1) * 8;
    const unsigned stride_nograd = align_up(o, 128);
    l.grad = 0;
    if (want_grad && v.n_dense > 0) {
        // the cached row dots must leave room for at least one restart next to a 2-stage ring
        if ((size_t)align_up(o + g_bytes, 128) + 256 + 2 * (size_t)v.n_dense * v.ld * 8 <= (size_t)smem_max) { l.grad = 1; o += g_bytes; }
    }
    l.warp_stride = l.grad ? align_up(o, 128) : stride_nograd;

    const unsigned stage_bytes = (unsigned)v.n_dense * v.ld * 8;
    int W, S = 0;
    if (v.n_dense > 0) {
        W = (R + sms - 1) / sms;          // spread the restarts over all SMs first, then share rows inside a CTA
        if (W < 1) W = 1;
        if (W > 7) W = 7;                 // 7 restart warps + the producer = 256 threads, 255 registers each
        // deepest ring (<= 4 stages) that still lets W restarts fit; otherwise fewer restarts per CTA
        for (;;) {
            S = 4;
            while (S > 2 && 256 + (size_t)align_up(S * stage_bytes, 128) + (size_t)W * l.warp_stride > (size_t)smem_max) S--;
            if (256 + (size_t)align_up(S * stage_bytes, 128) + (size_t)W * l.warp_stride <= (size_t)smem_max) break;
            if (W == 1) return fail(QCQP_ERR_CAPACITY, "qcqp_cd_improve: one restart plus the dense-row ring exceeds shared memory");
            W--;
        }
        unsigned fixed = 256 + align_up(S * stage_bytes, 128);
        l.off_bar = 0;
        l.off_ring = 256;
        l.off_warp0 = fixed;
    } else {
        if (l.warp_stride > (unsigned)smem_max)
            return fail(QCQP_ERR_CAPACITY, "qcqp_cd_improve: per-restart state exceeds shared memory (too many two-interval constraints on one coordinate)");
        int wmax = (int)(smem_max / l.warp_stride);
        W = wmax < 4 ? wmax : 4;
        if (W < 1) W = 1;
        l.off_bar = 0; l.off_ring = 0; l.off_warp0 = 0;
    }
    l.W = W; l.S = S;
    l.total = l.off_warp0 + (unsigned)W * l.warp_stride;
    *L = l;
    return QCQP_OK;
}

int cd_launch(qcqp_pack* p, const qcqp_cd_params* prm, const double* dX0, int R, qcqp_rng_state* drng, double* dX, double* df0,
              double* dmv, qcqp_cd_stats* dstats, cudaStream_t stream)
{
    if (R <= 0) return QCQP_OK;
    if (p->d_ctr) QCQP_CUDA_TRY(cudaMemsetAsync(p->d_ctr, 0, 8 * sizeof(unsigned long long), stream));
    CdLayout L;
    // CTA-per-restart kernel (cd_blk.cu): sparse problems whose coordinates meet many constraints; strict = 4 / 5 force it
    // (fast / strict summation) for A/B runs and the parity tests at small sizes
    if (prm->strict == 4 || prm->strict == 5 || ((prm->strict == 0 || prm->strict == 1) && !(prm->strict == 0 && p->lpc_ok) && blk_wanted(p))) {
        CdK kb;
        kb.num_iters = prm->num_iters; kb.viol_tol = prm->viol_tol; kb.tol = prm->tol; kb.phase1 = prm->phase1;
        kb.mode = (prm->strict == 1 || prm->strict == 5) ? MODE_STRICT : MODE_FRESH;
        kb.refresh_every = prm->refresh_every > 0 ? prm->refresh_every : 64;
        const char* ft = getenv("QCQP_BLK_THREADS");
        return blk_launch(p, kb, dX0, R, drng, dX, df0, dmv, dstats, stream, ft ? atoi(ft) : 0);
    }
    const int mode = (prm->strict == 1) ? MODE_STRICT : (prm->strict == 2 ? MODE_FRESH : MODE_GRAD);
    if (prm->strict == 0 && p->lpc_ok) {
        // separable problem (one single-coordinate constraint per coordinate): the lane-per-coordinate kernel
        CdK k0;
        k0.num_iters = prm->num_iters; k0.viol_tol = prm->viol_tol; k0.tol = prm->tol; k0.phase1 = prm->phase1; k0.mode = MODE_GRAD;
        k0.refresh_every = prm->refresh_every > 0 ? prm->refresh_every : 64;
        return lpc_launch(p, k0, dX0, R, drng, dX, df0, dmv, dstats, stream);
    }
    int rc = plan_layout(p, R, mode == MODE_GRAD, &L);
    if (rc != QCQP_OK) return rc;
    const PackView& v = p->v;
    size_t fval_bytes = L.fval_smem ? 0 : (size_t)R * (v.m + 1) * 8;
    size_t scr_bytes = (v.max_inc <= L.sc_cap) ? 0 : (size_t)R * 3 * v.max_inc * 8;
    size_t rel_bytes = (v.max_inc <= L.sc_cap) ? 0 : (size_t)R * v.max_inc * 4;
    size_t a1 = (fval_bytes + 255) & ~(size_t)255, a2 = (scr_bytes + 255) & ~(size_t)255;
    rc = ensure_workspace(p, a1 + a2 + rel_bytes + 256);
    if (rc != QCQP_OK) return rc;
    double* ws_fval = (double*)p->ws;
    double* ws_scr = (double*)((char*)p->ws + a1);
    int* ws_rel = (int*)((char*)p->ws + a1 + a2);
    CdK k;
    k.num_iters = prm->num_iters; k.viol_tol = prm->viol_tol; k.tol = prm->tol;
    k.phase1 = prm->phase1;
    k.mode = (mode == MODE_GRAD && !L.grad) ? MODE_FRESH : mode;   // no room for the cached row dots: fresh parallel dots
    k.refresh_every = prm->refresh_every > 0 ? prm->refresh_every : 64;
    QCQP_CUDA_TRY(cudaFuncSetAttribute(cd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    int blocks = (R + L.W - 1) / L.W;
    int threads = (L.W + (v.n_dense > 0 ? 1 : 0)) * 32;
    cd_kernel<<<blocks, threads, L.total, stream>>>(v, k, L, dX0, R, drng, dX, df0, dmv, dstats, ws_fval, ws_scr, ws_rel);
    QCQP_CUDA_TRY(cudaGetLastError());
    return QCQP_OK;
}

}  // namespace qcqp
