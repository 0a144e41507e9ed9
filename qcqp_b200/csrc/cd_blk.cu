// cd_blk.cu -- two-phase coordinate descent, ONE CTA PER RESTART, for problems whose coordinates meet many constraints
// (circle packing: 201 forms per centre coordinate, 20 702 for the radius).  Same reference functions as cd.cu
// (improve_coord_descent qcqp.py:181-192; coord_descent_phase1 :101-148; coord_descent_phase2 :152-178; get_onevar_func
// utilities.py:99-105; onevar_qcqp :241-288; get_feasible_intervals :198-232) and the same decisions, stream
// consumption and cached-f arithmetic as cd.cu's general path, so both kernels are held to the same oracle runs.
//
// Mapping.  Gauss-Seidel is sequential in k; what is parallel inside one coordinate step is the list of incident forms.
// A warp (cd.cu) walks 201 incidences in 7 rounds and sorts the ~199 holes of a centre coordinate with 36 shared-memory
// passes on its own; here the T threads of a CTA own one incidence each (strided when there are more):
//   coefficients (t2, t1, t0) from the cached f_j(x)  ->  feasible intervals  ->  per-thread fold, warp all-reduce, one
//   shared-memory exchange  ->  holes compacted with one shared-memory atomic per warp  ->  bitonic sort whose stages
//   with partner distance < the warp's region need only __syncwarp  ->  chunked prefix-max scan, one chunk per warp
//   ->  thread 0: minimiser + MT19937 draws  ->  every thread updates the cached f_j of its incidences.
// T is chosen from the number of restarts so that all of them are resident at once (128 registers per thread):
// T = 512 for R <= 148, 256 for R <= 296, else 128 (4 CTAs per SM).
#include "cd_holes.cuh"
#include "cd_shared.cuh"
#include "common.cuh"
#include "forms_eval.cuh"
#include "onevar.cuh"

namespace qcqp {

struct BlkLayout {
    int sc_cap;       // coefficient scratch entries in smem; coordinates with more incidences recompute them per probe
    int fval_smem;    // cached f_j in smem?
    int hcap;         // hole capacity (power of two, >= 64)
    unsigned o_x, o_fval, o_mt, o_scp, o_scq, o_scr, o_screl, o_hx, o_clo, o_chi, o_wfd, o_wfi, o_cmax, o_ccnt, o_redd, o_redi, o_ictl,
        o_dctl;
    unsigned total;
};

enum { BPH_P1 = 0, BPH_P2 = 1, BPH_DONE = 2 };
// ictl words: [0],[1] hole counters (by call parity), [2],[3] early-exit flags (by call parity), [4] found, [5] err, [6] nC
enum { IC_NH = 0, IC_FLAG = 2, IC_FOUND = 4, IC_ERR = 5 };

struct BlkCtx {
    double* x; double* fval;
    double* scp; double* scq; double* scr; int* screl;
    double2* hx; double* clo; double* chi;
    double* wfd; int* wfi;          // per-warp folds, two parities: [par][warp][2] and [par][warp][4]
    double* cmax; int* ccnt;        // per 32-hole chunk: max end, pieces found
    double* redd; int* redi;        // block reductions
    volatile int* ictl; double* dctl;
    int tid, warp, lane, sc_cap;
    int par;                        // call parity of solve_level
};

template <int NW>
__device__ __forceinline__ void blk_max_sum(const BlkCtx& c, double& vmax, int& isum)
{
    vmax = warp_max(vmax);
    isum = warp_sum_i(isum);
    if (c.lane == 0) { c.redd[c.warp] = vmax; c.redi[c.warp] = isum; }
    __syncthreads();
    double mx = c.redd[0];
    int s = c.redi[0];
#pragma unroll
    for (int i = 1; i < NW; i++) { const double v = c.redd[i]; mx = (v > mx) ? v : mx; s += c.redi[i]; }
    __syncthreads();
    vmax = mx; isum = s;
}

// (t2, t1, t0) of get_onevar_func (utilities.py:99-105) for incidence e; the row dot is summed by the owning thread in column
// order with separately rounded multiply/add (SciPy's csr_matvec order, so strict and fast mode coincide here); t0 from the
// cached f_j(x)
__device__ __forceinline__ void blk_coeffs(const PackView& P, const BlkCtx& c, int e, double xk, double& p, double& q, double& r, int& rel,
                                           int& j)
{
    const uint32_t fw = P.inc_form[e];
    j = (int)(fw & INC_FORM_MASK);
    rel = (int)((fw >> INC_RELOP_SHIFT) & 3);
    const double t2 = P.inc_t2[e], qk = P.inc_qk[e];
    const int rbeg = P.inc_rbeg[e], rlen = P.inc_rlen[e];
    double dot = 0.0;
    for (int t = rbeg; t < rbeg + rlen; t++) dot = dot + P.row_val[t] * c.x[P.row_col[t]];
    const double t1 = 2 * dot + qk;
    p = t2; q = t1;
    r = c.fval[j] - xk * (t2 * xk + t1);
}

// block bitonic sort of N2 (power of two, >= 64) holes by their start; stages whose partner distance stays inside a warp's
// region of the array synchronise with __syncwarp only
template <int T>
__device__ __forceinline__ void blk_sort_holes(double2* h, int N2, int tid)
{
    constexpr int NW = T / 32;
    const int warp = tid >> 5, lane = tid & 31;
    int chunk = N2 / NW;
    if (chunk < 64) chunk = 64;
    const int nlw = N2 / chunk;
    __syncthreads();
    for (int k = 2; k <= N2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (2 * j <= chunk) {
                if (warp < nlw) {
                    const int pbase = (warp * chunk) >> 1;
                    for (int t = lane; t < (chunk >> 1); t += 32) {
                        const int tt = pbase + t;
                        const int i = ((tt & ~(j - 1)) << 1) | (tt & (j - 1));
                        const int p = i | j;
                        const double2 a = h[i], b = h[p];
                        const bool up = ((i & k) == 0);
                        if ((a.x > b.x) == up && a.x != b.x) { h[i] = b; h[p] = a; }
                    }
                }
                __syncwarp();
            } else {
                __syncthreads();
                for (int tt = tid; tt < (N2 >> 1); tt += T) {
                    const int i = ((tt & ~(j - 1)) << 1) | (tt & (j - 1));
                    const int p = i | j;
                    const double2 a = h[i], b = h[p];
                    const bool up = ((i & k) == 0);
                    if ((a.x > b.x) == up && a.x != b.x) { h[i] = b; h[p] = a; }
                }
                __syncthreads();
            }
        }
    }
    __syncthreads();
}

// one 32-hole chunk of the sorted list (lane = position): is my hole's start the right end of a feasible piece, and where
// does that piece begin (scan_hole_chunk of cd_holes.cuh, with the carry taken from the per-chunk maxima)
__device__ __forceinline__ bool blk_chunk_eval(const BlkCtx& c, const Fold& f, int ch, int N2, double& a, double& b, double& M)
{
    const double2 t = c.hx[(ch << 5) + c.lane];
    a = t.x; b = t.y;
    double inc = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(FULL, inc, o);
        if (c.lane >= o && u > inc) inc = u;
    }
    const double exc = __shfl_up_sync(FULL, inc, 1);
    double carry = -QCQP_INF;
    for (int cc = c.lane; cc < ch; cc += 32) { const double v = c.cmax[cc]; carry = (v > carry) ? v : carry; }
    carry = warp_max(carry);
    if (f.L > carry) carry = f.L;
    M = carry;
    if (c.lane > 0 && exc > M) M = exc;
    const double ap = __shfl_up_sync(FULL, a, 1);
    const bool tiep = (c.lane > 0) ? (ap == a) : (ch > 0 && c.hx[(ch << 5) - 1].x == a);
    double an = __shfl_down_sync(FULL, a, 1);
    if (c.lane == 31) an = (((ch + 1) << 5) < N2) ? c.hx[(ch + 1) << 5].x : QCQP_INF;
    return !tiep && (an != a) && (M < a) && (a < f.H);
}

// onevar_qcqp(f0 = (p0, q0, r0), the mk constraints of this coordinate, level s) by the whole CTA.  Returns found
// (block-uniform); *xout / *err valid in every thread.
template <int T>
__device__ __forceinline__ int blk_solve_level(const PackView& P, BlkCtx& c, int cbeg, int mk, bool use_sc, double xk, double s, double p0,
                                               double q0, double r0, MtRng& rng, double* xout, int* err)
{
    constexpr int NW = T / 32;
    *xout = 0.0;
    *err = 0;
    const int par = c.par;
    c.par ^= 1;
    volatile int* nh_ctr = c.ictl + IC_NH + par;
    volatile int* flag = c.ictl + IC_FLAG + par;
    Fold f;
    f.init();
    const unsigned lt = (1u << c.lane) - 1u;
    int it = 0;
    for (int base = c.warp * 32; base < mk; base += T, it++) {
        const int i = base + c.lane;
        bool hole = false;
        Hole hh;
        hh.a = hh.b = 0.0;
        if (i < mk) {
            double p, q, r;
            int rel, j;
            if (use_sc) { p = c.scp[i]; q = c.scq[i]; r = c.scr[i]; rel = c.screl[i]; }
            else blk_coeffs(P, c, cbeg + i, xk, p, q, r, rel, j);
            if (!(p == 0.0 && q == 0.0)) {   // nfs filter of qcqp.py:116,166
                Ival I[2];
                const int cnt = feasible_intervals(p, q, r, rel, s, I);
                hole = fold_constraint(f, cnt, I, &hh);
            }
        }
        const unsigned hb = __ballot_sync(FULL, hole);
        if (hb) {
            int pos0 = 0;
            if (c.lane == 0) pos0 = atomicAdd((int*)nh_ctr, __popc(hb));
            pos0 = __shfl_sync(FULL, pos0, 0);
            if (hole) c.hx[pos0 + __popc(hb & lt)] = make_double2(hh.a, hh.b);
        }
        // a constraint with no feasible point at this level, or an empty partial intersection of the singles: the final set
        // is empty whatever the other warps find (long lists only: the radius of circle packing meets 20 701 constraints)
        if (base + T < mk) {
            bool bad = __any_sync(FULL, f.nempty > 0);
            if (!bad && (it & 3) == 3) {
                const double Lp = warp_max(f.L), Hp = -warp_max(-f.H);
                bad = !(Lp < Hp);
            }
            if (bad) *flag = 1;
            if (*flag) break;
        }
    }
    fold_allreduce(f);
    if (c.lane == 0) {
        double* wd = c.wfd + (size_t)(par * NW + c.warp) * 2;
        int* wi = c.wfi + (size_t)(par * NW + c.warp) * 4;
        wd[0] = f.L; wd[1] = f.H; wi[0] = f.mu; wi[1] = f.m1; wi[2] = f.mcnt; wi[3] = f.nempty;
    }
    __syncthreads();   // #1: folds, holes and their count are in shared memory
    f.init();
#pragma unroll
    for (int w = 0; w < NW; w++) {
        const double* wd = c.wfd + (size_t)(par * NW + w) * 2;
        const int* wi = c.wfi + (size_t)(par * NW + w) * 4;
        f.merge(wd[0], wd[1], wi[0], wi[1], wi[2], wi[3]);
    }
    const int nh = *nh_ctr;
    if (c.tid == 0) { c.ictl[IC_NH + (par ^ 1)] = 0; c.ictl[IC_FLAG + (par ^ 1)] = 0; }   // the next call's slots
    if (f.mcnt > 0 && (f.nempty > 0 || !(f.L < f.H))) return 0;   // the running total never gets full

    int nC = 0;          // meaningful in thread 0
    if (f.mcnt == 0) {
        if (c.tid == 0) { c.clo[0] = -QCQP_INF; c.chi[0] = QCQP_INF; }   // only the sentinel pair: all of R
        nC = 1;
    } else if (nh <= 32) {
        if (c.warp == 0) {
            HoleScan hs;
            hs.carryM = f.L; hs.prevA = 0.0; hs.havePrev = false; hs.blocked = false; hs.stH = -QCQP_INF; hs.nC = 0;
            if (nh > 0) {
                double a = QCQP_INF, b = QCQP_INF;
                if (c.lane < nh) { const double2 t = c.hx[c.lane]; a = t.x; b = t.y; }
                bitonic_sort_holes_reg(a, b, c.lane);
                scan_hole_chunk(c.clo, c.chi, f, a, b, QCQP_INF, hs, c.lane);
            }
            nC = hs.nC;
            if (f.mu == 1 && f.H < QCQP_INF) {
                const bool blocked = __any_sync(FULL, hs.blocked);
                double st = warp_max(hs.stH);
                if (f.L > st) st = f.L;
                if (!blocked) {
                    if (c.lane == 0) { c.clo[nC] = st; c.chi[nC] = f.H; }
                    nC++;
                }
            }
            __syncwarp();
        }
    } else {
        int N2 = 64;
        while (N2 < nh) N2 <<= 1;
        for (int i = nh + c.tid; i < N2; i += T) c.hx[i] = make_double2(QCQP_INF, QCQP_INF);   // pads are inert
        blk_sort_holes<T>(c.hx, N2, c.tid);
        const int nchunk = (nh + 31) >> 5;
        for (int ch = c.warp; ch < nchunk; ch += NW) {
            const double mx = warp_max(c.hx[(ch << 5) + c.lane].y);
            if (c.lane == 0) c.cmax[ch] = mx;
        }
        __syncthreads();
        bool blocked = false;
        double stH = -QCQP_INF;
        for (int ch = c.warp; ch < nchunk; ch += NW) {
            double a, b, M;
            const bool valid = blk_chunk_eval(c, f, ch, N2, a, b, M);
            const unsigned vb = __ballot_sync(FULL, valid);
            if (c.lane == 0) c.ccnt[ch] = __popc(vb);
            if (a <= f.H && f.H <= b) blocked = true;
            if (b < f.H && b > stH) stH = b;
        }
        blocked = __any_sync(FULL, blocked);
        stH = warp_max(stH);
        if (c.lane == 0) { c.redi[c.warp] = blocked ? 1 : 0; c.redd[c.warp] = stH; }
        __syncthreads();
        for (int ch = c.warp; ch < nchunk; ch += NW) {
            double a, b, M;
            const bool valid = blk_chunk_eval(c, f, ch, N2, a, b, M);
            const unsigned vb = __ballot_sync(FULL, valid);
            int off = 0;
            for (int cc = c.lane; cc < ch; cc += 32) off += c.ccnt[cc];
            off = warp_sum_i(off);
            if (valid) {
                const int pos = off + __popc(vb & lt);
                c.clo[pos] = M; c.chi[pos] = a;
            }
        }
        __syncthreads();
        if (c.tid == 0) {
            for (int ch = 0; ch < nchunk; ch++) nC += c.ccnt[ch];
            if (f.mu == 1 && f.H < QCQP_INF) {
                bool bl = false;
                double st = f.L;
                for (int w = 0; w < NW; w++) { bl = bl || (c.redi[w] != 0); if (c.redd[w] > st) st = c.redd[w]; }
                if (!bl) { c.clo[nC] = st; c.chi[nC] = f.H; nC++; }
            }
        }
    }
    if (c.tid == 0) {
        int e = 0;
        double xv = 0.0;
        const int found = choose_point(p0, q0, r0, c.clo, c.chi, nC, rng, &xv, &e);
        c.dctl[0] = xv;
        c.ictl[IC_FOUND] = found;
        c.ictl[IC_ERR] = e;
    }
    __syncthreads();   // #2
    *xout = c.dctl[0];
    *err = c.ictl[IC_ERR];
    return c.ictl[IC_FOUND];
}

// cached f_j(x) from scratch for forms j >= j0 (warps take blocks of 32 forms, same per-form summation as cd.cu's
// refresh_fvals); returns the max constraint violation
template <int T>
__device__ __forceinline__ double blk_refresh_fvals(const PackView& P, const BlkCtx& c, int j0, bool strict)
{
    constexpr int NW = T / 32;
    double* fv = c.fval;
    __syncthreads();
    for (int base = j0 + c.warp * 32; base <= P.m; base += NW * 32) {
        const int hi = (base + 31 < P.m) ? base + 31 : P.m;
        eval_forms(P, c.x, base, hi, strict, c.lane, [&](int j, double v) { fv[j] = v; }, false);
    }
    __syncthreads();
    double mv = -QCQP_INF;
    for (int j = 1 + c.tid; j <= P.m; j += T) {
        const double v = violation_of(P.relop[j], fv[j]);
        mv = (v > mv) ? v : mv;
    }
    int dummy = 0;
    blk_max_sum<NW>(c, mv, dummy);
    return mv;
}

template <int T>
__global__ void __launch_bounds__(T, 512 / T) cd_blk_kernel(PackView P, CdK prm, BlkLayout lay, const double* __restrict__ X0, int R,
                                                            qcqp_rng_state* rngs, double* __restrict__ X, double* __restrict__ f0_out,
                                                            double* __restrict__ mv_out, qcqp_cd_stats* stats_out, double* ws_fval)
{
    constexpr int NW = T / 32;
    extern __shared__ __align__(128) unsigned char smem[];
    const int n = P.n, m = P.m;
    const size_t rr = blockIdx.x;
    BlkCtx c;
    c.tid = threadIdx.x; c.warp = c.tid >> 5; c.lane = c.tid & 31;
    c.x = reinterpret_cast<double*>(smem + lay.o_x);
    c.fval = lay.fval_smem ? reinterpret_cast<double*>(smem + lay.o_fval) : ws_fval + rr * (size_t)(m + 1);
    uint32_t* mt = reinterpret_cast<uint32_t*>(smem + lay.o_mt);
    c.scp = reinterpret_cast<double*>(smem + lay.o_scp);
    c.scq = reinterpret_cast<double*>(smem + lay.o_scq);
    c.scr = reinterpret_cast<double*>(smem + lay.o_scr);
    c.screl = reinterpret_cast<int*>(smem + lay.o_screl);
    c.hx = reinterpret_cast<double2*>(smem + lay.o_hx);
    c.clo = reinterpret_cast<double*>(smem + lay.o_clo);
    c.chi = reinterpret_cast<double*>(smem + lay.o_chi);
    c.wfd = reinterpret_cast<double*>(smem + lay.o_wfd);
    c.wfi = reinterpret_cast<int*>(smem + lay.o_wfi);
    c.cmax = reinterpret_cast<double*>(smem + lay.o_cmax);
    c.ccnt = reinterpret_cast<int*>(smem + lay.o_ccnt);
    c.redd = reinterpret_cast<double*>(smem + lay.o_redd);
    c.redi = reinterpret_cast<int*>(smem + lay.o_redi);
    c.ictl = reinterpret_cast<volatile int*>(smem + lay.o_ictl);
    c.dctl = reinterpret_cast<double*>(smem + lay.o_dctl);
    c.sc_cap = lay.sc_cap;
    c.par = 0;
    const int tid = c.tid;

    for (int i = tid; i < n; i += T) c.x[i] = X0[rr * n + i];
    for (int i = tid; i < 624; i += T) mt[i] = rngs[rr].key[i];
    if (tid < 8) c.ictl[tid] = 0;
    MtRng rng;
    rng.key = mt;
    rng.pos = rngs[rr].pos;
    __syncthreads();

    qcqp_cd_stats st;
    st.steps_p1 = st.steps_p2 = st.updates_p1 = st.updates_p2 = st.steps_skipped = 0;
    st.sweeps_p1 = st.sweeps_p2 = 0; st.status = QCQP_RUN_OK; st.ran_phase2 = 0;
    const bool strict = (prm.mode == MODE_STRICT);
    const double tol = prm.tol, viol_tol = prm.viol_tol;
    int phase = BPH_P1;
    int t = 0;                    // sweeps done in the current phase
    long long update_counter = 0;
    double viol_last = QCQP_INF;  // phase 1
    double viol_p2 = 0.0;         // phase 2: frozen at entry (qcqp.py:157)
    const bool p1_over = !prm.phase1;

    for (;;) {
        // ---------------- sweep boundary: phase transitions (block-uniform) ----------------
        if (phase == BPH_P1) {
            if (p1_over || t >= prm.num_iters || viol_last < viol_tol) {
                // improve_coord_descent: if max(prob.violations(x)) < viol_tol: phase 2   (qcqp.py:189-190)
                const double mv = blk_refresh_fvals<T>(P, c, 0, strict);
                if (m == 0) { st.status = QCQP_RUN_EMPTY_MAX; phase = BPH_DONE; }
                else if (mv < viol_tol) { phase = BPH_P2; viol_p2 = mv; t = 0; update_counter = 0; st.ran_phase2 = 1; }
                else phase = BPH_DONE;
            } else if (t == 0) {
                blk_refresh_fvals<T>(P, c, 1, strict);
            }
        }
        if (phase == BPH_P2 && t >= prm.num_iters) phase = BPH_DONE;
        if (phase == BPH_DONE) break;
        if (phase == BPH_P1) st.sweeps_p1++;
        if (phase == BPH_P2) st.sweeps_p2++;
        bool skip = false;   // phase 1 'failed' break: the rest of this sweep is not executed (qcqp.py:138-141)
        const long long upd_before = st.updates_p1;

        for (int k = 0; k < n && phase != BPH_DONE && !skip; k++) {
            const int beg = P.inc_ptr[k], end = P.inc_ptr[k + 1];
            const bool obj_inc = (end > beg) && ((P.inc_form[beg] & INC_FORM_MASK) == 0);   // the objective, when incident, comes first
            const int cbeg = beg + (obj_inc ? 1 : 0);
            const int mk = end - cbeg;
            const bool use_sc = mk <= c.sc_cap;
            const double xk = c.x[k];
            double new_xi = xk;
            bool move = false;
            bool dead = false;   // the reference would have raised: stop this restart, leave x as it was
            double p0 = 0.0, q0 = 0.0, r0 = 0.0;
            if (phase == BPH_P2) {
                r0 = c.fval[0];
                if (obj_inc) {
                    const int orb = P.inc_rbeg[beg], orl = P.inc_rlen[beg];
                    double dot = 0.0;
                    if (strict || orl <= 64) {
                        for (int u = orb; u < orb + orl; u++) dot = dot + P.row_val[u] * c.x[P.row_col[u]];
                    } else {
                        double part = 0.0;
                        for (int u = orb + tid; u < orb + orl; u += T) part = fma(P.row_val[u], c.x[P.row_col[u]], part);
                        part = warp_sum(part);
                        if (c.lane == 0) c.redd[c.warp] = part;
                        __syncthreads();
                        for (int w = 0; w < NW; w++) dot += c.redd[w];
                        __syncthreads();
                    }
                    p0 = P.inc_t2[beg];
                    q0 = 2 * dot + P.inc_qk[beg];
                    r0 = c.fval[0] - xk * (p0 * xk + q0);
                }
            }
            // coefficients of my incidences (thread tid owns i = tid, tid + T, ...: the scratch entries are private to it)
            if (use_sc) {
                for (int i = tid; i < mk; i += T) {
                    double p, q, r;
                    int rel, j;
                    blk_coeffs(P, c, cbeg + i, xk, p, q, r, rel, j);
                    c.scp[i] = p; c.scq[i] = q; c.scr[i] = r; c.screl[i] = rel;
                }
            }
            if (phase == BPH_P1) {
                st.steps_p1++;
                double vmax = -QCQP_INF;
                int cz = 0;
                for (int i = tid; i < mk; i += T) {
                    double p, q, r;
                    int rel, j;
                    if (use_sc) { p = c.scp[i]; q = c.scq[i]; r = c.scr[i]; rel = c.screl[i]; }
                    else blk_coeffs(P, c, cbeg + i, xk, p, q, r, rel, j);
                    if (p == 0.0 && q == 0.0) continue;
                    cz++;
                    const double v = violation_of(rel, onevar_eval(p, q, r, xk));
                    vmax = (v > vmax) ? v : vmax;
                }
                blk_max_sum<NW>(c, vmax, cz);
                if (cz == 0) { st.status = QCQP_RUN_EMPTY_MAX; dead = true; }
                else {
                    const double viol = vmax;
                    double new_viol = viol;
                    double ss = -tol, es = viol - viol_tol;
                    while (es - ss > tol) {
                        const double s = (ss + es) / 2;
                        double xi;
                        int err;
                        const int ok = blk_solve_level<T>(P, c, cbeg, mk, use_sc, xk, s, 0.0, 0.0, 0.0, rng, &xi, &err);
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
                const int ok = blk_solve_level<T>(P, c, cbeg, mk, use_sc, xk, viol_p2, p0, q0, r0, rng, &xi, &err);
                if (err) { st.status = err; dead = true; }
                else if (ok && fabs(xi - xk) > tol) { move = true; new_xi = xi; update_counter = 0; st.updates_p2++; }
                else {
                    update_counter++;
                    if (update_counter == n) phase = BPH_DONE;   // converged
                }
            }
            if (move) {
                // f_j(x) = t0 + b (t2 b + t1) for every incident form
                const double b = new_xi;
                for (int i = tid; i < mk; i += T) {
                    double p, q, r;
                    int rel, j;
                    if (use_sc) { p = c.scp[i]; q = c.scq[i]; r = c.scr[i]; j = (int)(P.inc_form[cbeg + i] & INC_FORM_MASK); }
                    else blk_coeffs(P, c, cbeg + i, xk, p, q, r, rel, j);
                    c.fval[j] = r + b * (p * b + q);
                }
                if (tid == 0) {
                    if (phase != BPH_P1 && obj_inc) c.fval[0] = r0 + b * (p0 * b + q0);
                    c.x[k] = new_xi;
                }
            }
            if (dead) phase = BPH_DONE;
            __syncthreads();   // x and the cached f_j are current for the next coordinate
        }
        // ---------------- end of sweep ----------------
        if (phase == BPH_P1) {
            const double mv = blk_refresh_fvals<T>(P, c, 1, strict);   // viol = max(prob.violations(x))  (qcqp.py:142)
            viol_last = mv;
            t++;
            if (!skip && st.updates_p1 == upd_before && !(viol_last < viol_tol) && t < prm.num_iters) {
                // a full sweep that moved nothing drew no random number either: every remaining iteration of qcqp.py:110
                // would repeat it exactly (cd.cu has the argument)
                st.steps_skipped += (long long)(prm.num_iters - t) * n;
                t = prm.num_iters;
            }
        } else if (phase == BPH_P2) {
            t++;
            if (prm.refresh_every > 0 && (t % prm.refresh_every) == 0) blk_refresh_fvals<T>(P, c, 0, strict);
        }
    }

    // ---------------- results: x, (f0.eval(x), max(violations(x))) as QCQP._improve returns them (qcqp.py:415-417) -------
    const double mv = blk_refresh_fvals<T>(P, c, 0, strict);
    for (int i = tid; i < n; i += T) X[rr * n + i] = c.x[i];
    for (int i = tid; i < 624; i += T) rngs[rr].key[i] = mt[i];
    if (tid == 0) {
        rngs[rr].pos = rng.pos;
        f0_out[rr] = c.fval[0];
        mv_out[rr] = (m > 0) ? mv : 0.0;
        if (stats_out) stats_out[rr] = st;
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static unsigned blk_align(unsigned v, unsigned a) { return (v + a - 1) / a * a; }

bool blk_wanted(const qcqp_pack* p)
{
    const PackView& v = p->v;
    if (v.n_dense > 0 || v.n <= 0) return false;
    return p->info.incidences >= (int64_t)48 * v.n;   // on average a warp's worth of incident forms (or more) per coordinate
}

template <int T>
static int blk_launch_t(qcqp_pack* p, const CdK& k, const BlkLayout& L, const double* dX0, int R, qcqp_rng_state* drng, double* dX, double* df0,
                        double* dmv, qcqp_cd_stats* dstats, double* ws_fval, cudaStream_t stream)
{
    QCQP_CUDA_TRY(cudaFuncSetAttribute(cd_blk_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    cd_blk_kernel<T><<<R, T, L.total, stream>>>(p->v, k, L, dX0, R, drng, dX, df0, dmv, dstats, ws_fval);
    QCQP_CUDA_TRY(cudaGetLastError());
    return QCQP_OK;
}

int blk_launch(qcqp_pack* p, const CdK& k, const double* dX0, int R, qcqp_rng_state* drng, double* dX, double* df0, double* dmv,
               qcqp_cd_stats* dstats, cudaStream_t stream, int force_threads)
{
    const PackView& v = p->v;
    if (v.n_dense > 0) return fail(QCQP_ERR_INVALID, "qcqp_cd_improve: the CTA-per-restart kernel handles sparse forms only");
    const int sms = num_sms(p->device);
    int T = (R <= sms) ? 512 : (R <= 2 * sms ? 256 : 128);
    if (force_threads == 128 || force_threads == 256 || force_threads == 512) T = force_threads;
    const int NW = T / 32;
    BlkLayout l;
    memset(&l, 0, sizeof(l));
    int hcap = 64;
    while (hcap < v.max_two) hcap <<= 1;
    l.hcap = hcap;
    l.sc_cap = (v.max_inc <= 1024) ? (v.max_inc > 0 ? v.max_inc : 1) : (v.max_inc_small > 0 ? v.max_inc_small : 1);
    l.fval_smem = ((size_t)(v.m + 1) * 8 <= 16 * 1024) ? 1 : 0;
    unsigned o = 0;
    l.o_x = o; o += blk_align((unsigned)v.n * 8, 16);
    l.o_fval = o; if (l.fval_smem) o += blk_align((unsigned)(v.m + 1) * 8, 16);
    l.o_mt = o; o += 624 * 4;
    l.o_scp = o; o += blk_align((unsigned)l.sc_cap * 8, 16);
    l.o_scq = o; o += blk_align((unsigned)l.sc_cap * 8, 16);
    l.o_scr = o; o += blk_align((unsigned)l.sc_cap * 8, 16);
    l.o_screl = o; o += blk_align((unsigned)l.sc_cap * 4, 16);
    l.o_hx = o; o += (unsigned)hcap * 16;
    l.o_clo = o; o += blk_align((unsigned)(hcap + 2) * 8, 16);
    l.o_chi = o; o += blk_align((unsigned)(hcap + 2) * 8, 16);
    l.o_wfd = o; o += (unsigned)(2 * NW * 2) * 8;
    l.o_wfi = o; o += (unsigned)(2 * NW * 4) * 4;
    l.o_cmax = o; o += blk_align((unsigned)(hcap / 32) * 8, 16);
    l.o_ccnt = o; o += blk_align((unsigned)(hcap / 32) * 4, 16);
    l.o_redd = o; o += (unsigned)NW * 8;
    l.o_redi = o; o += blk_align((unsigned)NW * 4, 16);
    l.o_ictl = o; o += 32;
    l.o_dctl = o; o += 16;
    l.total = blk_align(o, 128);
    if (l.total > (unsigned)max_smem_optin(p->device))
        return fail(QCQP_ERR_CAPACITY, "qcqp_cd_improve: per-restart state exceeds shared memory (too many two-interval constraints on one coordinate)");
    const size_t fval_bytes = l.fval_smem ? 0 : (size_t)R * (v.m + 1) * 8;
    int rc = ensure_workspace(p, fval_bytes + 256);
    if (rc != QCQP_OK) return rc;
    double* ws_fval = (double*)p->ws;
    if (T == 512) return blk_launch_t<512>(p, k, l, dX0, R, drng, dX, df0, dmv, dstats, ws_fval, stream);
    if (T == 256) return blk_launch_t<256>(p, k, l, dX0, R, drng, dX, df0, dmv, dstats, ws_fval, stream);
    return blk_launch_t<128>(p, k, l, dX0, R, drng, dX, df0, dmv, dstats, ws_fval, stream);
}

}  // namespace qcqp
