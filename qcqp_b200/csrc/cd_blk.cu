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
// Phase 1 (the ~14 bisection probes per coordinate of qcqp.py:113-140, the whole of circle packing's default run) does not repeat
// that CTA-wide probe 14 times:
//   * constraints proven inert over the whole bisection (concave, negative discriminant at the lowest level) are compacted away;
//   * a probe over the compact list is WARP-local and sort-free (blk_warp_probe), each warp of the CTA probing another level;
//   * a "solid" level (the feasible sets share no open interval) certifies every lower level infeasible -- feasible sets only grow
//     with the level -- so the next non-trivial probe of the reference's loop is found by a search over the chain of levels,
//     deepest first: a coordinate that cannot move costs one probe round.  Every step is exact; DESIGN.md 4.1.c has the argument.
// Lists with more than 128 kept constraints (the radius) keep the CTA-wide sequential probes.
// T is chosen from the number of restarts so that all of them are resident at once (128 registers per thread):
// T = 512 for R <= 148, 256 for R <= 296, else 128 (4 CTAs per SM; the 64-register build with 8 CTAs per SM only when every probe is
// CTA-wide, QCQP_BLK_WARP=0).
#include "cd_holes.cuh"
#include "cd_shared.cuh"
#include "common.cuh"
#include "forms_eval.cuh"
#include "onevar.cuh"

namespace qcqp {

#ifdef BLK_PROF
#define STAMP(c, slot) do { if ((c).on) { const long long now__ = clock64(); (c).prof[(c).base + (slot)] += now__ - (c).last; (c).last = now__; } } while (0)
#else
#define STAMP(c, slot) do { } while (0)
#endif

struct BlkLayout {
    int sc_cap;       // coefficient scratch entries in smem; coordinates with more incidences use the per-restart HBM block
    int fval_smem;    // cached f_j in smem?
    int hcap;         // hole capacity (power of two, >= 64)
    int act_cap;      // phase 1: capacity of the compacted list of constraints that can still matter in this coordinate's bisection
    unsigned o_ap, o_aq, o_ar, o_arel;
    unsigned o_whx, o_wclo, o_wchi, o_wres, o_wlev;   // per-warp hole / piece buffers of the speculative phase-1 probes, and their results
    unsigned o_x, o_fval, o_mt, o_scp, o_scq, o_scr, o_screl, o_hx, o_clo, o_chi, o_wfd, o_wfi, o_cmax, o_ccnt, o_redd, o_redi, o_ictl,
        o_dctl;
    unsigned total;
};

enum { BPH_P1 = 0, BPH_P2 = 1, BPH_DONE = 2 };
// ictl words: [0],[1] hole counters (by call parity), [2],[3] early-exit flags (by call parity), [4] found, [5] err,
// [6] phase-1 compaction: constraints kept, [7] constraints proven inert
enum { IC_NH = 0, IC_FLAG = 2, IC_FOUND = 4, IC_ERR = 5, IC_NACT = 6, IC_NINERT = 7 };

struct BlkCtx {
    double* x; double* fval;
    double2* hx; double* clo; double* chi;
    double* wfd; int* wfi;          // per-warp folds, two parities: [par][warp][2] and [par][warp][4]
    double* cmax; int* ccnt;        // per 32-hole chunk: max end, pieces found
    double* redd; int* redi;        // block reductions
    volatile int* ictl; double* dctl;
    int tid, warp, lane;
    int par;                        // call parity of solve_level
#ifdef BLK_PROF
    long long* prof; long long last; int on, base;
#endif
};

// one-variable coefficients of the incident forms of the current coordinate: slot i belongs to thread i mod T.
// rj = relop | form << 2.  Shared memory for coordinates with <= sc_cap incident forms, the restart's HBM block beyond.
struct Scr {
    double* p; double* q; double* r; int* rj;
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

// static metadata of one incidence (form word, diagonal, q_j[k], off-diagonal row with its first entry); fetched ahead of
// its use, so that the only global round trip left inside a step is the gather of the cached f_j
struct PMeta {
    uint32_t fw;
    int rbeg, rlen, col0;
    double t2, qk, val0;
};
__device__ __forceinline__ PMeta blk_load_meta(const PackView& P, int e, bool valid)
{
    PMeta mt;
    mt.fw = 0xffffffffu; mt.rbeg = 0; mt.rlen = 0; mt.col0 = 0; mt.t2 = 0.0; mt.qk = 0.0; mt.val0 = 0.0;
    if (valid) {
        mt.fw = P.inc_form[e]; mt.rbeg = P.inc_rbeg[e]; mt.rlen = P.inc_rlen[e];
        mt.t2 = P.inc_t2[e]; mt.qk = P.inc_qk[e];
    }
    return mt;
}
// second phase of the prefetch, issued once rbeg / rlen have arrived: the first off-diagonal entry of the row
__device__ __forceinline__ void blk_load_row0(const PackView& P, PMeta& mt)
{
    if (mt.rlen > 0) { mt.col0 = P.row_col[mt.rbeg]; mt.val0 = P.row_val[mt.rbeg]; }
}

// (t2, t1, t0) of get_onevar_func (utilities.py:99-105) for one incidence; the row dot is summed by the owning thread in
// column order with separately rounded multiply/add (SciPy's csr_matvec order, so strict and fast mode coincide here); t0
// from the cached f_j(x) (fv, gathered by the caller so that the loads of several slots are in flight together).
// The objective (form 0) is never a constraint: its slot holds (0, 0, 0), which the nfs filter drops.
__device__ __forceinline__ void blk_coeffs_from(const PackView& P, const BlkCtx& c, const PMeta& mt, double fv, double xk, double& p,
                                                double& q, double& r, int& rj)
{
    const int j = (int)(mt.fw & INC_FORM_MASK);
    rj = (int)((mt.fw >> INC_RELOP_SHIFT) & 3) | (j << 2);
    double dot = 0.0;
    if (mt.rlen > 0) {
        dot = dot + mt.val0 * c.x[mt.col0];
        for (int t = mt.rbeg + 1; t < mt.rbeg + mt.rlen; t++) dot = dot + P.row_val[t] * c.x[P.row_col[t]];
    }
    const double t1 = 2 * dot + mt.qk;
    p = mt.t2; q = t1;
    r = fv - xk * (mt.t2 * xk + t1);
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

// onevar_qcqp(f0 = (p0, q0, r0), the cnt incident forms of this coordinate in sc, level s) by the whole CTA.  Returns found
// (block-uniform); *xout / *err valid in every thread.
template <int T>
__device__ __forceinline__ int blk_solve_level(BlkCtx& c, const Scr& sc, int cnt, double s, double p0, double q0, double r0, MtRng& rng,
                                               double* xout, int* err)
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
    {
        // my slots, one per round, the next one's coefficients in flight while this one is folded
        int i = c.warp * 32 + c.lane;
        double pn = 0.0, qn = 0.0, rn = 0.0;
        int rjn = 0;
        if (i < cnt) { pn = sc.p[i]; qn = sc.q[i]; rn = sc.r[i]; rjn = sc.rj[i]; }
        int it = 0;
        for (int base = c.warp * 32; base < cnt; base += T, it++) {
            const double p = pn, q = qn, r = rn;
            const int rel = rjn & 3;
            i += T;
            pn = 0.0; qn = 0.0; rn = 0.0; rjn = 0;
            if (i < cnt) { pn = sc.p[i]; qn = sc.q[i]; rn = sc.r[i]; rjn = sc.rj[i]; }
            bool hole = false;
            Hole hh;
            hh.a = hh.b = 0.0;
            if (!(p == 0.0 && q == 0.0)) {   // nfs filter of qcqp.py:116,166 (also drops the objective's and the empty slots)
                Ival I[2];
                const int ni = feasible_intervals(p, q, r, rel, s, I);
                hole = fold_constraint(f, ni, I, &hh);
            }
            const unsigned hb = __ballot_sync(FULL, hole);
            if (hb) {
                int pos0 = 0;
                if (c.lane == 0) pos0 = atomicAdd((int*)nh_ctr, __popc(hb));
                pos0 = __shfl_sync(FULL, pos0, 0);
                if (hole) c.hx[pos0 + __popc(hb & lt)] = make_double2(hh.a, hh.b);
            }
            // a constraint with no feasible point at this level, or an empty partial intersection of the singles: the final
            // set is empty whatever the other warps find (long lists only: circle packing's radius meets 20 701 constraints)
            if ((it & 3) == 3 && base + T < cnt) {
                bool bad = __any_sync(FULL, f.nempty > 0);
                if (!bad) {
                    const double Lp = warp_max(f.L), Hp = -warp_max(-f.H);
                    bad = !(Lp < Hp);
                }
                if (bad) *flag = 1;
                if (*flag) break;
            }
        }
    }
    STAMP(c, 12);
    fold_allreduce(f);
    if (c.lane == 0) {
        double* wd = c.wfd + (size_t)(par * NW + c.warp) * 2;
        int* wi = c.wfi + (size_t)(par * NW + c.warp) * 4;
        wd[0] = f.L; wd[1] = f.H; wi[0] = f.mu; wi[1] = f.m1; wi[2] = f.mcnt; wi[3] = f.nempty;
    }
    __syncthreads();   // #1: folds, holes and their count are in shared memory
    // block fold: lane w of every warp picks up warp w's fold, then the same all-reduce (no second barrier)
    f.init();
    if (c.lane < NW) {
        const double* wd = c.wfd + (size_t)(par * NW + c.lane) * 2;
        const int* wi = c.wfi + (size_t)(par * NW + c.lane) * 4;
        f.L = wd[0]; f.H = wd[1]; f.mu = wi[0]; f.m1 = wi[1]; f.mcnt = wi[2]; f.nempty = wi[3];
    }
    fold_allreduce(f);
    STAMP(c, 13);
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
                int kmax = 2;
                while (kmax < nh) kmax <<= 1;
                bitonic_sort_holes_reg(a, b, c.lane, kmax);
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
    STAMP(c, 14);
    if (c.tid == 0) {
        int e = 0;
        double xv = 0.0;
        const int found = choose_point(p0, q0, r0, c.clo, c.chi, nC, rng, &xv, &e);
        c.dctl[0] = xv;
        c.ictl[IC_FOUND] = found;
        c.ictl[IC_ERR] = e;
    }
    STAMP(c, 15);
    __syncthreads();   // #2
    STAMP(c, 16);
    *xout = c.dctl[0];
    *err = c.ictl[IC_ERR];
    return c.ictl[IC_FOUND];
}

// One phase-1 probe (onevar_qcqp at level s with the zero objective, up to the random draw) by ONE warp over the compacted
// constraint list, into the warp's own hole / piece buffers: returns the number of feasible pieces, ascending in (clo, chi).
// Per probe: the kept constraints' feasible sets (lanes over constraints) -> fold of the singles and hulls (warp all-reduce) ->
// holes into shared memory -> the feasible pieces WITHOUT sorting the holes: with
//     M_i = max(L, max{b_j : a_j < a_i})   and   tie_i = (another hole starts at a_i),
// hole i ends a piece [M_i, a_i] iff !tie_i and M_i < a_i < H -- exactly what the scan over the sorted list reports (for an untied
// hole the holes sorted before it are those with a smaller start; a tied one is never reported, so its M is irrelevant) -- and each
// lane gets the M and the tie count of its own holes from one pass over the list (independent compares instead of the dependent
// compare-exchange stages of a sorting network).  Pieces are then ranked by their right end; the piece ending at H comes last
// (onevar.cuh "HOLE formulation").  Kept out of line so that the CTA-wide kernel keeps its register budget.
// *solid_out = 1: the intersection of the feasible sets at this level contains NO open interval -- a constraint with an empty set, an
// empty box, or the box (L, H) covered by the closed holes: no gap (M_i, a_i) with M_i < a_i < H, ties or not, and a hole over the
// left neighbourhood of H.  That is a statement about sets, free of the reference's reporting quirks (tied starts, coincident right
// ends), and every constraint's feasible set only grows with the level (each operation of get_feasible_intervals is monotone in
// s under round-to-nearest): a solid level certifies that every LOWER level reports no piece either.
__device__ __noinline__ int blk_warp_probe(double2* hx, double* clo, double* chi, const double* ap, const double* aq, const double* ar,
                                           const int* arel, int n_act, int n_inert, double s, int lane, long long* prof, int* solid_out)
{
    const unsigned lt = (1u << lane) - 1u;
#ifdef BLK_PROF
    long long t_last = clock64();
#define BSTAMP(slot) do { if (prof) { const long long now__ = clock64(); prof[slot] += now__ - t_last; t_last = now__; } } while (0)
#define BCOUNT(slot, v) do { if (prof) prof[slot] += (v); } while (0)
#else
#define BCOUNT(slot, v) do { } while (0)
#define BSTAMP(slot) do { } while (0)
#endif
    // ---- feasible sets at level s ----
    Fold f;
    f.init();
    int nh = 0;
    for (int base = 0; base < n_act; base += 32) {
        const int i = base + lane;
        bool hole = false;
        Hole hh;
        hh.a = hh.b = 0.0;
        if (i < n_act) {
            Ival I[2];
            const int c = feasible_intervals(ap[i], aq[i], ar[i], arel[i], s, I);
            hole = fold_constraint(f, c, I, &hh);
        }
        const unsigned hb = __ballot_sync(FULL, hole);
        if (hole) hx[nh + __popc(hb & lt)] = make_double2(hh.a, hh.b);
        nh += __popc(hb);
    }
    BSTAMP(30);
    fold_allreduce(f);
    f.mcnt += n_inert; f.m1 += n_inert;
    if (f.H == QCQP_INF) f.mu += n_inert;
    BSTAMP(31);
    // no feasible point for one constraint, or an empty intersection of the singles and hulls: the total never gets full
    BCOUNT(45, 1); BCOUNT(46, nh);
    *solid_out = 1;
    if (f.nempty > 0 || !(f.L < f.H)) { BCOUNT(47, 1); return 0; }
    __syncwarp();
    // ---- pieces ending at a hole start ----
    // four inert pads behind the list: the walk below reads the list four holes at a time, one group ahead of its use
    if (lane < 4) hx[nh + lane] = make_double2(QCQP_INF, QCQP_INF);
    __syncwarp();
    int nC = 0;
    bool blocked = false, gap = false, cov = false;
    double stH = -QCQP_INF;
    const int nh32 = (nh + 31) & ~31;
    for (int ib = 0; ib < nh32; ib += 32) {
        const int i = ib + lane;
        double a = QCQP_INF, b = QCQP_INF;
        if (i < nh) { const double2 t = hx[i]; a = t.x; b = t.y; }
        // one hole over the whole of (L, H): nothing is left whatever the others do
        if (__any_sync(FULL, i < nh && a <= f.L && f.H <= b)) { BSTAMP(32); BCOUNT(48, 1); return 0; }
        // M = max(L, max{b_j : a_j < a}) for my hole: one pass over the list (pads: a_j = +inf is never smaller)
        double M = f.L;
        {
            double2 n0 = hx[0], n1 = hx[1], n2 = hx[2], n3 = hx[3];
#pragma unroll 1
            for (int j = 0; j < nh; j += 4) {
                const double2 u0 = n0, u1 = n1, u2 = n2, u3 = n3;
                if (j + 4 < nh) { n0 = hx[j + 4]; n1 = hx[j + 5]; n2 = hx[j + 6]; n3 = hx[j + 7]; }
                if (u0.x < a && u0.y > M) M = u0.y;
                if (u1.x < a && u1.y > M) M = u1.y;
                if (u2.x < a && u2.y > M) M = u2.y;
                if (u3.x < a && u3.y > M) M = u3.y;
            }
        }
        const bool valid0 = (i < nh) && (M < a) && (a < f.H);
        if (a <= f.H && f.H <= b) blocked = true;      // pads: a = +inf, never
        if (a < f.H && f.H <= b) cov = true;
        if (b < f.H && b > stH) stH = b;
        const unsigned vb0 = __ballot_sync(FULL, valid0);
        if (vb0) {
            gap = true;
            // a hole whose start another hole shares ends no piece (the reference's dict nets the two events to -2)
            int ties = 0;
            if (nh <= 32) {
                for (unsigned mask = vb0; mask; mask &= mask - 1) {
                    const int src = __ffs(mask) - 1;
                    const double av = __shfl_sync(FULL, a, src);
                    const int t = __popc(__ballot_sync(FULL, a == av));
                    if (lane == src) ties = t;
                }
            } else {
                for (int j = 0; j < nh; j++) ties += (hx[j].x == a) ? 1 : 0;
            }
            const bool valid = valid0 && ties == 1;
            const unsigned vb = __ballot_sync(FULL, valid);
            if (vb && ib == 0) {
                // ascending in a: rank = valid holes of the chunk with a smaller start
                int rank = 0;
                for (unsigned mask = vb; mask; mask &= mask - 1) {
                    const int src = __ffs(mask) - 1;
                    const double av = __shfl_sync(FULL, a, src);
                    rank += (av < a) ? 1 : 0;
                }
                if (valid) { clo[rank] = M; chi[rank] = a; }
                nC = __popc(vb);
                __syncwarp();
            } else if (vb) {
                // later chunks are rare (more than 32 holes): lane 0 inserts their pieces one by one, keeping chi ascending
                for (unsigned mask = vb; mask; mask &= mask - 1) {
                    const int src = __ffs(mask) - 1;
                    const double av = __shfl_sync(FULL, a, src), Mv = __shfl_sync(FULL, M, src);
                    if (lane == 0) {
                        int q = nC;
                        while (q > 0 && chi[q - 1] > av) { clo[q] = clo[q - 1]; chi[q] = chi[q - 1]; q--; }
                        clo[q] = Mv; chi[q] = av;
                    }
                    nC++;
                }
                __syncwarp();
            }
        }
    }
    BSTAMP(32);
    // ---- the piece ending at H ----
    if (f.mu == 1 && f.H < QCQP_INF) {
        blocked = __any_sync(FULL, blocked);
        double st = warp_max(stH);
        if (f.L > st) st = f.L;
        if (!blocked) {
            if (lane == 0) { clo[nC] = st; chi[nC] = f.H; }
            nC++;
        }
    }
    __syncwarp();
    BSTAMP(33);
    BCOUNT(51, nC > 0 ? 1 : 0);
    *solid_out = (!gap && __any_sync(FULL, cov)) ? 1 : 0;
    return nC;
}

// cached f_j(x) from scratch for forms j >= j0 (warps take blocks of 32 forms, same per-form summation as cd.cu's
// refresh_fvals); returns the max constraint violation.  One copy, called from every sweep boundary.
template <int T>
__device__ __noinline__ double blk_refresh_fvals(const PackView& P, const double* x, double* fv, double* redd, int j0, int strict)
{
    constexpr int NW = T / 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    __syncthreads();
    for (int base = j0 + warp * 32; base <= P.m; base += NW * 32) {
        const int hi = (base + 31 < P.m) ? base + 31 : P.m;
        eval_forms(P, x, base, hi, strict != 0, lane, [&](int j, double v) { fv[j] = v; }, false);
    }
    __syncthreads();
    double mv = -QCQP_INF;
    for (int j = 1 + tid; j <= P.m; j += T) {
        const double v = violation_of(P.relop[j], fv[j]);
        mv = (v > mv) ? v : mv;
    }
    mv = warp_max(mv);
    if (lane == 0) redd[warp] = mv;
    __syncthreads();
    double mx = redd[0];
#pragma unroll
    for (int i = 1; i < NW; i++) { const double v = redd[i]; mx = (v > mx) ? v : mx; }
    __syncthreads();
    return mx;
}

template <int T, int MINB>
__global__ void __launch_bounds__(T, MINB) cd_blk_kernel(const __grid_constant__ PackView P, CdK prm, BlkLayout lay,
                                                            const double* __restrict__ X0, int R, qcqp_rng_state* rngs, double* __restrict__ X,
                                                            double* __restrict__ f0_out, double* __restrict__ mv_out, qcqp_cd_stats* stats_out,
                                                            double* ws_fval, double* ws_scr, int* ws_scrj)
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
    Scr sc_s, sc_g;
    sc_s.p = reinterpret_cast<double*>(smem + lay.o_scp);
    sc_s.q = reinterpret_cast<double*>(smem + lay.o_scq);
    sc_s.r = reinterpret_cast<double*>(smem + lay.o_scr);
    sc_s.rj = reinterpret_cast<int*>(smem + lay.o_screl);
    sc_g = sc_s;
    if (P.max_inc > lay.sc_cap) {
        sc_g.p = ws_scr + rr * 3 * (size_t)P.max_inc;
        sc_g.q = sc_g.p + P.max_inc;
        sc_g.r = sc_g.q + P.max_inc;
        sc_g.rj = ws_scrj + rr * (size_t)P.max_inc;
    }
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
    c.par = 0;
    const int tid = c.tid;
    Scratch act;      // phase 1: the constraints of the current coordinate that are not inert over the whole bisection
    act.p = reinterpret_cast<double*>(smem + lay.o_ap);
    act.q = reinterpret_cast<double*>(smem + lay.o_aq);
    act.r = reinterpret_cast<double*>(smem + lay.o_ar);
    act.rel = reinterpret_cast<int*>(smem + lay.o_arel);

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
    const int strict = (prm.mode == MODE_STRICT) ? 1 : 0;
    const double tol = prm.tol, viol_tol = prm.viol_tol;
#ifdef BLK_PROF
    long long prof[64];
    for (int i = 0; i < 64; i++) prof[i] = 0;
    c.prof = prof; c.on = 0; c.last = 0; c.base = 0; prof[57] = clock64();
#endif
    int phase = BPH_P1;
    int t = 0;                    // sweeps done in the current phase
    long long update_counter = 0;
    double viol_last = QCQP_INF;  // phase 1
    double viol_p2 = 0.0;         // phase 2: frozen at entry (qcqp.py:157)
    const bool p1_over = !prm.phase1;
    // sweep-boundary actions are folded into one evaluation site: what = 0 none, 1 phase-1 start (j0 = 1),
    // 2 phase-1 -> phase-2 test (j0 = 0), 3 end of a phase-1 sweep (j0 = 1), 4 periodic phase-2 refresh (j0 = 0), 5 final (j0 = 0)
    int what = 0;
    bool skip = false;            // phase 1 'failed' break: the rest of the sweep is not executed (qcqp.py:138-141)
    long long upd_before = 0;
    double final_mv = 0.0;

    for (;;) {
        // ---------------- sweep boundary (block-uniform) ----------------
        if (what == 0) {
            if (phase == BPH_P1) what = (p1_over || t >= prm.num_iters || viol_last < viol_tol) ? 2 : (t == 0 ? 1 : 0);
            else if (phase == BPH_DONE) what = 5;
        }
        if (what != 0) {
#ifdef BLK_PROF
            const long long trf = clock64();
#endif
            const double mv = blk_refresh_fvals<T>(P, c.x, c.fval, c.redd, (what == 1 || what == 3) ? 1 : 0, strict);
#ifdef BLK_PROF
            prof[58] += clock64() - trf; prof[59]++;
#endif
            const int w = what;
            what = 0;
            if (w == 2) {
                // improve_coord_descent: if max(prob.violations(x)) < viol_tol: phase 2   (qcqp.py:189-190)
                if (m == 0) { st.status = QCQP_RUN_EMPTY_MAX; phase = BPH_DONE; }
                else if (mv < viol_tol) { phase = BPH_P2; viol_p2 = mv; t = 0; update_counter = 0; st.ran_phase2 = 1; }
                else phase = BPH_DONE;
                if (phase == BPH_DONE) { what = 5; continue; }
            } else if (w == 3) {
                viol_last = mv;   // viol = max(prob.violations(x))  (qcqp.py:142)
                t++;
                if (!skip && st.updates_p1 == upd_before && !(viol_last < viol_tol) && t < prm.num_iters) {
                    // a full sweep that moved nothing drew no random number either: every remaining iteration of qcqp.py:110
                    // would repeat it exactly (cd.cu has the argument)
                    st.steps_skipped += (long long)(prm.num_iters - t) * n;
                    t = prm.num_iters;
                }
                continue;
            } else if (w == 5) {
                final_mv = mv;
                break;
            }
        }
        if (phase == BPH_P2 && t >= prm.num_iters) { phase = BPH_DONE; continue; }
        if (phase == BPH_P1) st.sweeps_p1++;
        if (phase == BPH_P2) st.sweeps_p2++;
        skip = false;
        upd_before = st.updates_p1;

        // incidence ranges and the static metadata of this thread's first two slots run one coordinate ahead of the work
        int pb0 = P.inc_ptr[0], pb1 = P.inc_ptr[1], pb2 = P.inc_ptr[n >= 2 ? 2 : 1];
        PMeta pf0 = blk_load_meta(P, pb0 + tid, pb0 + tid < pb1);
        PMeta pf1 = blk_load_meta(P, pb0 + tid + T, pb0 + tid + T < pb1);
        blk_load_row0(P, pf0);
        blk_load_row0(P, pf1);
        for (int k = 0; k < n && phase != BPH_DONE && !skip; k++) {
#ifdef BLK_PROF
            c.on = (phase == BPH_P1); c.base = (pb1 - pb0 <= lay.sc_cap) ? 0 : 20; c.last = clock64();
            if (c.on) prof[60 + (c.base ? 1 : 0)]++;
#endif
            const PMeta cur0 = pf0, cur1 = pf1;
            const int beg = pb0, end = pb1;
            pb0 = pb1; pb1 = pb2;
            if (k + 3 <= n) pb2 = P.inc_ptr[k + 3];
            pf0 = blk_load_meta(P, pb0 + tid, (k + 1 < n) && (pb0 + tid < pb1));
            pf1 = blk_load_meta(P, pb0 + tid + T, (k + 1 < n) && (pb0 + tid + T < pb1));
            STAMP(c, 10);
            const int cnt = end - beg;            // incident forms; the objective, when incident, is slot 0 (thread 0's)
            const Scr sc = (cnt <= lay.sc_cap) ? sc_s : sc_g;
            const double xk = c.x[k];
            const bool in_p1 = (phase == BPH_P1);
            // objective coefficients: thread 0 only (it owns slot 0 and the RNG); phase 1 never looks at the objective
            const bool obj_mine = (tid == 0) && (cnt > 0) && ((cur0.fw & INC_FORM_MASK) == 0);
            double p0 = 0.0, q0 = 0.0, r0 = 0.0;
            // ---- coefficients of my slots (tid, tid + T, ...): private to this thread until the next coordinate ----
            {
                const bool has0 = tid < cnt, has1 = tid + T < cnt;
                const double fv0 = has0 ? c.fval[cur0.fw & INC_FORM_MASK] : 0.0;
                const double fv1 = has1 ? c.fval[cur1.fw & INC_FORM_MASK] : 0.0;
                double pa = 0.0, qa = 0.0, ra = 0.0, pb = 0.0, qb = 0.0, rb = 0.0;
                int rja = 0, rjb = 0;
                if (has0) blk_coeffs_from(P, c, cur0, fv0, xk, pa, qa, ra, rja);
                if (has1) blk_coeffs_from(P, c, cur1, fv1, xk, pb, qb, rb, rjb);
                if (obj_mine) {
                    if (!in_p1) { p0 = pa; q0 = qa; r0 = ra; }
                    pa = 0.0; qa = 0.0; ra = 0.0; rja = 0;
                }
                if (has0) { sc.p[tid] = pa; sc.q[tid] = qa; sc.r[tid] = ra; sc.rj[tid] = rja; }
                if (has1) { sc.p[tid + T] = pb; sc.q[tid + T] = qb; sc.r[tid + T] = rb; sc.rj[tid + T] = rjb; }
            }
            if (tid + 2 * T < cnt) {
                // long lists: metadata two rounds ahead, the cached f_j one round ahead of the arithmetic
                int i = tid + 2 * T;
                PMeta m0 = blk_load_meta(P, beg + i, true);
                PMeta m1 = blk_load_meta(P, beg + i + T, i + T < cnt);
                double fv0 = c.fval[m0.fw & INC_FORM_MASK];
                for (; i < cnt; i += T) {
                    const PMeta m2 = blk_load_meta(P, beg + i + 2 * T, i + 2 * T < cnt);
                    const double fv1 = (i + T < cnt) ? c.fval[m1.fw & INC_FORM_MASK] : 0.0;
                    const int j = (int)(m0.fw & INC_FORM_MASK);
                    double dot = 0.0;
                    for (int u = m0.rbeg; u < m0.rbeg + m0.rlen; u++) dot = dot + P.row_val[u] * c.x[P.row_col[u]];
                    const double t1 = 2 * dot + m0.qk;
                    sc.p[i] = m0.t2; sc.q[i] = t1; sc.r[i] = fv0 - xk * (m0.t2 * xk + t1);
                    sc.rj[i] = (int)((m0.fw >> INC_RELOP_SHIFT) & 3) | (j << 2);
                    m0 = m1; m1 = m2; fv0 = fv1;
                }
            }
            STAMP(c, 11);
            // ---- the probes: phase 1 bisects the violation level (qcqp.py:113-141), phase 2 asks once at the frozen level
            //      (qcqp.py:162-176); one call site, so the solver's code exists once ----
            double new_xi = xk;
            bool move = false;
            bool dead = false;   // the reference would have raised: stop this restart, leave x as it was
            double viol = 0.0, new_viol = 0.0, ss = 0.0, es = 0.0;
            if (in_p1) {
                st.steps_p1++;
                double vmax = -QCQP_INF;
                int cz = 0;
                for (int i = tid; i < cnt; i += T) {
                    const double p = sc.p[i], q = sc.q[i];
                    if (p == 0.0 && q == 0.0) continue;
                    cz++;
                    const double v = violation_of(sc.rj[i] & 3, onevar_eval(p, q, sc.r[i], xk));
                    vmax = (v > vmax) ? v : vmax;
                }
                blk_max_sum<NW>(c, vmax, cz);
                if (cz == 0) { st.status = QCQP_RUN_EMPTY_MAX; dead = true; }
                viol = vmax; new_viol = vmax;
                ss = -tol; es = viol - viol_tol;
                STAMP(c, 1);
            } else {
                st.steps_p2++;
            }
            // ---- phase 1, compaction.  Every probe level of this coordinate satisfies s >= ss0 = -tol.  A concave constraint
            //      (p < -1e-4, not an equality) whose discriminant q*q - (4p)(r - ss0) is negative has a negative discriminant at
            //      every such s too -- r - s, (4p)(r - s) and the difference are each monotone in s under round-to-nearest -- so
            //      its feasible set is the whole line in every probe: it only counts (mcnt, m1).  What is left (circle packing:
            //      the circles whose band |dy| < 2r the centre can actually hit, plus the two box constraints) usually fits one
            //      warp, and warp 0 then runs the whole bisection with the warp-level solver -- no block barrier per probe.
            bool warp_path = false;
            int n_act = 0, n_inert = 0;
            if (in_p1 && !dead && lay.act_cap > 0 && (es - ss > tol) && cnt <= 32 * lay.act_cap) {
                const double ss0 = ss;
                int my_inert = 0;
                for (int base = c.warp * 32; base < cnt; base += T) {
                    const int i = base + c.lane;
                    double p = 0.0, q = 0.0, r = 0.0;
                    int rel = 0;
                    if (i < cnt) { p = sc.p[i]; q = sc.q[i]; r = sc.r[i]; rel = sc.rj[i] & 3; }
                    const bool counted = !(p == 0.0 && q == 0.0);
                    const bool inert = counted && rel != QCQP_RELOP_EQ && p < -IVAL_TOL && (q * q - (4 * p) * (r - ss0) < 0);
                    const bool keep = counted && !inert;
                    my_inert += inert ? 1 : 0;
                    const unsigned kb = __ballot_sync(FULL, keep);
                    if (kb) {
                        int pos0 = 0;
                        if (c.lane == 0) pos0 = atomicAdd((int*)(c.ictl + IC_NACT), __popc(kb));
                        pos0 = __shfl_sync(FULL, pos0, 0);
                        const int pos = pos0 + __popc(kb & ((1u << c.lane) - 1u));
                        if (keep && pos < lay.act_cap) { act.p[pos] = p; act.q[pos] = q; act.r[pos] = r; act.rel[pos] = rel; }
                    }
                }
                my_inert = warp_sum_i(my_inert);
                if (c.lane == 0 && my_inert) atomicAdd((int*)(c.ictl + IC_NINERT), my_inert);
                __syncthreads();
                n_act = c.ictl[IC_NACT]; n_inert = c.ictl[IC_NINERT];
                warp_path = (n_act <= lay.act_cap);
                STAMP(c, 2);
#ifdef BLK_PROF
                if (c.on) { prof[62] += n_act; prof[63] += warp_path ? 1 : 0; }
#endif
            }
            if (warp_path) {
                // The reference probes one level at a time (qcqp.py:122-131): s = (ss + es) / 2, infeasible -> ss = s, feasible -> es = s.
                // While its probes are infeasible ss climbs the chain c_1 = (ss + es) / 2, c_{m+1} = (c_m + es) / 2 (K levels until
                // es - c_K <= tol).  A level that is SOLID (blk_warp_probe: the feasible sets share no open interval) certifies every
                // lower level infeasible, so the loop's next feasible-or-quirky probe is the first non-solid level of the chain, found
                // by a search over the chain index: each round the CTA's warps probe up to NW levels spread over the open index range
                // (always including the deepest one, so a coordinate that cannot move -- 4 of 5 on circle packing -- costs ONE round
                // instead of 14 probes).  The exact result of the first non-solid level (its pieces stay in the buffers of the warp
                // that probed it, which sits out the rest of the search) then plays the reference's step: feasible -> thread 0
                // draws (the reference's stream order) and es drops to it; not feasible (a piece hidden by the reference's tie
                // rules) -> ss rises to it.  The levels are replayed with the reference's own arithmetic, bit for bit.
#ifdef BLK_PROF
                long long* bprof = (c.on && c.warp == 0) ? prof + c.base : nullptr;
#else
                long long* bprof = nullptr;
#endif
                double2* my_hx = reinterpret_cast<double2*>(smem + lay.o_whx) + (size_t)c.warp * (lay.act_cap + 4);
                double* my_clo = reinterpret_cast<double*>(smem + lay.o_wclo) + (size_t)c.warp * (lay.act_cap + 2);
                double* my_chi = reinterpret_cast<double*>(smem + lay.o_wchi) + (size_t)c.warp * (lay.act_cap + 2);
                volatile int* wres = reinterpret_cast<volatile int*>(smem + lay.o_wres);     // [2 parities][result, index][NW]
                double* my_lev = reinterpret_cast<double*>(smem + lay.o_wlev) + (size_t)c.warp * 64;
                int rpar = 0;
                while (es - ss > tol) {
                    // the chain, once per bracket: every warp keeps its own copy of c_0 = ss, c_1 .. c_K (the first 63; beyond, replayed)
                    int K = 0;
                    {
                        double cl = ss;
                        __syncwarp();                       // the previous chain's levels have been read by every lane
                        if (c.lane == 0) my_lev[0] = cl;
                        for (; es - cl > tol && K < 4096; ) { cl = (cl + es) / 2; K++; if (c.lane == 0 && K < 64) my_lev[K] = cl; }
                        __syncwarp();
                    }
                    auto level = [&](int mm) {
                        if (mm < 64) return my_lev[mm];
                        double cl = ss;
                        for (int i = 0; i < mm; i++) cl = (cl + es) / 2;
                        return cl;
                    };
                    int lo = 0, hi = K + 1, hi_nC = 0, keep = -1;
                    while (hi - lo > 1) {
                        const int span = hi - lo - 1;
                        const int nev = (keep < 0) ? NW : NW - 1;
                        const int slot = (keep < 0) ? c.warp : ((c.warp == keep) ? -1 : (c.warp < keep ? c.warp : c.warp - 1));
                        int mm = -1;
                        if (slot >= 0) {
                            if (span <= nev) mm = (slot < span) ? lo + 1 + slot : -1;
                            else mm = (slot == 0 || nev == 1) ? lo + 1      // the lowest open level (a coordinate that improves is usually feasible there) ...
                                                                : lo + 1 + (int)(((long long)slot * (span - 1) + nev - 2) / (nev - 1));   // ... the rest spread up to hi - 1
                        }
                        int r = -1;
                        if (mm > 0) {
                            int solid = 0;
                            const int nC = blk_warp_probe(my_hx, my_clo, my_chi, act.p, act.q, act.r, act.rel, n_act, n_inert, level(mm), c.lane, bprof, &solid);
                            r = nC | (solid << 30);
                        }
                        if (c.lane == 0) { wres[rpar * 2 * NW + c.warp] = r; wres[rpar * 2 * NW + NW + c.warp] = mm; }
                        __syncthreads();
                        for (int w = 0; w < NW; w++) {
                            const int rw = wres[rpar * 2 * NW + w], mw = wres[rpar * 2 * NW + NW + w];
                            if (mw <= 0 || rw < 0) continue;
                            if ((rw >> 30) & 1) { if (mw > lo) lo = mw; }
                            else if (mw < hi) { hi = mw; hi_nC = rw & 0x3fffffff; keep = w; }
                        }
                        // a solid level above a non-solid one: the certificate wins (everything up to lo is infeasible)
                        if (hi <= lo) { hi = K + 1; hi_nC = 0; keep = -1; }
                        rpar ^= 1;
                    }
                    const double s_lo = level(lo), s_hi = (hi <= K) ? level(hi) : 0.0;
                    if (lo >= 1) ss = s_lo;
                    int e = 0;
                    if (hi <= K) {
                        if (hi_nC > 0) {
                            if (tid == 0) {
                                // np.random.uniform(*C[np.random.choice(len(C))])  (utilities.py:266-267)
                                const double* wl = reinterpret_cast<double*>(smem + lay.o_wclo) + (size_t)keep * (lay.act_cap + 2);
                                const double* wh = reinterpret_cast<double*>(smem + lay.o_wchi) + (size_t)keep * (lay.act_cap + 2);
                                const int idx = rng.choice(hi_nC);
                                const double plo = wl[idx], phi = wh[idx];
                                int ee = 0;
                                double xv = 0.0;
                                if (is_inf(plo) || is_inf(phi)) ee = QCQP_RUN_UNBOUNDED_UNIFORM;
                                else xv = rng.uniform(plo, phi);
                                c.dctl[0] = xv; c.ictl[IC_ERR] = ee;
                            }
                            __syncthreads();
                            e = c.ictl[IC_ERR];
                            new_xi = c.dctl[0]; new_viol = s_hi; es = s_hi;
                            __syncthreads();      // dctl / the piece buffers are free again
                        } else {
                            ss = s_hi;
                        }
                    }
                    if (e) { st.status = e; dead = true; break; }
                }
                STAMP(c, 3);
            }
            bool asked = false;
            while (!dead && !warp_path) {
                double s;
                if (in_p1) {
                    if (!(es - ss > tol)) break;
                    s = (ss + es) / 2;
                } else {
                    if (asked) break;
                    s = viol_p2;
                    asked = true;
                }
                double xi;
                int err;
                const int ok = blk_solve_level<T>(c, sc, cnt, s, p0, q0, r0, rng, &xi, &err);
                if (err) { st.status = err; dead = true; break; }
                if (in_p1) {
                    if (!ok) ss = s;
                    else { new_xi = xi; new_viol = s; es = s; }
                } else if (ok && fabs(xi - xk) > tol) {
                    move = true; new_xi = xi;
                }
            }
            if (!dead) {
                if (in_p1) {
                    if (new_viol < viol) { move = true; update_counter = 0; st.updates_p1++; }
                    else {
                        update_counter++;
                        if (update_counter == n) skip = true;
                    }
                } else if (move) {
                    update_counter = 0; st.updates_p2++;
                } else {
                    update_counter++;
                    if (update_counter == n) phase = BPH_DONE;   // converged
                }
            }
            STAMP(c, 17);
            if (move) {
                // f_j(x) = t0 + b (t2 b + t1) for every incident constraint (and, in phase 2, the objective)
                const double b = new_xi;
                for (int i = tid; i < cnt; i += T) {
                    const int j = sc.rj[i] >> 2;
                    if (j > 0) c.fval[j] = sc.r[i] + b * (sc.p[i] * b + sc.q[i]);
                }
                if (tid == 0) {
                    if (!in_p1 && obj_mine) c.fval[0] = r0 + b * (p0 * b + q0);
                    c.x[k] = new_xi;
                }
            }
            if (tid == 0 && in_p1) { c.ictl[IC_NACT] = 0; c.ictl[IC_NINERT] = 0; }
            if (dead) phase = BPH_DONE;
            blk_load_row0(P, pf0);
            blk_load_row0(P, pf1);
            STAMP(c, 18);
            __syncthreads();   // x and the cached f_j are current for the next coordinate
            STAMP(c, 19);
        }
        // ---------------- end of sweep ----------------
        if (phase == BPH_P1) what = 3;
        else if (phase == BPH_P2) {
            t++;
            if (prm.refresh_every > 0 && (t % prm.refresh_every) == 0) what = 4;
        }
    }

    // ---------------- results: x, (f0.eval(x), max(violations(x))) as QCQP._improve returns them (qcqp.py:415-417) -------
    for (int i = tid; i < n; i += T) X[rr * n + i] = c.x[i];
    for (int i = tid; i < 624; i += T) rngs[rr].key[i] = mt[i];
    if (tid == 0) {
        rngs[rr].pos = rng.pos;
        f0_out[rr] = c.fval[0];
        mv_out[rr] = (m > 0) ? final_mv : 0.0;
        if (stats_out) stats_out[rr] = st;
#ifdef BLK_PROF
        if (rr == 0) {
            printf("centre steps %lld radius steps %lld; kept constraints / centre step %.1f; warp path in %lld steps\n", prof[60], prof[61],
                   (double)prof[62] / (double)(prof[60] > 0 ? prof[60] : 1), prof[63]);
            printf("warp-0 probes %lld (%.1f / centre step): holes / probe %.1f; empty box %lld, one hole covers %lld, chunks without a start inside the box %lld, "
                   "starts inside the box / chunk %.2f, feasible %lld\n", prof[45], (double)prof[45] / (double)(prof[60] > 0 ? prof[60] : 1),
                   (double)prof[46] / (double)(prof[45] > 0 ? prof[45] : 1), prof[47], prof[48], prof[49],
                   (double)prof[50] / (double)(prof[45] - prof[47] - prof[48] > 0 ? prof[45] - prof[47] - prof[48] : 1), prof[51]);
            printf("refresh of the cached f_j: %.0f cycles each, %lld calls; whole kernel %lld cycles\n", (double)prof[58] / (double)(prof[59] > 0 ? prof[59] : 1), prof[59], clock64() - prof[57]);
            for (int i = 30; i < 35; i++) printf("bisect slot %d: %8.0f cycles / centre step\n", i, (double)prof[i] / (double)(prof[60] > 0 ? prof[60] : 1));
            for (int i = 0; i < 20; i++)
                if (prof[i] || prof[20 + i]) printf("slot %2d: centre %8.0f cycles / step   radius %10.0f cycles / step\n", i, (double)prof[i] / (double)(prof[60] > 0 ? prof[60] : 1),
                                                    (double)prof[20 + i] / (double)(prof[61] > 0 ? prof[61] : 1));
        }
#endif
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

template <int T, int MINB>
static int blk_launch_t(qcqp_pack* p, const CdK& k, const BlkLayout& L, const double* dX0, int R, qcqp_rng_state* drng, double* dX, double* df0,
                        double* dmv, qcqp_cd_stats* dstats, double* ws_fval, double* ws_scr, int* ws_scrj, cudaStream_t stream)
{
    QCQP_CUDA_TRY(cudaFuncSetAttribute(cd_blk_kernel<T, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    cd_blk_kernel<T, MINB><<<R, T, L.total, stream>>>(p->v, k, L, dX0, R, drng, dX, df0, dmv, dstats, ws_fval, ws_scr, ws_scrj);
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
    if (force_threads == 64 || force_threads == 128 || force_threads == 256 || force_threads == 512) T = force_threads;
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
    // hole / piece buffers: the CTA-wide solver's (hcap holes) and, over the same bytes, the per-warp ones of the speculative phase-1
    // probes (act_cap holes per warp) -- a coordinate step uses one or the other
    l.act_cap = 128;
    { const char* e = getenv("QCQP_BLK_WARP"); if (e && atoi(e) == 0) l.act_cap = 0; }   // A/B: every probe by the whole CTA
    {
        const unsigned o0 = o;
        l.o_hx = o; o += (unsigned)hcap * 16;
        l.o_clo = o; o += blk_align((unsigned)(hcap + 2) * 8, 16);
        l.o_chi = o; o += blk_align((unsigned)(hcap + 2) * 8, 16);
        unsigned w = o0;
        l.o_whx = w; w += (unsigned)(NW * (l.act_cap + 4)) * 16;   // + the four pads of blk_warp_probe
        l.o_wclo = w; w += (unsigned)(NW * (l.act_cap + 2)) * 8;
        l.o_wchi = w; w += (unsigned)(NW * (l.act_cap + 2)) * 8;
        if (w > o) o = w;
    }
    l.o_wfd = o; o += (unsigned)(2 * NW * 2) * 8;
    l.o_wfi = o; o += (unsigned)(2 * NW * 4) * 4;
    l.o_cmax = o; o += blk_align((unsigned)(hcap / 32) * 8, 16);
    l.o_ccnt = o; o += blk_align((unsigned)(hcap / 32) * 4, 16);
    l.o_redd = o; o += (unsigned)NW * 8;
    l.o_redi = o; o += blk_align((unsigned)NW * 4, 16);
    l.o_ictl = o; o += 32;
    l.o_dctl = o; o += 16;
    l.o_ap = o; o += (unsigned)l.act_cap * 8;
    l.o_aq = o; o += (unsigned)l.act_cap * 8;
    l.o_ar = o; o += (unsigned)l.act_cap * 8;
    l.o_arel = o; o += blk_align((unsigned)l.act_cap * 4, 16);
    l.o_wres = o; o += blk_align((unsigned)(4 * NW) * 4, 16);
    l.o_wlev = o; o += (unsigned)(NW * 64) * 8;                  // per warp: the levels of the current chain
    l.total = blk_align(o, 128);
    if (l.total > (unsigned)max_smem_optin(p->device))
        return fail(QCQP_ERR_CAPACITY, "qcqp_cd_improve: per-restart state exceeds shared memory (too many two-interval constraints on one coordinate)");
    const size_t fval_bytes = l.fval_smem ? 0 : (size_t)R * (v.m + 1) * 8;
    const bool spill = v.max_inc > l.sc_cap;
    const size_t scr_bytes = spill ? (size_t)R * 3 * v.max_inc * 8 : 0;
    const size_t scrj_bytes = spill ? (size_t)R * v.max_inc * 4 : 0;
    const size_t a1 = (fval_bytes + 255) & ~(size_t)255, a2 = (scr_bytes + 255) & ~(size_t)255;
    int rc = ensure_workspace(p, a1 + a2 + scrj_bytes + 256);
    if (rc != QCQP_OK) return rc;
    double* ws_fval = (double*)p->ws;
    double* ws_scr = (double*)((char*)p->ws + a1);
    int* ws_scrj = (int*)((char*)p->ws + a1 + a2);
    // more restarts than 4 CTAs per SM can hold: the 64-register build of the 128-thread kernel, 8 CTAs per SM -- only when every probe
    // is CTA-wide (QCQP_BLK_WARP=0): with the speculative phase-1 probes all four warps are busy and the 128-register build wins
    // (C5, 4096 restarts: 2.5 s against 4.0 s)
    const char* f8 = getenv("QCQP_BLK_CTAS");
    const bool dense8 = (T == 128) && (f8 ? atoi(f8) == 8 : (R > 4 * sms && l.act_cap == 0));
    if (T == 64) return blk_launch_t<64, 8>(p, k, l, dX0, R, drng, dX, df0, dmv, dstats, ws_fval, ws_scr, ws_scrj, stream);
    if (dense8) return blk_launch_t<128, 8>(p, k, l, dX0, R, drng, dX, df0, dmv, dstats, ws_fval, ws_scr, ws_scrj, stream);
    if (T == 512) return blk_launch_t<512, 1>(p, k, l, dX0, R, drng, dX, df0, dmv, dstats, ws_fval, ws_scr, ws_scrj, stream);
    if (T == 256) return blk_launch_t<256, 2>(p, k, l, dX0, R, drng, dX, df0, dmv, dstats, ws_fval, ws_scr, ws_scrj, stream);
    return blk_launch_t<128, 4>(p, k, l, dX0, R, drng, dX, df0, dmv, dstats, ws_fval, ws_scr, ws_scrj, stream);
}

}  // namespace qcqp
