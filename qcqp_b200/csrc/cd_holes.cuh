// cd_holes.cuh -- warp-level pieces of the GPU sweep line shared by the coordinate-descent kernels (cd.cu: one warp per
// restart; cd_blk.cu: one CTA per restart): sorting the holes of the two-interval constraints, the prefix-max scan that
// turns sorted holes into feasible pieces, and the warp all-reduce of the single-interval fold (onevar.cuh has the argument).
#pragma once
#include "common.cuh"
#include "onevar.cuh"

namespace qcqp {

// ---------------------------------------------------------------------------------------------------------
// Sweep line, GPU formulation (onevar.cuh "HOLE formulation" has the argument and the scalar statement that the CPU
// tests hold against the reference): singles and hulls are folded into (L, H, mu); only the holes of two-interval
// constraints are sorted -- by their start, 16-byte (a, b) records -- and one prefix-max scan of their ends yields
// the feasible pieces in ascending order.  <= 32 holes: registers + shuffles; more: shared memory.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bitonic_sort_holes_smem(double2* h, int N, int lane)
{
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll 4
            for (int t = lane; t < (N >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const double2 a = h[i], b = h[p];
                const bool up = ((i & k) == 0);
                if ((a.x > b.x) == up && a.x != b.x) { h[i] = b; h[p] = a; }
            }
            __syncwarp();
        }
    }
}
// holes in lanes 0..kmax-1 (kmax a power of two <= 32), +inf pads elsewhere: the sub-network of the first kmax lanes is enough
__device__ __forceinline__ void bitonic_sort_holes_reg(double& a, double& b, int lane, int kmax = 32)
{
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
        if (k > kmax) break;
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const double pa = __shfl_xor_sync(FULL, a, j), pb = __shfl_xor_sync(FULL, b, j);
            const bool keepmin = (((lane & j) == 0) == ((lane & k) == 0));
            const bool swap = keepmin ? (pa < a) : (pa > a);
            if (swap) { a = pa; b = pb; }
        }
    }
}

struct HoleScan {
    double carryM;    // max(L, ends of the holes of earlier chunks)          (warp-uniform)
    double prevA;     // start of the last hole of the previous chunk         (warp-uniform)
    bool havePrev;
    bool blocked;     // per lane: one of my holes has a <= H <= b
    double stH;       // per lane: max{b : b < H} over my holes
    int nC;           // pieces written so far                                (warp-uniform)
};

// one chunk of 32 holes in ascending order of a (lane = position); nextA = start of the first hole of the next chunk
__device__ __forceinline__ void scan_hole_chunk(double* clo, double* chi, const Fold& f, double a, double b, double nextA, HoleScan& hs, int lane)
{
    double inc = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o && t > inc) inc = t;
    }
    const double exc = __shfl_up_sync(FULL, inc, 1);
    double M = hs.carryM;
    if (lane > 0 && exc > M) M = exc;
    const double ap = __shfl_up_sync(FULL, a, 1);
    const bool tiep = (lane > 0) ? (ap == a) : (hs.havePrev && hs.prevA == a);
    double an = __shfl_down_sync(FULL, a, 1);
    if (lane == 31) an = nextA;
    const bool valid = !tiep && (an != a) && (M < a) && (a < f.H);
    const unsigned vb = __ballot_sync(FULL, valid);
    if (valid) {
        const int pos = hs.nC + __popc(vb & ((1u << lane) - 1u));
        clo[pos] = M; chi[pos] = a;
    }
    hs.nC += __popc(vb);
    if (a <= f.H && f.H <= b) hs.blocked = true;
    if (b < f.H && b > hs.stH) hs.stH = b;
    const double tot = __shfl_sync(FULL, inc, 31);
    if (tot > hs.carryM) hs.carryM = tot;
    hs.prevA = __shfl_sync(FULL, a, 31);
    hs.havePrev = true;
}

// order-preserving map double -> uint64 (never fed NaN: the fold's L and H only ever take values that won a comparison).
// -0.0 is mapped onto +0.0 first: the reference's sweep keys its events in a dict, where 0.0 and -0.0 are ONE key (and the
// lane-by-lane Fold::merge compares with ==), so two right ends +0.0 and -0.0 in different lanes must share a key here too --
// otherwise their multiplicities would not add up and a piece ending at 0 that the reference hides would be reported.
__device__ __forceinline__ unsigned long long dkey(double v)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v + 0.0);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k)
{
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
// warp max / min of a double with two REDUX instructions each (high word, then the low words of the lanes that hold it)
__device__ __forceinline__ unsigned long long warp_max_key(unsigned long long k)
{
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned mh = __reduce_max_sync(FULL, hi);
    const unsigned ml = __reduce_max_sync(FULL, (hi == mh) ? lo : 0u);
    return ((unsigned long long)mh << 32) | ml;
}
__device__ __forceinline__ unsigned long long warp_min_key(unsigned long long k)
{
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned mh = __reduce_min_sync(FULL, hi);
    const unsigned ml = __reduce_min_sync(FULL, (hi == mh) ? lo : 0xffffffffu);
    return ((unsigned long long)mh << 32) | ml;
}

// all-reduce of the per-lane folds.  Same result as merging lane by lane (Fold::merge): L = max, H = min, mu = the
// multiplicities of the lanes that hold H (lanes that folded nothing carry H = +inf with mu = 0), counts add up.
// (-0.0 and +0.0 compare equal in the lane-by-lane merge; dkey maps both to one key.)
__device__ __forceinline__ void fold_allreduce(Fold& f)
{
    const unsigned long long kh = dkey(f.H);
    const unsigned long long kH = warp_min_key(kh);
    f.L = dunkey(warp_max_key(dkey(f.L)));
    f.mu = __reduce_add_sync(FULL, (kh == kH) ? f.mu : 0);
    f.H = dunkey(kH);
    f.m1 = __reduce_add_sync(FULL, f.m1);
    f.mcnt = __reduce_add_sync(FULL, f.mcnt);
    f.nempty = __reduce_add_sync(FULL, f.nempty);
}
__device__ __forceinline__ void fold_bcast(Fold& f, int src)
{
    f.L = bcast(f.L, src); f.H = bcast(f.H, src); f.mu = bcast_i(f.mu, src); f.m1 = bcast_i(f.m1, src);
    f.mcnt = bcast_i(f.mcnt, src); f.nempty = bcast_i(f.nempty, src);
}

// per-restart shared-memory views of the warp-level solver (cd.cu: one warp per restart; cd_blk.cu: warp 0 of the restart's CTA)
struct WarpMem {
    double* x;
    double* fval;
    uint32_t* mt;
    double2* hx;      // holes (a, b) of the two-interval constraints of this step
    double* clo; double* chi;
    double* dd;       // dots of the dense rows of this step, by dense slot
    double* g;        // gradient mode: g[d][npad] = P_d x for every dense slot
    int* misc;
};

// tail of onevar_qcqp: f = the all-reduced fold (singles + hulls), nh holes -- in (ra, rb), one per lane, when
// in_regs, else in w.hx[0..nh).  Lane 0 owns the RNG.  Returns found (warp-uniform); xout/err valid in every lane.
__device__ __forceinline__ int holes_finish(const WarpMem& w, const Fold& f, int nh, bool in_regs, double ra, double rb, double p0, double q0,
                                            double r0, MtRng& rng, int lane, double* xout, int* err)
{
    *xout = 0.0;
    *err = 0;
    int nC = 0;
    if (f.mcnt == 0) {
        if (lane == 0) { w.clo[0] = -QCQP_INF; w.chi[0] = QCQP_INF; }   // only the sentinel pair: all of R
        nC = 1;
    } else {
        if (f.nempty > 0 || !(f.L < f.H)) return 0;   // the running total never gets full
        HoleScan hs;
        hs.carryM = f.L; hs.prevA = 0.0; hs.havePrev = false; hs.blocked = false; hs.stH = -QCQP_INF; hs.nC = 0;
        if (nh > 0) {
            if (in_regs || nh <= 32) {
                double a = ra, b = rb;
                if (!in_regs) {
                    a = QCQP_INF; b = QCQP_INF;
                    if (lane < nh) { const double2 t = w.hx[lane]; a = t.x; b = t.y; }
                }
                bitonic_sort_holes_reg(a, b, lane);
                scan_hole_chunk(w.clo, w.chi, f, a, b, QCQP_INF, hs, lane);
            } else {
                int N2 = 64;
                while (N2 < nh) N2 <<= 1;
                for (int i = nh + lane; i < N2; i += 32) w.hx[i] = make_double2(QCQP_INF, QCQP_INF);   // pads are inert
                __syncwarp();
                bitonic_sort_holes_smem(w.hx, N2, lane);
                const int nchunk = (nh + 31) >> 5;
                double2 cur = w.hx[lane];
                for (int c = 0; c < nchunk; c++) {
                    const double2 nxt = w.hx[(c + 1 < (N2 >> 5)) ? ((c + 1) << 5) + lane : lane];
                    const double nextA = (c + 1 < (N2 >> 5)) ? __shfl_sync(FULL, nxt.x, 0) : QCQP_INF;
                    scan_hole_chunk(w.clo, w.chi, f, cur.x, cur.y, nextA, hs, lane);
                    cur = nxt;
                }
            }
        }
        nC = hs.nC;
        if (f.mu == 1 && f.H < QCQP_INF) {
            const bool blocked = __any_sync(FULL, hs.blocked);
            double st = warp_max(hs.stH);
            if (f.L > st) st = f.L;
            if (!blocked) {
                if (lane == 0) { w.clo[nC] = st; w.chi[nC] = f.H; }
                nC++;
            }
        }
    }
    __syncwarp();
    int found = 0, e = 0;
    double xv = 0.0;
    if (lane == 0) found = choose_point(p0, q0, r0, w.clo, w.chi, nC, rng, &xv, &e);
    __syncwarp();
    *xout = bcast(xv, 0);
    *err = bcast_i(e, 0);
    return bcast_i(found, 0);
}

// ---------------------------------------------------------------------------------------------------------
// onevar_qcqp, general path: the mk constraint coefficients are in scratch (any mk).
// ---------------------------------------------------------------------------------------------------------
struct Scratch {
    double* p; double* q; double* r; int* rel;
};

// inert: constraints the caller has proven to contribute the whole line at this level (cd_blk.cu's phase-1 compaction): they count
// in mcnt / m1, and in the multiplicity of H only while H is still +inf
__device__ __forceinline__ int solve_level(const WarpMem& w, const Scratch& sc, int mk, double s, double p0, double q0, double r0, MtRng& rng,
                                           int lane, double* xout, int* err, int inert = 0)
{
    *err = 0;
    *xout = 0.0;
    Fold f;
    f.init();
    int nh = 0;
    const unsigned lt = (1u << lane) - 1u;
    for (int base = 0; base < mk; base += 32) {
        const int i = base + lane;
        bool hole = false;
        Hole hh;
        hh.a = hh.b = 0.0;
        if (i < mk) {
            const double p = sc.p[i], q = sc.q[i];
            if (!(p == 0.0 && q == 0.0)) {   // nfs filter of qcqp.py:116,166
                Ival I[2];
                const int c = feasible_intervals(p, q, sc.r[i], sc.rel[i], s, I);
                hole = fold_constraint(f, c, I, &hh);
            }
        }
        // some constraint has no feasible point at this level: the total never reaches m + 1
        if (__any_sync(FULL, f.nempty > 0)) return 0;
        const unsigned hb = __ballot_sync(FULL, hole);
        if (hb) {
            if (hole) w.hx[nh + __popc(hb & lt)] = make_double2(hh.a, hh.b);
            nh += __popc(hb);
        }
        // long lists (circle packing's r: one interval from each of 20 701 constraints): once the partial intersection of the
        // singles is empty the final one is too
        if ((base & 511) == 480 && base + 32 < mk) {
            const double Lp = warp_max(f.L), Hp = -warp_max(-f.H);
            if (!(Lp < Hp)) return 0;
        }
    }
    fold_allreduce(f);
    if (inert > 0) { f.mcnt += inert; f.m1 += inert; if (f.H == QCQP_INF) f.mu += inert; }
    __syncwarp();
    return holes_finish(w, f, nh, false, 0.0, 0.0, p0, q0, r0, rng, lane, xout, err);
}

}  // namespace qcqp
