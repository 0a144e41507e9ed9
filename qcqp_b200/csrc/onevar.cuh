// onevar.cuh -- the 1-D machinery of the coordinate-descent sweep, written for one lane of a warp:
//   * MT19937 + NumPy's legacy uniform/choice transforms (the reference consumes np.random, SURVEY a-7)
//   * feasible set of one scalar quadratic constraint            (get_feasible_intervals, utilities.py:198-232)
//   * the sweep-line intersection with the reference's quirks     (onevar_qcqp, utilities.py:241-261)
//   * the choice of the minimiser over the feasible set           (onevar_qcqp, utilities.py:263-288)
//
// GPU formulation (not the reference's): constraints whose feasible set is ONE interval are folded into a running
// (L = max lo, H = min hi, multiplicity of H, count); only two-interval constraints contribute explicit events.
// The fold is exact with respect to the reference's dict/sorted sweep, including its quirks -- a feasible piece is
// reported only where the running total drops to m by exactly -1, so two coincident right ends hide it, a piece
// unbounded to the right is never seen, zero-width pieces cancel (DESIGN.md "1-D solver" has the argument).
//
// Everything here is __host__ __device__ so that tests/ can run the same code on the CPU build box through
// csrc/host_shim.cpp (unit tests of device code; the product never runs it on the host).
// Compile with -fmad=false: NumPy rounds every multiply and add separately and the branch thresholds depend on it.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define QCQP_HD __host__ __device__ __forceinline__
#else
#define QCQP_HD inline
#endif

namespace qcqp {

#ifndef QCQP_INF
#define QCQP_INF (__builtin_huge_val())
#endif

// ---------------------------------------------------------------------------------------------------------
// MT19937, sequential (one lane owns the stream)
// ---------------------------------------------------------------------------------------------------------
struct MtRng {
    uint32_t* key;  // [624]
    int pos;

    QCQP_HD void refill()
    {
        uint32_t* mt = key;
        int i;
        for (i = 0; i < 624 - 397; i++) {
            uint32_t y = (mt[i] & 0x80000000u) | (mt[i + 1] & 0x7fffffffu);
            mt[i] = mt[i + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        for (; i < 623; i++) {
            uint32_t y = (mt[i] & 0x80000000u) | (mt[i + 1] & 0x7fffffffu);
            mt[i] = mt[i - 227] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        pos = 0;
    }
    QCQP_HD uint32_t next()
    {
        if (pos >= 624) refill();
        uint32_t y = key[pos++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    // random_sample(): 53-bit double from two draws
    QCQP_HD double next_double()
    {
        uint32_t a = next() >> 5, b = next() >> 6;
        return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
    }
    // np.random.uniform(lo, hi)
    QCQP_HD double uniform(double lo, double hi)
    {
        double range = hi - lo;
        return lo + range * next_double();
    }
    // np.random.choice(cnt) == legacy randint(0, cnt): masked rejection; cnt == 1 draws nothing
    QCQP_HD int choice(int cnt)
    {
        uint32_t top = (uint32_t)(cnt - 1);
        if (top == 0) return 0;
        uint32_t mask = top;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t v;
        do { v = next() & mask; } while (v > top);
        return (int)v;
    }
};

// ---------------------------------------------------------------------------------------------------------
// scalar quadratic p x^2 + q x + r
// ---------------------------------------------------------------------------------------------------------
QCQP_HD bool is_inf(double x) { return x == QCQP_INF || x == -QCQP_INF; }

// OneVarQuadraticFunction.eval (utilities.py:115-120)
QCQP_HD double onevar_eval(double p, double q, double r, double x)
{
    if (is_inf(x)) {
        if (p != 0.0) return p * x * x;
        if (q != 0.0) return q * x;
        return r;
    }
    return x * (p * x + q) + r;
}

// QuadraticFunction.violation (utilities.py:56-62)
QCQP_HD double violation_of(int relop, double v)
{
    if (relop == QCQP_RELOP_EQ) return fabs(v);
    return (v > 0.0) ? v : 0.0;
}

struct Ival { double lo, hi; };

constexpr double IVAL_TOL = 1e-4;  // the default tol of get_feasible_intervals; the reference never overrides it

// {x : p x^2 + q x + rr <= 0}; the '<=' branch of utilities.py:210-231 with s already folded into rr
QCQP_HD int ivals_le0(double p, double q, double rr, Ival* out)
{
    // q == 0 (x_k^2 = c constraints: Boolean LS, MAXCUT): -q - rD and -q + rD are exactly -rD and +rD, and IEEE division is
    // sign-symmetric, so the two roots are one quotient and its negation -- same bits, one division less
    if (p > IVAL_TOL) {
        double D = q * q - (4 * p) * rr;
        if (D >= 0) {
            double rD = sqrt(D), den = 2 * p;
            if (q == 0.0) { const double h = rD / den; out[0].lo = -h; out[0].hi = h; return 1; }
            out[0].lo = (-q - rD) / den;
            out[0].hi = (-q + rD) / den;
            return 1;
        }
        return 0;
    }
    if (p < -IVAL_TOL) {
        double D = q * q - (4 * p) * rr;
        if (D >= 0) {
            double rD = sqrt(D), den = 2 * p;
            if (q == 0.0) {
                const double h = rD / den;
                out[0].lo = -QCQP_INF; out[0].hi = h;
                out[1].lo = -h; out[1].hi = QCQP_INF;
                return 2;
            }
            out[0].lo = -QCQP_INF; out[0].hi = (-q + rD) / den;
            out[1].lo = (-q - rD) / den; out[1].hi = QCQP_INF;
            return 2;
        }
        out[0].lo = -QCQP_INF; out[0].hi = QCQP_INF;
        return 1;
    }
    if (q > IVAL_TOL) { out[0].lo = -QCQP_INF; out[0].hi = (0.0 - rr) / q; return 1; }
    if (q < -IVAL_TOL) { out[0].lo = (0.0 - rr) / q; out[0].hi = QCQP_INF; return 1; }
    out[0].lo = -QCQP_INF; out[0].hi = QCQP_INF;
    return 1;
}

// get_feasible_intervals(f, s) (utilities.py:198-232); at most two intervals come out
QCQP_HD int feasible_intervals(double p, double q, double r, int relop, double s, Ival* out)
{
    if (relop != QCQP_RELOP_EQ) {
        // (r - s) folded: the reference computes q*q - 4*p*(r-s) and (s-r)/q.  (s-r) == -(r-s) exactly in IEEE, and
        // 0.0 - (r-s) reproduces it including the sign of zero.
        return ivals_le0(p, q, r - s, out);
    }
    Ival a[2], b[2];
    a[0].lo = a[0].hi = a[1].lo = a[1].hi = b[0].lo = b[0].hi = b[1].lo = b[1].hi = 0.0;
    const int na = ivals_le0(p, q, r - s, a);       // f1 = (p, q, r - s) <= 0
    const int nb = ivals_le0(-p, -q, -r - s, b);    // f2 = (-p, -q, -r - s) <= 0
    // the reference's double loop over (i, k) in the order (0,0), (0,1), (1,0), (1,1), keeping the first two non-empty
    // intersections -- written with static indices only, so that nothing here lives in local memory on the device
    Ival t[4];
    bool v[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const int i = e >> 1, k = e & 1;
        t[e].lo = (b[k].lo > a[i].lo) ? b[k].lo : a[i].lo;
        t[e].hi = (b[k].hi < a[i].hi) ? b[k].hi : a[i].hi;
        v[e] = (i < na) && (k < nb) && (t[e].lo <= t[e].hi);
    }
    const int c = (int)v[0] + (int)v[1] + (int)v[2] + (int)v[3];
    // first valid entry, then the first valid one after it
    const Ival f1 = v[0] ? t[0] : (v[1] ? t[1] : (v[2] ? t[2] : t[3]));
    const Ival s3 = t[3];
    const Ival s2 = v[2] ? t[2] : s3;                 // first valid among {2, 3}
    const Ival s1 = v[1] ? t[1] : s2;                 // first valid among {1, 2, 3}
    const Ival f2 = v[0] ? s1 : (v[1] ? s2 : s3);      // first valid after the first valid
    if (c >= 1) out[0] = f1;
    if (c >= 2) out[1] = f2;
    return c > 2 ? 2 : c;
}

// ---------------------------------------------------------------------------------------------------------
// running fold of the single-interval constraints
// ---------------------------------------------------------------------------------------------------------
struct Fold {
    double L, H;    // max of the left ends, min of the right ends
    int mu;         // how many right ends equal H exactly
    int m1;         // single-interval constraints folded
    int mcnt;       // constraints counted (those with (p, q) != (0, 0): qcqp.py:116,166)
    int nempty;     // constraints with an empty feasible set

    QCQP_HD void init() { L = -QCQP_INF; H = QCQP_INF; mu = 0; m1 = 0; mcnt = 0; nempty = 0; }
    QCQP_HD void add_single(double lo, double hi)
    {
        m1++;
        if (lo > L) L = lo;
        if (hi < H) { H = hi; mu = 1; }
        else if (hi == H) mu++;
    }
    QCQP_HD void merge(double L2, double H2, int mu2, int m12, int mcnt2, int nempty2)
    {
        if (L2 > L) L = L2;
        if (m12 > 0) {
            if (m1 == 0 || H2 < H) { H = H2; mu = mu2; }
            else if (H2 == H) mu += mu2;
        }
        m1 += m12; mcnt += mcnt2; nempty += nempty2;
    }
};

// ---------------------------------------------------------------------------------------------------------
// sweep line over the explicit events (utilities.py:245-261).
// ev_key/ev_del hold nev events already SORTED by key (ties in any order).  mcnt = len(fs).
// Feasible pieces are appended to (c_lo, c_hi); returns their number.
// ---------------------------------------------------------------------------------------------------------
QCQP_HD int sweep_sorted(const double* ev_key, const int* ev_del, int nev, int mcnt, double* c_lo, double* c_hi)
{
    int nC = 0;
    long tot = 0;
    double prev_key = 0.0;   // key of the previous entry with a nonzero net count
    bool have_prev = false;
    int i = 0;
    while (i < nev) {
        double key = ev_key[i];
        int d = 0;
        while (i < nev && ev_key[i] == key) { d += ev_del[i]; i++; }   // dict: equal keys share one counter
        if (d == 0) continue;                                          // zero-net entries are dropped
        tot += d;
        if (tot == mcnt && d == -1 && have_prev) { c_lo[nC] = prev_key; c_hi[nC] = key; nC++; }
        prev_key = key;
        have_prev = true;
    }
    return nC;
}

QCQP_HD void insertion_sort_events(double* key, int* del, int nev)
{
    for (int i = 1; i < nev; i++) {
        double k = key[i];
        int d = del[i];
        int j = i - 1;
        while (j >= 0 && key[j] > k) { key[j + 1] = key[j]; del[j + 1] = del[j]; j--; }
        key[j + 1] = k;
        del[j + 1] = d;
    }
}

// appends the sentinels and the folded single-interval constraints to the explicit (two-interval) events
QCQP_HD int finish_events(const Fold& f, double* ev_key, int* ev_del, int nev)
{
    ev_key[nev] = -QCQP_INF; ev_del[nev] = +1; nev++;
    ev_key[nev] = QCQP_INF; ev_del[nev] = -1; nev++;
    if (f.m1 > 0) {
        ev_key[nev] = f.L; ev_del[nev] = f.m1; nev++;
        ev_key[nev] = f.H; ev_del[nev] = -f.mu; nev++;
    }
    return nev;
}

// ---------------------------------------------------------------------------------------------------------
// register-resident sweep for at most 8 events (the common case: no or one two-interval constraint):
// Batcher's 19-comparator network, then the same merge/scan as sweep_sorted.  Pieces (at most 4) go to c_lo/c_hi.
// ---------------------------------------------------------------------------------------------------------
#define QCQP_CSWAP(a, b)                                                          \
    do {                                                                          \
        if (k[a] > k[b]) { double tk = k[a]; k[a] = k[b]; k[b] = tk; int td = d[a]; d[a] = d[b]; d[b] = td; } \
    } while (0)

QCQP_HD int sweep_small8(const Fold& f, bool has_two, const Ival& I0, const Ival& I1, double* c_lo, double* c_hi)
{
    double k[8];
    int d[8];
    k[0] = -QCQP_INF; d[0] = +1;
    k[1] = QCQP_INF; d[1] = -1;
    if (f.m1 > 0) { k[2] = f.L; d[2] = f.m1; k[3] = f.H; d[3] = -f.mu; }
    else { k[2] = QCQP_INF; d[2] = 0; k[3] = QCQP_INF; d[3] = 0; }       // pads join the +inf sentinel and add 0
    if (has_two) { k[4] = I0.lo; d[4] = +1; k[5] = I0.hi; d[5] = -1; k[6] = I1.lo; d[6] = +1; k[7] = I1.hi; d[7] = -1; }
    else { k[4] = k[5] = k[6] = k[7] = QCQP_INF; d[4] = d[5] = d[6] = d[7] = 0; }
    QCQP_CSWAP(0, 1); QCQP_CSWAP(2, 3); QCQP_CSWAP(4, 5); QCQP_CSWAP(6, 7);
    QCQP_CSWAP(0, 2); QCQP_CSWAP(1, 3); QCQP_CSWAP(4, 6); QCQP_CSWAP(5, 7);
    QCQP_CSWAP(1, 2); QCQP_CSWAP(5, 6);
    QCQP_CSWAP(0, 4); QCQP_CSWAP(1, 5); QCQP_CSWAP(2, 6); QCQP_CSWAP(3, 7);
    QCQP_CSWAP(2, 4); QCQP_CSWAP(3, 5);
    QCQP_CSWAP(1, 2); QCQP_CSWAP(3, 4); QCQP_CSWAP(5, 6);
    int nC = 0;
    long tot = 0;
    double prev_key = 0.0, cur_key = k[0];
    int cur_d = d[0];
    bool have_prev = false;
#pragma unroll
    for (int i = 1; i <= 8; i++) {
        if (i < 8 && k[i] == cur_key) { cur_d += d[i]; continue; }
        if (cur_d != 0) {
            tot += cur_d;
            if (tot == f.mcnt && cur_d == -1 && have_prev) { c_lo[nC] = prev_key; c_hi[nC] = cur_key; nC++; }
            prev_key = cur_key;
            have_prev = true;
        }
        if (i < 8) { cur_key = k[i]; cur_d = d[i]; }
    }
    return nC;
}
#undef QCQP_CSWAP

// ---------------------------------------------------------------------------------------------------------
// HOLE formulation of the sweep line (what cd.cu's warp-parallel solver implements; this scalar statement of it is
// what tests/test_onevar_host.py checks against the reference's dict/sorted sweep, ties and quirks included).
//
// A two-interval feasible set [a0,b0] u [a1,b1] is its hull [a0,b1] -- folded like any single interval -- minus the
// open hole (b0,a1).  The hole adds +1 at -inf and -1 at +inf, i.e. one more always-covering constraint, so the
// events at every finite key are those of the reference.  A zero-width hole nets 0 at its key and is dropped, as the
// reference drops zero-net dict entries.  With L = max lo, H = min hi (multiplicity mu) over the singles and hulls, the
// reference reports a piece ending at key e iff the running total is full just before e and the net count at e is -1:
//   * e = a_i, a hole start:  no other hole starts at e,  max(L, max{b_j : a_j < e}) < e < H;
//   * e = H (finite):         mu == 1, L < H, no hole with a_j <= H <= b_j;
// and the piece starts at the previous nonzero-net key, which is max(L, max{b_j : b_j < e}).  Pieces come out in
// ascending order of e (the H piece last).  No sorting of the 4-per-constraint events is needed: only of the holes.
// ---------------------------------------------------------------------------------------------------------
struct Hole { double a, b; };

// appends constraint i's feasible set (c intervals from feasible_intervals) to the fold / hole list
QCQP_HD bool fold_constraint(Fold& f, int c, const Ival* I, Hole* hole)
{
    f.mcnt++;
    if (c == 0) { f.nempty++; return false; }
    if (c == 1) { f.add_single(I[0].lo, I[0].hi); return false; }
    f.add_single(I[0].lo, I[1].hi);
    if (I[0].hi < I[1].lo) { hole->a = I[0].hi; hole->b = I[1].lo; return true; }
    return false;
}

QCQP_HD int pieces_from_holes(const Fold& f, Hole* h, int nh, double* c_lo, double* c_hi)
{
    if (f.mcnt == 0) { c_lo[0] = -QCQP_INF; c_hi[0] = QCQP_INF; return 1; }   // only the sentinel pair: one piece, all of R
    for (int i = 1; i < nh; i++) {      // sort the holes by their start
        Hole t = h[i];
        int j = i - 1;
        while (j >= 0 && h[j].a > t.a) { h[j + 1] = h[j]; j--; }
        h[j + 1] = t;
    }
    int nC = 0;
    double M = f.L;                     // max(L, b_j of the holes before i)
    for (int i = 0; i < nh; i++) {
        const double e = h[i].a;
        const bool tie = (i > 0 && h[i - 1].a == e) || (i + 1 < nh && h[i + 1].a == e);
        if (!tie && M < e && e < f.H) { c_lo[nC] = M; c_hi[nC] = e; nC++; }
        if (h[i].b > M) M = h[i].b;
    }
    if (f.mu == 1 && f.H < QCQP_INF && f.L < f.H) {
        bool blocked = false;
        double st = f.L;
        for (int i = 0; i < nh; i++) {
            if (h[i].a <= f.H && f.H <= h[i].b) blocked = true;
            if (h[i].b < f.H && h[i].b > st) st = h[i].b;
        }
        if (!blocked) { c_lo[nC] = st; c_hi[nC] = f.H; nC++; }
    }
    return nC;
}

// The same pieces WITHOUT sorting the holes (scalar statement of cd_blk.cu's blk_warp_probe, held against pieces_from_holes by
// tests/test_onevar_host.py): with M_i = max(L, max{b_j : a_j < a_i}), hole i ends a piece [M_i, a_i] iff no other hole starts at
// a_i and M_i < a_i < H.  For an untied hole the holes sorted before it are exactly those with a smaller start; a tied hole is
// never reported, so the M the sorted scan would give it does not matter.  Pieces are ranked by their right end, the H piece last.
// One hole over the whole of (L, H) leaves nothing.  The caller has checked nempty == 0.
QCQP_HD int pieces_from_holes_nosort(const Fold& f, const Hole* h, int nh, double* c_lo, double* c_hi)
{
    if (f.mcnt == 0) { c_lo[0] = -QCQP_INF; c_hi[0] = QCQP_INF; return 1; }
    if (!(f.L < f.H)) return 0;
    for (int i = 0; i < nh; i++)
        if (h[i].a <= f.L && f.H <= h[i].b) return 0;
    int nC = 0;
    bool blocked = false;
    double st = f.L;
    for (int i = 0; i < nh; i++) {
        const double a = h[i].a, b = h[i].b;
        if (a <= f.H && f.H <= b) blocked = true;
        if (b < f.H && b > st) st = b;
        if (!(a > f.L && a < f.H)) continue;        // M_i >= L: only a start inside (L, H) can end a piece
        double M = f.L;
        int eq = 0;
        for (int j = 0; j < nh; j++) {
            if (h[j].a < a && h[j].b > M) M = h[j].b;
            eq += (h[j].a == a) ? 1 : 0;
        }
        if (eq != 1 || !(M < a)) continue;
        int q = nC;
        while (q > 0 && c_hi[q - 1] > a) { c_lo[q] = c_lo[q - 1]; c_hi[q] = c_hi[q - 1]; q--; }
        c_lo[q] = M; c_hi[q] = a;
        nC++;
    }
    if (f.mu == 1 && f.H < QCQP_INF && !blocked) { c_lo[nC] = st; c_hi[nC] = f.H; nC++; }
    return nC;
}

// "Solid" level (scalar statement of the flag blk_warp_probe returns): the feasible sets at this level share NO open interval -- a
// constraint with an empty set, an empty box, one hole over the whole box, or no gap (M_i, a_i) with M_i < a_i < H (tied starts or
// not) together with a hole over the left neighbourhood of H.  A statement about the sets, free of the reference's reporting rules
// (tied starts, coincident right ends); since every constraint's feasible set only grows with the level, a solid level certifies
// that every LOWER level reports no piece.  f: the fold of all constraints at the level; h: its holes.
QCQP_HD bool level_is_solid(const Fold& f, const Hole* h, int nh)
{
    if (f.mcnt == 0) return false;
    if (f.nempty > 0 || !(f.L < f.H)) return true;
    bool cov = false;
    for (int i = 0; i < nh; i++) {
        const double a = h[i].a, b = h[i].b;
        if (a <= f.L && f.H <= b) return true;
        if (a < f.H && f.H <= b) cov = true;
        double M = f.L;
        for (int j = 0; j < nh; j++)
            if (h[j].a < a && h[j].b > M) M = h[j].b;
        if (M < a && a < f.H) return false;           // a gap ends at a_i
    }
    return cov;                                        // no gap at a hole start: solid iff the stretch below H is covered too
}

// ---------------------------------------------------------------------------------------------------------
// the minimiser of f0 = (p, q, r) over the pieces (utilities.py:263-288).  Returns 1 and *xout, 0 for None.
// *err: QCQP_RUN_UNBOUNDED_UNIFORM when the reference would raise OverflowError.
// ---------------------------------------------------------------------------------------------------------
QCQP_HD int choose_point(double p, double q, double r, const double* c_lo, const double* c_hi, int nC, MtRng& rng,
                         double* xout, int* err)
{
    if (nC == 0) return 0;
    if (p == 0.0 && q == 0.0) {
        int idx = rng.choice(nC);
        double lo = c_lo[idx], hi = c_hi[idx];
        if (is_inf(lo) || is_inf(hi)) { *err = QCQP_RUN_UNBOUNDED_UNIFORM; return 0; }
        *xout = rng.uniform(lo, hi);
        return 1;
    }
    const bool have_x0 = (p > 0.0);
    const double x0 = have_x0 ? (-q / (2. * p)) : 0.0;
    if (nC <= 2) {
        // register path for the common one- or two-piece feasible set: every endpoint value is computed once
        const bool two = (nC == 2);
        const double lo0 = c_lo[0], hi0 = c_hi[0];
        const double lo1 = two ? c_lo[1] : 0.0, hi1 = two ? c_hi[1] : 0.0;
        if (have_x0 && ((lo0 <= x0 && x0 <= hi0) || (two && lo1 <= x0 && x0 <= hi1))) { *xout = x0; return 1; }
        const double v0 = onevar_eval(p, q, r, lo0), v1 = onevar_eval(p, q, r, hi0);
        const double v2 = two ? onevar_eval(p, q, r, lo1) : QCQP_INF, v3 = two ? onevar_eval(p, q, r, hi1) : QCQP_INF;
        double bestf = QCQP_INF;
        if (v0 < bestf) bestf = v0;
        if (v1 < bestf) bestf = v1;
        if (v2 < bestf) bestf = v2;
        if (v3 < bestf) bestf = v3;
        const bool m0 = (v0 == bestf), m1 = (v1 == bestf), m2 = two && (v2 == bestf), m3 = two && (v3 == bestf);
        const int cnt = (int)m0 + (int)m1 + (int)m2 + (int)m3;
        if (cnt == 0) return 0;
        int pick = (cnt == 1) ? 0 : rng.choice(cnt);
        if (m0) { if (pick == 0) { *xout = lo0; return 1; } pick--; }
        if (m1) { if (pick == 0) { *xout = hi0; return 1; } pick--; }
        if (m2) { if (pick == 0) { *xout = lo1; return 1; } pick--; }
        *xout = hi1;
        return 1;
    }
    // pass 1: the unconstrained minimiser wins as soon as a piece contains it; otherwise the smallest endpoint value
    double bestf = QCQP_INF;
    for (int i = 0; i < nC; i++) {
        double lo = c_lo[i], hi = c_hi[i];
        if (have_x0 && lo <= x0 && x0 <= hi) { *xout = x0; return 1; }
        double fl = onevar_eval(p, q, r, lo), fr = onevar_eval(p, q, r, hi);
        if (fl < bestf) bestf = fl;
        if (fr < bestf) bestf = fr;
    }
    // pass 2: endpoints attaining it, in order (the reference's bestxs list); NaN values never match
    int cnt = 0;
    double first = 0.0;
    for (int i = 0; i < nC; i++) {
        double lo = c_lo[i], hi = c_hi[i];
        if (onevar_eval(p, q, r, lo) == bestf) { if (cnt == 0) first = lo; cnt++; }
        if (onevar_eval(p, q, r, hi) == bestf) { if (cnt == 0) first = hi; cnt++; }
    }
    if (cnt == 0) return 0;
    if (cnt == 1) { *xout = first; return 1; }    // np.random.choice on a 1-element list draws nothing
    int pick = rng.choice(cnt);
    for (int i = 0; i < nC; i++) {
        if (onevar_eval(p, q, r, c_lo[i]) == bestf) { if (pick == 0) { *xout = c_lo[i]; return 1; } pick--; }
        if (onevar_eval(p, q, r, c_hi[i]) == bestf) { if (pick == 0) { *xout = c_hi[i]; return 1; } pick--; }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// One constraint standing alone (the separable case: every coordinate has exactly one constraint that touches only it).
// Feasible pieces at level s, at most two, returned in registers.
// ---------------------------------------------------------------------------------------------------------
QCQP_HD int single_constraint_pieces(double p, double q, double r, int rel, double s, double* lo0, double* hi0, double* lo1,
                                     double* hi1)
{
    Ival I[2];
    I[0].lo = I[0].hi = I[1].lo = I[1].hi = 0.0;
    const int c = feasible_intervals(p, q, r, rel, s, I);
    *lo0 = *hi0 = *lo1 = *hi1 = 0.0;
    if (c == 0) return 0;
    // The intervals of one constraint come out in ascending order (lo0 <= hi0 <= lo1 <= hi1), so the event list
    // -inf, lo0, hi0[, lo1, hi1], +inf of the reference's sweep (utilities.py:245-261) is already sorted: scan it directly --
    // equal keys share one counter, zero-net keys are dropped, a piece ends where the total falls to m = 1 by exactly -1.
    const bool sorted = (c == 1) ? (I[0].lo <= I[0].hi) : (I[0].lo <= I[0].hi && I[0].hi <= I[1].lo && I[1].lo <= I[1].hi);
    if (sorted) {
        double k[6];
        int d[6];
        int ne;
        k[0] = -QCQP_INF; d[0] = +1;
        k[1] = I[0].lo; d[1] = +1;
        k[2] = I[0].hi; d[2] = -1;
        if (c == 2) { k[3] = I[1].lo; d[3] = +1; k[4] = I[1].hi; d[4] = -1; k[5] = QCQP_INF; d[5] = -1; ne = 6; }
        else { k[3] = QCQP_INF; d[3] = -1; k[4] = k[5] = QCQP_INF; d[4] = d[5] = 0; ne = 4; }
        double cl[2], ch[2];
        cl[0] = cl[1] = ch[0] = ch[1] = 0.0;
        int nC = 0, tot = 0;
        double prev_key = 0.0, cur_key = k[0];
        int cur_d = d[0];
        bool have_prev = false;
#pragma unroll
        for (int i = 1; i <= 6; i++) {
            if (i < ne && k[i] == cur_key) { cur_d += d[i]; continue; }
            if (i <= ne) {
                if (cur_d != 0) {
                    tot += cur_d;
                    if (tot == 1 && cur_d == -1 && have_prev) { if (nC < 2) { cl[nC] = prev_key; ch[nC] = cur_key; } nC++; }
                    prev_key = cur_key;
                    have_prev = true;
                }
                if (i < ne) { cur_key = k[i]; cur_d = d[i]; }
            }
        }
        *lo0 = cl[0]; *hi0 = ch[0]; *lo1 = cl[1]; *hi1 = ch[1];
        return nC;
    }
    Fold f;
    f.init();
    f.mcnt = 1;
    if (c == 1) f.add_single(I[0].lo, I[0].hi);
    double cl[4], ch[4];
    cl[0] = cl[1] = ch[0] = ch[1] = 0.0;
    const int nC = sweep_small8(f, c == 2, I[0], I[1], cl, ch);
    *lo0 = cl[0]; *hi0 = ch[0]; *lo1 = cl[1]; *hi1 = ch[1];
    return nC;
}

// The part of choose_point that needs no random number, for at most two pieces:
//   0 = None, 1 = *xout found deterministically, 2 = the reference would draw (flat objective or tied endpoints).
// FIN: the caller knows every endpoint is finite (memoised with the pieces), so OneVarQuadraticFunction.eval's +-inf branches drop out
template <bool FIN>
QCQP_HD int choose_point_det_t(double p, double q, double r, double lo0, double hi0, double lo1, double hi1, int nC, double* xout)
{
    if (nC == 0) return 0;
    if (p == 0.0 && q == 0.0) return 2;
    const bool two = (nC == 2);
    if (p > 0.0) {
        // x0 = -q / (2p) matters only if it lies in a piece.  With d = 2p > 0, x0 in [lo, hi] needs d lo <= -q <= d hi up to
        // rounding: when -q misses both products by far more than any rounding error (2^-50 relative), the reference's test
        // is false whatever the last bit of the quotient, and the 125-cycle division is skipped.  Otherwise: the exact test.
        const double d = 2. * p, mq = -q;
        const double a0 = d * lo0, b0 = d * hi0;
        bool out = (mq < a0 - fabs(a0) * 8.9e-16) || (mq > b0 + fabs(b0) * 8.9e-16);
        if (out && two) {
            const double a1 = d * lo1, b1 = d * hi1;
            out = (mq < a1 - fabs(a1) * 8.9e-16) || (mq > b1 + fabs(b1) * 8.9e-16);
        }
        if (!out) {
            const double x0 = -q / (2. * p);
            if ((lo0 <= x0 && x0 <= hi0) || (two && lo1 <= x0 && x0 <= hi1)) { *xout = x0; return 1; }
        }
    }
    const double v0 = FIN ? lo0 * (p * lo0 + q) + r : onevar_eval(p, q, r, lo0);
    const double v1 = FIN ? hi0 * (p * hi0 + q) + r : onevar_eval(p, q, r, hi0);
    const double v2 = two ? (FIN ? lo1 * (p * lo1 + q) + r : onevar_eval(p, q, r, lo1)) : QCQP_INF;
    const double v3 = two ? (FIN ? hi1 * (p * hi1 + q) + r : onevar_eval(p, q, r, hi1)) : QCQP_INF;
    // smallest non-NaN value, +inf if there is none: what the reference's running `if v < bestf` leaves behind (fmin ignores NaN)
    const double bestf = fmin(fmin(fmin(v0, v1), fmin(v2, v3)), QCQP_INF);
    const bool m0 = (v0 == bestf), m1 = (v1 == bestf), m2 = two && (v2 == bestf), m3 = two && (v3 == bestf);
    const int cnt = (int)m0 + (int)m1 + (int)m2 + (int)m3;
    if (cnt == 0) return 0;
    if (cnt > 1) return 2;
    *xout = m0 ? lo0 : (m1 ? hi0 : (m2 ? lo1 : hi1));
    return 1;
}
QCQP_HD int choose_point_det(double p, double q, double r, double lo0, double hi0, double lo1, double hi1, int nC, double* xout)
{
    return choose_point_det_t<false>(p, q, r, lo0, hi0, lo1, hi1, nC, xout);
}

}  // namespace qcqp
