// host_shim.cpp -- TEST-ONLY: exposes the __host__ __device__ 1-D solver of onevar.cuh to the CPU test-suite
// (tests/test_onevar_host.py), so the device logic (intervals, fold, sweep line, minimiser choice, MT19937) can be
// checked against the golden vectors on the GPU-less build box.  Built as libqcqp_b200_hostshim.so; it is NOT part of
// libqcqp_b200.so and the qcqp_b200 package never loads it.
#include <stdint.h>

#include <vector>

#include "../../include/qcqp_b200.h"
#include "onevar.cuh"

using namespace qcqp;

// single-thread composition of the pieces the CD kernel runs per warp (cd.cu: solve_level)
extern "C" int qcqp_shim_onevar_qcqp(const double* f0 /*p,q,r*/, const double* fs /*[m][3]*/, const int32_t* relops, int32_t m,
                                     double s, qcqp_rng_state* st, double* xout, int32_t force_general)
{
    std::vector<double> ev_key(4 * (size_t)m + 8), c_lo(2 * (size_t)m + 4), c_hi(2 * (size_t)m + 4);
    std::vector<int> ev_del(4 * (size_t)m + 8);
    Fold fold;
    fold.init();
    int nev = 0, ntwo = 0;
    Ival T0{0, 0}, T1{0, 0};
    for (int i = 0; i < m; i++) {
        double p = fs[3 * i], q = fs[3 * i + 1], r = fs[3 * i + 2];
        fold.mcnt++;   // the caller has already dropped (p, q) == (0, 0) forms, as qcqp.py:116 does
        Ival I[2];
        int c = feasible_intervals(p, q, r, relops[i], s, I);
        if (c == 0) fold.nempty++;
        else if (c == 1) fold.add_single(I[0].lo, I[0].hi);
        else {
            ntwo++; T0 = I[0]; T1 = I[1];
            ev_key[nev] = I[0].lo; ev_del[nev++] = +1; ev_key[nev] = I[0].hi; ev_del[nev++] = -1;
            ev_key[nev] = I[1].lo; ev_del[nev++] = +1; ev_key[nev] = I[1].hi; ev_del[nev++] = -1;
        }
    }
    if (fold.nempty > 0) return 0;
    int nC;
    if (force_general == 2) {
        // the hole formulation (cd.cu: solve_level / holes_finish)
        std::vector<Hole> holes((size_t)m + 1);
        Fold hf;
        hf.init();
        int nh = 0;
        for (int i = 0; i < m; i++) {
            Ival I[2];
            int c = feasible_intervals(fs[3 * i], fs[3 * i + 1], fs[3 * i + 2], relops[i], s, I);
            if (fold_constraint(hf, c, I, &holes[nh])) nh++;
        }
        nC = pieces_from_holes(hf, holes.data(), nh, c_lo.data(), c_hi.data());
    } else if (force_general == 3) {
        // the sort-free hole formulation (cd_blk.cu: blk_warp_probe)
        std::vector<Hole> holes((size_t)m + 1);
        Fold hf;
        hf.init();
        int nh = 0;
        for (int i = 0; i < m; i++) {
            Ival I[2];
            int c = feasible_intervals(fs[3 * i], fs[3 * i + 1], fs[3 * i + 2], relops[i], s, I);
            if (fold_constraint(hf, c, I, &holes[nh])) nh++;
        }
        nC = pieces_from_holes_nosort(hf, holes.data(), nh, c_lo.data(), c_hi.data());
    } else if (ntwo <= 1 && !force_general) {
        nC = sweep_small8(fold, ntwo == 1, T0, T1, c_lo.data(), c_hi.data());   // the kernel's register path
    } else {
        nev = finish_events(fold, ev_key.data(), ev_del.data(), nev);
        insertion_sort_events(ev_key.data(), ev_del.data(), nev);
        nC = sweep_sorted(ev_key.data(), ev_del.data(), nev, fold.mcnt, c_lo.data(), c_hi.data());
    }
    MtRng rng{st->key, st->pos};
    int err = 0;
    int ok = choose_point(f0[0], f0[1], f0[2], c_lo.data(), c_hi.data(), nC, rng, xout, &err);
    st->pos = rng.pos;
    return err ? -err : ok;
}

extern "C" int qcqp_shim_feasible_intervals(double p, double q, double r, int32_t relop, double s, double* out4)
{
    Ival I[2];
    int c = feasible_intervals(p, q, r, relop, s, I);
    for (int i = 0; i < c; i++) { out4[2 * i] = I[i].lo; out4[2 * i + 1] = I[i].hi; }
    return c;
}

extern "C" double qcqp_shim_uniform(qcqp_rng_state* st, double lo, double hi)
{
    MtRng rng{st->key, st->pos};
    double v = rng.uniform(lo, hi);
    st->pos = rng.pos;
    return v;
}

extern "C" int qcqp_shim_choice(qcqp_rng_state* st, int32_t n)
{
    MtRng rng{st->key, st->pos};
    int v = rng.choice(n);
    st->pos = rng.pos;
    return v;
}

// separable case: pieces of a single constraint + the RNG-free part of the minimiser choice (cd_lpc.cu)
extern "C" int qcqp_shim_single_det(const double* f0, const double* f, int32_t relop, double s, double* xout, double* pieces4, int32_t* nC)
{
    double lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
    int c = single_constraint_pieces(f[0], f[1], f[2], relop, s, &lo0, &hi0, &lo1, &hi1);
    pieces4[0] = lo0; pieces4[1] = hi0; pieces4[2] = lo1; pieces4[3] = hi1;
    *nC = c;
    const int rc = choose_point_det(f0[0], f0[1], f0[2], lo0, hi0, lo1, hi1, c, xout);
    // the finite-endpoint variant the kernel uses on its hot path must agree with the general one whenever it applies
    const bool fin = c > 0 && c <= 2 && !is_inf(lo0) && !is_inf(hi0) && (c < 2 || (!is_inf(lo1) && !is_inf(hi1)));
    if (fin) {
        double x2 = 0.0;
        const int rc2 = choose_point_det_t<true>(f0[0], f0[1], f0[2], lo0, hi0, lo1, hi1, c, &x2);
        if (rc2 != rc || (rc == 1 && !(x2 == *xout))) return -100;
    }
    return rc;
}

// ---------------------------------------------------------------------------------------------------------
// Phase-1 bisection of one coordinate (qcqp.py:113-140) with the zero objective, two ways:
//   mode 0: the reference's loop, one probe per level (hole formulation + choose_point);
//   mode 1: cd_blk.cu's search -- the chain c_1 = (ss + es) / 2, c_{m+1} = (c_m + es) / 2; a SOLID level certifies every lower level
//           infeasible, so only the first non-solid level of the chain is probed for real; `nw` levels are examined per round, spread
//           over the open index range (lowest and deepest included), as the CTA's warps do.
// out[0..3] = new_xi, new_viol, ss, es at the end; returns the number of levels examined (mode 1) / probes (mode 0), < 0 on error.
// tests/test_onevar_host.py holds the two against each other (results and MT19937 position) on random and tie-heavy constraint sets.
// ---------------------------------------------------------------------------------------------------------
namespace {
struct ProbeOut { int nC; bool solid; std::vector<double> lo, hi; };
ProbeOut probe_level(const double* fs, const int32_t* relops, int m, double s)
{
    ProbeOut o;
    std::vector<Hole> holes((size_t)m + 1);
    Fold f;
    f.init();
    int nh = 0;
    for (int i = 0; i < m; i++) {
        Ival I[2];
        const int c = feasible_intervals(fs[3 * i], fs[3 * i + 1], fs[3 * i + 2], relops[i], s, I);
        if (fold_constraint(f, c, I, &holes[nh])) nh++;
    }
    o.lo.assign((size_t)m + 4, 0.0); o.hi.assign((size_t)m + 4, 0.0);
    o.nC = (f.nempty > 0) ? 0 : pieces_from_holes_nosort(f, holes.data(), nh, o.lo.data(), o.hi.data());
    o.solid = level_is_solid(f, holes.data(), nh);
    return o;
}
}  // namespace

extern "C" int qcqp_shim_phase1_bisect(const double* fs /*[m][3]*/, const int32_t* relops, int32_t m, double ss, double es, double tol,
                                       qcqp_rng_state* st, int32_t mode, int32_t nw, double* out4)
{
    MtRng rng{st->key, st->pos};
    double new_xi = 0.0, new_viol = es;
    int work = 0;
    auto draw = [&](const ProbeOut& o, double* x) {
        const int idx = rng.choice(o.nC);
        if (is_inf(o.lo[idx]) || is_inf(o.hi[idx])) return false;
        *x = rng.uniform(o.lo[idx], o.hi[idx]);
        return true;
    };
    if (mode == 0) {
        while (es - ss > tol) {
            const double s = (ss + es) / 2;
            const ProbeOut o = probe_level(fs, relops, m, s);
            work++;
            if (o.nC == 0) ss = s;
            else { double x; if (!draw(o, &x)) { st->pos = rng.pos; return -2; } new_xi = x; new_viol = s; es = s; }
        }
    } else {
        while (es - ss > tol) {
            std::vector<double> lev(1, ss);
            for (double cl = ss; es - cl > tol && lev.size() < 4096;) { cl = (cl + es) / 2; lev.push_back(cl); }
            const int K = (int)lev.size() - 1;
            int lo = 0, hi = K + 1, keep = -1;
            ProbeOut hi_out;
            while (hi - lo > 1) {
                const int span = hi - lo - 1, nev = (keep < 0) ? nw : nw - 1;
                std::vector<int> idx;
                for (int slot = 0; slot < nev; slot++) {
                    int mm = -1;
                    if (span <= nev) mm = (slot < span) ? lo + 1 + slot : -1;
                    else mm = (slot == 0 || nev == 1) ? lo + 1 : lo + 1 + (int)(((long long)slot * (span - 1) + nev - 2) / (nev - 1));
                    if (mm > 0) idx.push_back(mm);
                }
                int nlo = lo, nhi = hi;
                for (int mm : idx) {
                    const ProbeOut o = probe_level(fs, relops, m, lev[(size_t)mm]);
                    work++;
                    if (o.solid) { if (mm > nlo) nlo = mm; }
                    else if (mm < nhi) { nhi = mm; hi_out = o; keep = 0; }
                }
                lo = nlo; hi = nhi;
                if (hi <= lo) { hi = K + 1; keep = -1; hi_out = ProbeOut(); }
            }
            const double s_lo = lev[(size_t)lo], s_hi = (hi <= K) ? lev[(size_t)hi] : 0.0;
            if (lo >= 1) ss = s_lo;
            if (hi <= K) {
                if (hi_out.nC > 0) { double x; if (!draw(hi_out, &x)) { st->pos = rng.pos; return -2; } new_xi = x; new_viol = s_hi; es = s_hi; }
                else ss = s_hi;
            }
        }
    }
    st->pos = rng.pos;
    out4[0] = new_xi; out4[1] = new_viol; out4[2] = ss; out4[3] = es;
    return work;
}
