// forms_eval.cuh -- f_j(x) = x'P_j x + q_j'x + r_j from scratch, for one point held by one warp
// (QuadraticFunction.eval, utilities.py:49-50).  Used by the CD kernel (cache refresh, final (f0, maxviol)),
// the batched eval kernel and the SDR sampler.
#pragma once

#include "common.cuh"

namespace qcqp {

constexpr int EVAL_LONG_FORM = 96;   // sparse forms with more stored entries than this are summed by the whole warp

// One lane, one sparse form, in the reference's order: rows ascending, (P x)_i summed over ascending columns with
// separately rounded multiply/add (SciPy csr_matvec), then acc += ((P x)_i + q_i) * x_i, then + r.  The operations come as the
// form's evaluation program (common.cuh: EvOp, built by pack.cu in exactly that order): one 16-byte load per operation whose
// address does not depend on earlier loads, four in flight -- walking the (row, col, val) and (idx, val) lists with their
// data-dependent loop conditions cost one L2 round trip per entry (circle packing's sweep-boundary refresh of 20 701 forms:
// 1.5 M cycles).  Same arithmetic, same order, same bits.
__device__ __forceinline__ double eval_sparse_form_seq(const PackView& P, int j, const double* x)
{
    const long long e0 = P.ev_ptr[j], e1 = P.ev_ptr[j + 1];
    const int4* ops = reinterpret_cast<const int4*>(P.ev_op);
    double acc = 0.0, y = 0.0;
    auto step = [&](const int4& o) {
        const double val = __hiloint2double(o.y, o.x);
        y = (o.z >= 0) ? (y + val * x[o.z]) : (y + val);
        if (o.w >= 0) { acc = acc + y * x[o.w]; y = 0.0; }
    };
    long long e = e0;
    for (; e + 4 <= e1; e += 4) {
        const int4 o0 = __ldg(ops + e), o1 = __ldg(ops + e + 1), o2 = __ldg(ops + e + 2), o3 = __ldg(ops + e + 3);
        step(o0); step(o1); step(o2); step(o3);
    }
    if (e < e1) {
        const int n = (int)(e1 - e);
        const int4 o0 = __ldg(ops + e), o1 = __ldg(ops + (n > 1 ? e + 1 : e)), o2 = __ldg(ops + (n > 2 ? e + 2 : e));
        step(o0);
        if (n > 1) step(o1);
        if (n > 2) step(o2);
    }
    return acc + P.r[j];
}

// whole warp, one sparse form (fast mode): fma partial sums, one shuffle reduction
__device__ __forceinline__ double eval_sparse_form_warp(const PackView& P, int j, const double* x, int lane)
{
    double acc = 0.0;
    for (long long e = P.f_ptr[j] + lane; e < P.f_ptr[j + 1]; e += 32) acc = fma(P.f_val[e] * x[P.f_row[e]], x[P.f_col[e]], acc);
    for (long long e = P.q_ptr[j] + lane; e < P.q_ptr[j + 1]; e += 32) acc = fma(P.q_val[e], x[P.q_idx[e]], acc);
    return warp_sum(acc) + P.r[j];
}

// whole warp, one dense form (fast mode): lanes own column pairs (16-byte loads), four rows in flight per pass so
// the loads of one pass overlap instead of queueing behind each other's L2 latency
__device__ __forceinline__ double eval_dense_form_warp(const PackView& P, int j, const double* x, int lane)
{
    const int n = P.n, ld = P.ld;
    const double* M = P.dense_P + (size_t)P.dense_slot[j] * n * ld;
    const int n2 = n >> 1;
    double acc = 0.0;
    int i = 0;
    for (; i + 4 <= n; i += 4) {
        const double2* r0 = reinterpret_cast<const double2*>(M + (size_t)i * ld);
        const double2* r1 = reinterpret_cast<const double2*>(M + (size_t)(i + 1) * ld);
        const double2* r2 = reinterpret_cast<const double2*>(M + (size_t)(i + 2) * ld);
        const double2* r3 = reinterpret_cast<const double2*>(M + (size_t)(i + 3) * ld);
        double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
        for (int c = lane; c < n2; c += 32) {
            const double2 a0 = r0[c], a1 = r1[c], a2 = r2[c], a3 = r3[c];
            const double xa = x[2 * c], xb = x[2 * c + 1];
            p0 = fma(a0.x, xa, p0); p0 = fma(a0.y, xb, p0);
            p1 = fma(a1.x, xa, p1); p1 = fma(a1.y, xb, p1);
            p2 = fma(a2.x, xa, p2); p2 = fma(a2.y, xb, p2);
            p3 = fma(a3.x, xa, p3); p3 = fma(a3.y, xb, p3);
        }
        if ((n & 1) && lane == 0) {
            const double xl = x[n - 1];
            p0 = fma(M[(size_t)i * ld + n - 1], xl, p0); p1 = fma(M[(size_t)(i + 1) * ld + n - 1], xl, p1);
            p2 = fma(M[(size_t)(i + 2) * ld + n - 1], xl, p2); p3 = fma(M[(size_t)(i + 3) * ld + n - 1], xl, p3);
        }
        acc = fma(p0, x[i], acc); acc = fma(p1, x[i + 1], acc); acc = fma(p2, x[i + 2], acc); acc = fma(p3, x[i + 3], acc);
    }
    for (; i < n; i++) {
        const double* row = M + (size_t)i * ld;
        double part = 0.0;
        for (int c = lane; c < n; c += 32) part = fma(row[c], x[c], part);
        acc = fma(part, x[i], acc);
    }
    for (long long e = P.q_ptr[j] + lane; e < P.q_ptr[j + 1]; e += 32) acc = fma(P.q_val[e], x[P.q_idx[e]], acc);
    return warp_sum(acc) + P.r[j];
}

// one lane, one dense form, reference order (strict mode; slow, test sizes only)
__device__ __forceinline__ double eval_dense_form_seq(const PackView& P, int j, const double* x)
{
    const int n = P.n, ld = P.ld;
    const double* M = P.dense_P + (size_t)P.dense_slot[j] * n * ld;
    long long qe = P.q_ptr[j], qend = P.q_ptr[j + 1];
    double acc = 0.0;
    for (int i = 0; i < n; i++) {
        const double* row = M + (size_t)i * ld;
        double y = 0.0;
        for (int c = 0; c < n; c++) {
            double v = row[c];
            if (v != 0.0) y = y + v * x[c];   // CSR stores no explicit zeros
        }
        if (qe < qend && P.q_idx[qe] == i) { y = y + P.q_val[qe]; qe++; }
        acc = acc + y * x[i];
    }
    return acc + P.r[j];
}

// f_j(x) for j in [j0, j1], every result handed to sink(j, value) by exactly one lane.
// Short sparse forms: one lane each.  Long sparse and dense forms: the whole warp (fast) or one lane (strict).
template <class Sink>
__device__ __forceinline__ void eval_forms(const PackView& P, const double* x, int j0, int j1, bool strict, int lane, Sink sink,
                                           bool skip_dense = false)
{
    for (int base = j0; base <= j1; base += 32) {
        int j = base + lane;
        bool mine = j <= j1;
        bool dense = mine && P.dense_slot[j] >= 0;
        if (dense && skip_dense) { mine = false; dense = false; }
        bool big = mine && !dense && (P.f_ptr[j + 1] - P.f_ptr[j]) > EVAL_LONG_FORM;
        if (mine && !dense && (!big || strict)) sink(j, eval_sparse_form_seq(P, j, x));
        if (mine && dense && strict) sink(j, eval_dense_form_seq(P, j, x));
        if (!strict) {
            unsigned coop = __ballot_sync(FULL, dense || big);
            while (coop) {
                int src = __ffs(coop) - 1;
                coop &= coop - 1;
                int jj = base + src;
                double v = (P.dense_slot[jj] >= 0) ? eval_dense_form_warp(P, jj, x, lane) : eval_sparse_form_warp(P, jj, x, lane);
                if (lane == src) sink(jj, v);
            }
        }
    }
}

}  // namespace qcqp
