// cd_lpc2.cu -- phase 2 of improve_coord_descent (coord_descent_phase2, qcqp.py:152-178) for SEPARABLE problems with a DENSE
// objective (Boolean least squares: the C2 benchmark path).  Same decisions and the same results, bit for bit, as stage 2 of
// cd_lpc_kernel (cd_lpc.cu) -- a restart still walks the coordinates 32 at a time, blocked Gauss-Seidel through the 32 x 32
// diagonal block of P_0 -- but a restart is now a CTA of four warps with different jobs that talk through mbarriers only:
//
//   warp 0, the RESOLVER, owns the serial chain of one-variable decisions (get_onevar_func utilities.py:99-105 through the
//           cached g = P_0 x, onevar_qcqp :241-288 on the memoised pieces; in the common case a certified threshold test on g_k,
//           see "certified classification" below).  Its loop is branch-free SIMD work plus one ballot and two shuffles per
//           move; everything else is POSTED as a command (two shared-memory stores by lane 0 and an mbarrier arrive by all 32 lanes,
//           no fence): "row k moved by delta", "publish g of block b", "stage the block and constants of pass kk".
//   warp 1, the COPY warp, turns commands into TMA traffic: for a move, a bulk copy (cp.async.bulk) of the moved row of P_0 into
//           a ring of S row slots in shared memory; for a stage command, the 32 x 32 diagonal block of P_0 (one
//           cp.async.bulk.tensor.2d box) and the 48 bytes of constants of each of its coordinates.  Up to S rows are in flight
//           with no registers held.
//   warps 2..3, the HELPERS, own g = P_0 x IN REGISTERS (each half of the columns: helper h, lane l, register u  <->  16-byte
//           chunk 32 (NH u + h) + l of g) and consume the row ring in order: g += delta * row from shared memory.  On a publish
//           command the helper that owns the next 32 coordinates writes their g to an exchange area; the commands are consumed
//           in order, so its completion also tells the resolver that every earlier row has been applied.
//
// Nothing on the resolver's path is a fence (st.release / ld.acquire on shared memory compile to MEMBAR.ALL.CTA, which waits for
// the resolver's outstanding global loads: ~1000 cycles per move in the first version of this kernel), an elected-lane section or
// a uniform-datapath TMA issue.  x lives in the output array X itself (lane k mod 32 is the only thread that ever touches x_k)
// and the MT19937 state stays in HBM (phase 2 draws only on exact ties).  Shared memory per restart at n = 1000: 16 KB blocks +
// 3 KB constants + 4 x 8000 B row slots: four restarts per SM.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cd_shared.cuh"
#include "common.cuh"
#include "onevar.cuh"

namespace qcqp {

constexpr int LPC2_BLK = 32 * 32;
constexpr int LPC2_CQ = 64;           // command slots (at most 35 are outstanding: the resolver drains the queue every pass)
// commands: k >= 0 "row k moved"; -1 - b "publish g of block b"; <= STAGE "stage pass": -(STAGE) = 2 kk + buffer
enum { LPC2_CMD_FIN = -2000000000, LPC2_CMD_REFRESH = -1999999999, LPC2_CMD_STAGE = -1000000000 };

// one 2-D box of the tensor map -> shared memory, completion (bytes) on an mbarrier.  SASS: UTMALDG
__device__ __forceinline__ void tma_box_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}

// mbarrier wait that lets the hardware suspend the warp for up to ~1 ms per attempt (the default attempt of try_wait is short, and
// a polling worker takes issue slots from the resolvers that share its scheduler)
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)
            : "memory");
    }
}

struct Lpc2Mem {
    double* blk;       // [2][32][32]  diagonal blocks of P_0 (TMA destinations, 128-byte aligned)
    double* cbuf;      // [2][32][6]   per-coordinate constants of the pass (bulk-copy destinations, same mbarriers as blk)
    double* rows;      // [npad]       the moved row of P_0 in flight (bulk-copy destination)
    double* gx;        // [32]         g of the next 32 coordinates, published by the helper that owns them
    double* fpart;     // [NH]         partial sums of f_0 after a refresh
    double* cq_dl;     // [CQ]         command queue: delta of a move
    uint64_t* fullD;   // [2]          block + constants of a pass have landed
    uint64_t* cfull;   // [CQ]         command posted (32 arrivals: every resolver lane)
    uint64_t* full;    // [1]          the moved row has landed in the row slot
    uint64_t* pub;     // [1]          publish / refresh done
    int* cq_cmd;       // [CQ]
};

// ctr[0] rows of P_0 applied (moves), ctr[1] diagonal blocks fetched, ctr[2] rows read by from-scratch refreshes,
// ctr[3] bytes requested from L2 by this launch (rows + blocks + constants + refreshes + G)
template <int CH, bool PROF>   // CH: 16-byte chunks of g per worker lane (rows of up to 64 CH doubles); PROF: clock64 breakdown
__global__ void __launch_bounds__(64, 7)
cd_lpc2_kernel(const __grid_constant__ CUtensorMap tmapD, PackView P, LpcView V, CdK prm, int R, qcqp_rng_state* rngs, double* X,
               const double* __restrict__ G, qcqp_cd_stats* stats_out, unsigned long long* ctr, unsigned long long* prof_, double* xmirror)
{
    unsigned long long* const prof = PROF ? prof_ : nullptr;
    constexpr int CQ = LPC2_CQ;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = P.n;
    const int npad = (n + 1) & ~1;
    const int n2 = npad >> 1;
    const size_t rr = blockIdx.x;
    if ((int)rr >= R) return;
    Lpc2Mem w;
    w.blk = reinterpret_cast<double*>(smem);
    w.cbuf = w.blk + 2 * LPC2_BLK;
    w.rows = w.cbuf + 2 * 32 * 6;
    w.gx = w.rows + npad;
    w.fpart = w.gx + 32;
    w.cq_dl = w.fpart + 4;
    w.fullD = reinterpret_cast<uint64_t*>(w.cq_dl + CQ);
    w.cfull = w.fullD + 2;
    w.full = w.cfull + CQ;
    w.pub = w.full + 1;
    w.cq_cmd = reinterpret_cast<int*>(w.pub + 1);
    double* xg = X + rr * (size_t)n;                       // x of this restart, in place in the output array
    const uint32_t row_bytes = (uint32_t)npad * 8;

    if (threadIdx.x == 0) {
        mbar_init(&w.fullD[0], 1);
        mbar_init(&w.fullD[1], 1);
        mbar_init(&w.pub[0], 1);
        mbar_init(&w.full[0], 1);
        for (int e = 0; e < CQ; e++) mbar_init(&w.cfull[e], 32);
        mbar_fence_init();
    }
    __syncthreads();

    // ------------------------------------------------ worker ------------------------------------------------
    if (warp == 1) {
        double2 greg[CH];                                   // g = P_0 x: lane l, register u  <->  16-byte chunk 32 u + l
        {
            const double2* gr = reinterpret_cast<const double2*>(G + rr * (size_t)npad);   // one GEMM computed it for all restarts
#pragma unroll
            for (int u = 0; u < CH; u++) {
                const int c = 32 * u + lane;
                greg[u] = (c < n2) ? gr[c] : make_double2(0.0, 0.0);
            }
            if (npad > n) {                                  // the padding entry of an odd n takes no part
#pragma unroll
                for (int u = 0; u < CH; u++) if (32 * u + lane == n2 - 1) greg[u].y = 0.0;
            }
        }
        int e = 0, epar = 0;                 // command queue position and phase
        int rpar = 0;                        // phase of the row slot's mbarrier
        for (;;) {
            mbar_wait_sleep(&w.cfull[e], epar);
            const int cmd = w.cq_cmd[e];
            const double delta = w.cq_dl[e];
            if (++e == CQ) { e = 0; epar ^= 1; }
            if (cmd >= 0) {
                // g += delta * (row of P_0): one bulk copy of the row (row k == column k) into the slot, then fold it in
                if (lane == 0) {
                    mbar_arrive_expect_tx(&w.full[0], row_bytes);
                    bulk_g2s(w.rows, P.dense_P + (size_t)cmd * P.ld, row_bytes, &w.full[0]);
                }
                mbar_wait_sleep(&w.full[0], rpar); rpar ^= 1;
                const double2* row = reinterpret_cast<const double2*>(w.rows);
#pragma unroll
                for (int u = 0; u < CH; u++) {
                    const int c = 32 * u + lane;
                    if (c < n2) {
                        const double2 rv = row[c];
                        greg[u].x = fma(rv.x, delta, greg[u].x);
                        greg[u].y = fma(rv.y, delta, greg[u].y);
                    }
                }
                __syncwarp();                                    // every lane has read the slot before the next copy may overwrite it
            } else if (cmd <= LPC2_CMD_STAGE && cmd > LPC2_CMD_REFRESH) {
                // the 32 x 32 diagonal block of P_0 (one TMA box) and the six constants of its coordinates (one bulk copy), both
                // completing on the buffer's mbarrier
                const int code = LPC2_CMD_STAGE - cmd, b_ = code & 1, kk = code >> 1;
                if (lane == 0) {
                    const uint32_t cb = (uint32_t)((n - kk < 32) ? (n - kk) : 32) * 48u;
                    mbar_arrive_expect_tx(&w.fullD[b_], LPC2_BLK * 8 + cb);
                    tma_box_2d(w.blk + b_ * LPC2_BLK, &tmapD, kk, kk, &w.fullD[b_]);
                    bulk_g2s(w.cbuf + b_ * 32 * 6, V.cst6 + (size_t)kk * 6, cb, &w.fullD[b_]);
                }
            } else if (cmd == LPC2_CMD_FIN) {
                break;
            } else if (cmd == LPC2_CMD_REFRESH) {
                // g = P_0 x from scratch (x read past L1: the resolver's lanes wrote it), then x.(g + q_0)
                double2 acc[CH];
#pragma unroll
                for (int u = 0; u < CH; u++) acc[u] = make_double2(0.0, 0.0);
                for (int r0 = 0; r0 < n; r0++) {
                    const double xa = __ldcg(&xg[r0]);
                    const double2* ra = reinterpret_cast<const double2*>(P.dense_P + (size_t)r0 * P.ld);
#pragma unroll
                    for (int u = 0; u < CH; u++) {
                        const int c = 32 * u + lane;
                        const double2 va = (c < n2) ? __ldg(&ra[c]) : make_double2(0.0, 0.0);
                        acc[u].x = fma(va.x, xa, acc[u].x); acc[u].y = fma(va.y, xa, acc[u].y);
                    }
                }
                double part = 0.0;
#pragma unroll
                for (int u = 0; u < CH; u++) {
                    const int c = 32 * u + lane;
                    greg[u] = acc[u];
                    if (c < n2) {
                        const int k = 2 * c;
                        part = fma(__ldcg(&xg[k]), acc[u].x + V.o_q[k], part);
                        if (k + 1 < n) part = fma(__ldcg(&xg[k + 1]), acc[u].y + V.o_q[k + 1], part);
                        else greg[u].y = 0.0;
                    }
                }
                part = warp_sum(part);
                if (lane == 0) w.fpart[0] = part;
                __syncwarp();
                if (lane == 0) mbar_arrive(&w.pub[0]);
            } else {
                // PUBLISH block b: coordinates 32 b .. 32 b + 31 = chunks 16 b .. 16 b + 15: register b / 2 of lanes 16 (b & 1) ..
                const int b = -1 - cmd;
                const int uu = b >> 1;
                double2 v = make_double2(0.0, 0.0);
#pragma unroll
                for (int u = 0; u < CH; u++) if (u == uu) v = greg[u];
                if ((lane >> 4) == (b & 1)) reinterpret_cast<double2*>(w.gx)[lane & 15] = v;
                __syncwarp();
                if (lane == 0) mbar_arrive(&w.pub[0]);
            }
        }
        return;
    }

    // ------------------------------------------------ resolver ----------------------------------------------
    qcqp_cd_stats st = stats_out[rr];
    bool dead = st.status != QCQP_RUN_OK;
    int pos = rngs[rr].pos;
    const double tol = prm.tol, viol_tol = prm.viol_tol;
    unsigned long long c_rows = 0, c_blk = 0, c_ref = 0;
    long long pt_drain = 0, pt_mbar = 0, pt_res = 0, pt_pro = 0, pt_a = 0, pt_b = 0, pt_cls = 0, pt_push = 0, pt_c = 0;
    const long long pt_begin = prof ? clock64() : 0;
    // command queue producer state (warp-uniform).  A command is two words stored by lane 0 (predicated, no branch) and an arrive
    // by ALL 32 lanes: no election, no fence; the phase completes with lane 0's own arrive, which follows its stores.
    int ce = 0;
    auto post = [&](int cmd, double delta) {
        if (lane == 0) { w.cq_cmd[ce] = cmd; w.cq_dl[ce] = delta; }
        mbar_arrive(&w.cfull[ce]);
        ce = (ce + 1 == CQ) ? 0 : ce + 1;
    };
    int pubpar = 0;                    // phase of the publish barrier the next wait is for

    double mv = -QCQP_INF;             // improve_coord_descent's gate (qcqp.py:189)
    for (int k = lane; k < n; k += 32) {
        const double v = violation_of(V.c_rel[k], onevar_eval(V.c_p[k], V.c_q[k], V.c_r[k], xg[k]));
        mv = (v > mv) ? v : mv;
    }
    mv = warp_max(mv);
    if (!dead && mv < viol_tol) {
        st.ran_phase2 = 1;
        const double viol_p2 = mv;                               // frozen (qcqp.py:157)
        // f_0(x) = f0base + (sum over lanes of dfl): only the exact path reads it, so it is kept as per-lane increments
        double f0base, dfl = 0.0;
        {
            const double* gr = G + rr * (size_t)npad;
            double acc = 0.0;
            for (int k = lane; k < n; k += 32) acc = fma(xg[k], gr[k] + V.o_q[k], acc);
            f0base = warp_sum(acc) + V.o_r;
        }
        int uc = 0;                 // update_counter of phase 2 (qcqp.py:160): never exceeds n
        bool done = false;
        // per-lane memo of the constraint's pieces at the frozen level
        double mp = 0.0, mq = 0.0, mr = 0.0, ml0 = 0.0, mh0 = 0.0, ml1 = 0.0, mh1 = 0.0;
        int mrel = -1, mnC = 0;
        bool mfin = false;      // every endpoint of the memoised pieces is finite
        unsigned gp = 0;        // passes started so far: pass gp uses block buffer gp & 1, mbarrier phase (gp >> 1) & 1
        // x_k of a pass is loaded two passes ahead (it comes from L2; only lane k mod 32 ever writes it, in its own pass)
        double n_x = (lane < n) ? xg[lane] : 0.0, nn_x = (32 + lane < n) ? xg[32 + lane] : ((lane < n) ? xg[lane] : 0.0);
        post(LPC2_CMD_STAGE - 0, 0.0);                           // block and constants of the first pass
        c_blk++;
        post(-1 - 0, 0.0);                                       // g of the first 32 coordinates
        for (int t = 0; t < prm.num_iters && !done; t++) {
            st.sweeps_p2++;
            for (int k0 = 0; k0 < n && !done; k0 += 32, gp++) {
                const int B = (n - k0 < 32) ? (n - k0) : 32;
                const bool act = lane < B;
                const int k = k0 + lane;
                const int buf = gp & 1, par = (gp >> 1) & 1;
                if (prof) pt_a = clock64();
                const double* D = w.blk + buf * LPC2_BLK;
                // the block and the constants of the NEXT pass start their way now, into the buffers the previous pass has finished with
                const int kn0 = (k0 + 32 < n) ? (k0 + 32) : 0;
                post(LPC2_CMD_STAGE - (2 * kn0 + (buf ^ 1)), 0.0);
                c_blk++;
                double xk = 0.0, p0 = 0.0, oq = 0.0, gl = 0.0;
                const double c_xk = n_x;
                n_x = nn_x;
                {
                    int kn = kn0 + 32;                           // two passes ahead; the sweep starts over after the last block
                    if (kn >= n) kn = 0;
                    kn += lane;
                    if (kn < n) nn_x = xg[kn];
                }
                // ---- certified classification of the one-variable decision ----
                // With p0 > 0 and at most two finite, non-degenerate pieces [lo0,hi0] < [lo1,hi1] the minimiser of (p0, q0, .) over the
                // pieces is decided by where -q0 lies among  2p0 lo0 < 2p0 hi0 < p0 (hi0 + lo1) < 2p0 lo1 < 2p0 hi1  (the closest
                // endpoint to x0 = -q0 / 2p0, or x0 itself inside a piece).  -q0 = c - 2 g_k with c = 2 p0 x_k - q_0[k] is affine in the
                // cached g_k, so the five thresholds are constants of the lane for the pass, expressed on g_k: GA > GB > GM > GC > GE.
                // A lane whose g_k is farther than a relative 1e-9 from every threshold (rounding is ~1e-16) and not inside a piece
                // takes the endpoint of its region without evaluating anything; everything else -- near a threshold (exact ties of
                // the reference included: equality is always inside the band), x0 inside a piece, p0 <= 0, unbounded or degenerate
                // pieces -- goes through the reference's own arithmetic (choose_point_det / choose_point) when its turn comes.
                double GA = 0.0, GB = 0.0, GM = 0.0, GC = 0.0, GE = 0.0, gsc = 0.0;
                bool fastok = false;
                unsigned wmask = 0;          // bit r: the endpoint of region r is farther than tol from x_k (the step would move)
                // g of this pass's coordinates: published by its owner once every earlier row had been applied
                {
                    const long long q0c = prof ? clock64() : 0;
                    mbar_wait(&w.pub[0], pubpar); pubpar ^= 1;
                    if (prof) pt_drain += clock64() - q0c;
                }
                // this pass's block and constants (posted a pass ago); every armed phase of the mbarrier is waited on exactly once
                {
                    const long long q0c = prof ? clock64() : 0;
                    mbar_wait(&w.fullD[buf], par);
                    if (prof) pt_mbar += clock64() - q0c;
                }
                if (act) {
                    const double* cc = w.cbuf + (buf * 32 + lane) * 6;
                    const double p = cc[0], q = cc[1], r = cc[2];
                    const int rel = (int)cc[3];
                    const double c_odk = cc[4], c_oqk = cc[5];
                    xk = c_xk;
                    if (!(mrel == rel && mp == p && mq == q && mr == r)) {
                        mnC = single_constraint_pieces(p, q, r, rel, viol_p2, &ml0, &mh0, &ml1, &mh1);
                        mfin = mnC > 0 && mnC <= 2 && !is_inf(ml0) && !is_inf(mh0) && (mnC < 2 || (!is_inf(ml1) && !is_inf(mh1)));
                        mp = p; mq = q; mr = r; mrel = rel;
                    }
                    p0 = c_odk; oq = c_oqk; gl = w.gx[lane];
                    const bool two = (mnC == 2);
                    fastok = mfin && p0 > 0.0 && ml0 < mh0 && (!two || (mh0 < ml1 && ml1 < mh1));
                    if (fastok) {
                        const double d = 2. * p0, c = d * xk - oq;
                        GA = 0.5 * (c - d * ml0); GB = 0.5 * (c - d * mh0);
                        GM = two ? 0.5 * (c - p0 * (mh0 + ml1)) : -QCQP_INF;
                        GC = two ? 0.5 * (c - d * ml1) : -QCQP_INF;
                        GE = two ? 0.5 * (c - d * mh1) : -QCQP_INF;
                        gsc = 1e-9 * (fabs(c) + fabs(GA) + fabs(two ? GE : GB));
                        wmask = (fabs(ml0 - xk) > tol ? 1u : 0u) | (fabs(mh0 - xk) > tol ? 4u : 0u);
                        if (two) wmask |= (fabs(ml1 - xk) > tol ? 8u : 0u) | (fabs(mh1 - xk) > tol ? 32u : 0u);
                    }
                }
                __syncwarp();                 // every lane has read gx before the next publish can overwrite it
                bool moved_me = false;
                double mv_gl = 0.0, mv_xi = 0.0;
                // this lane's step as it stands: 0 certainly no move, 1 moves to fxi_me (the endpoint of its region), 2 reference arithmetic
                int code_me = 2;
                double fxi_me = 0.0, dl_me = 0.0;
                auto classify = [&]() {
                    const int reg = (int)(gl < GA) + (int)(gl < GB) + (int)(gl < GM) + (int)(gl < GC) + (int)(gl < GE);
                    const double mu = fma(1e-9, fabs(gl), gsc);
                    const bool cert = (fabs(gl - GA) > mu) & (fabs(gl - GB) > mu) & (fabs(gl - GM) > mu) & (fabs(gl - GC) > mu) & (fabs(gl - GE) > mu) &
                                      (reg != 1) & (reg != 4) & fastok;
                    fxi_me = (reg == 0) ? ml0 : ((reg == 2) ? mh0 : ((reg == 3) ? ml1 : mh1));
                    dl_me = fxi_me - xk;
                    code_me = cert ? (int)((wmask >> reg) & 1u) : 2;
                };
                classify();
                int cur = 0;            // next coordinate of the pass to resolve
                int steps32 = 0, upd32 = 0;
                if (prof) { pt_b = clock64(); pt_pro += pt_b - pt_a; }
                while (cur < B) {
                    if (prof) pt_c = clock64();
                    const bool pend = act && lane >= cur;
                    const unsigned m_mv = __ballot_sync(FULL, pend && code_me == 1), m_ex = __ballot_sync(FULL, pend && code_me == 2);
                    const unsigned stop = m_mv | m_ex;
                    const int first = stop ? (__ffs(stop) - 1) : B;
                    const int quiet = first - cur;      // steps that change nothing (qcqp.py:172-176)
                    if (n - uc <= quiet) { steps32 += (n - uc); done = true; break; }
                    uc += quiet;
                    steps32 += quiet;
                    if (first == B) break;
                    steps32++;
                    const int code = ((m_ex >> first) & 1u) ? 2 : 1;
                    double delta = bcast(dl_me, first);
                    double fxi = fxi_me;                        // meaningful in lane `first` only
                    bool moves = (code == 1);
                    if (code == 2) {
                        // the reference's own arithmetic for this one coordinate
                        const double f0val = f0base + warp_sum(dfl + (moved_me ? (mv_xi * (p0 * mv_xi + (2 * (mv_gl - p0 * xk) + oq)) - xk * (p0 * xk + (2 * (mv_gl - p0 * xk) + oq))) : 0.0));
                        int rc = 0;
                        double q0 = 0.0, r0 = 0.0, xi = 0.0;
                        const double fx = bcast(xk, first);
                        if (lane == first) {
                            q0 = 2 * (gl - p0 * xk) + oq;
                            r0 = f0val - xk * (p0 * xk + q0);
                            rc = mfin ? choose_point_det_t<true>(p0, q0, r0, ml0, mh0, ml1, mh1, mnC, &xi) : choose_point_det(p0, q0, r0, ml0, mh0, ml1, mh1, mnC, &xi);
                        }
                        const int frc = bcast_i(rc, first);
                        if (frc == 2) {
                            const double fp0 = bcast(p0, first), fq0 = bcast(q0, first), fr0 = bcast(r0, first);
                            double cl[2], ch[2];
                            cl[0] = bcast(ml0, first); ch[0] = bcast(mh0, first); cl[1] = bcast(ml1, first); ch[1] = bcast(mh1, first);
                            const int nC = bcast_i(mnC, first);
                            int err = 0, fnd = 0;
                            double xv = 0.0;
                            if (lane == 0) {
                                MtRng rng;
                                rng.key = rngs[rr].key; rng.pos = pos;       // the stream stays in HBM: phase 2 draws only on exact ties
                                fnd = choose_point(fp0, fq0, fr0, cl, ch, nC, rng, &xv, &err);
                                pos = rng.pos;
                            }
                            pos = bcast_i(pos, 0); err = bcast_i(err, 0); fnd = bcast_i(fnd, 0); fxi = bcast(xv, 0);
                            if (err) { st.status = err; dead = true; done = true; break; }
                            moves = (fnd != 0) && fabs(fxi - fx) > tol;
                            delta = fxi - fx;
                        } else {
                            const bool mv1 = (rc == 1) && fabs(xi - xk) > tol;
                            moves = bcast_i((int)mv1, first) != 0;
                            fxi = bcast(xi, first);
                            delta = bcast(xi - xk, first);
                        }
                    }
                    if (prof) { const long long q = clock64(); pt_cls += q - pt_c; pt_c = q; }
                    if (moves) {
                        // the moved row of P_0 starts its way to the helpers at once: its round trip overlaps the decisions that follow
                        post(k0 + first, delta);
                        if (prof) { const long long q = clock64(); pt_push += q - pt_c; pt_c = q; }
                        upd32++;
                        uc = 0;
                        // the coordinates still to come see the move through their own entry of column `first`
                        if (lane == first) { moved_me = true; mv_gl = gl; mv_xi = fxi; xg[k] = fxi; }   // a coordinate moves at most once per pass
                        if (act && lane > first) gl = fma(D[first * 32 + lane], delta, gl);     // D[lane][first] = D[first][lane]
                        classify();
                    } else {
                        uc++;
                        if (uc == n) { done = true; break; }
                    }
                    cur = first + 1;
                }
                if (prof) { pt_a = clock64(); pt_res += pt_a - pt_b; }
                st.steps_p2 += steps32; st.updates_p2 += upd32; c_rows += upd32;
                // off the decision chain: this lane's share of f_0(x) = t0 + b (t2 b + t1)
                if (moved_me) {
                    const double q0 = 2 * (mv_gl - p0 * xk) + oq;
                    dfl += mv_xi * (p0 * mv_xi + q0) - xk * (p0 * xk + q0);
                }
                // g of the next 32 coordinates, once every row posted so far has been applied
                if (!done) post(-1 - (kn0 >> 5), 0.0);
            }
            if (!done && prm.refresh_every > 0 && ((t + 1) % prm.refresh_every) == 0) {
                // g and f_0 from scratch by the helpers; the publish that is already queued precedes it, so it is re-issued
                mbar_wait(&w.pub[0], pubpar); pubpar ^= 1;
                __threadfence();                                 // the helpers read x from L2
                post(LPC2_CMD_REFRESH, 0.0);
                mbar_wait(&w.pub[0], pubpar); pubpar ^= 1;
                f0base = V.o_r;
                f0base += w.fpart[0];
                dfl = 0.0;
                c_ref += n;
                post(-1 - 0, 0.0);
            }
        }
        // the stage command posted for a pass that never started must land before the CTA's shared memory is released
        mbar_wait(&w.fullD[gp & 1], (gp >> 1) & 1);
    }
    post(LPC2_CMD_FIN, 0.0);
    __syncwarp();
    // the finished point straight into the caller's pinned result array (when there is one): the read-back of X overlaps the rest
    // of the launch instead of following it
    if (xmirror) {
        double* xm = xmirror + rr * (size_t)n;
        for (int k = lane; k < n; k += 32) xm[k] = xg[k];
    }
    if (lane == 0) {
        rngs[rr].pos = pos;
        stats_out[rr] = st;
        if (prof) {
            prof[rr * 8 + 0] = clock64() - pt_begin; prof[rr * 8 + 1] = pt_pro; prof[rr * 8 + 2] = pt_res; prof[rr * 8 + 3] = pt_mbar;
            prof[rr * 8 + 4] = pt_drain; prof[rr * 8 + 5] = pt_cls; prof[rr * 8 + 6] = pt_push; prof[rr * 8 + 7] = c_rows;
        }
        if (ctr) {
            const unsigned long long rb = row_bytes;
            atomicAdd(&ctr[0], c_rows);
            atomicAdd(&ctr[1], c_blk);
            atomicAdd(&ctr[2], c_ref);
            atomicAdd(&ctr[3], c_rows * rb + c_blk * (unsigned long long)(LPC2_BLK * 8 + 32 * 48) + c_ref * rb + 2 * rb);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side: the tensor map of P_0 (dense slot 0) and the launch
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int lpc2_make_tmap(qcqp_pack* p)
{
    if (p->tmap_state != 0) return p->tmap_state > 0 ? QCQP_OK : QCQP_ERR_CUDA;
    p->tmap_state = -1;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        return fail(QCQP_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    }
    static_assert(sizeof(CUtensorMap) <= sizeof(p->tmap), "tensor map storage");
    const cuuint64_t gdim[2] = {(cuuint64_t)p->v.n, (cuuint64_t)p->v.n};          // {columns (contiguous), rows}
    const cuuint64_t gstr[1] = {(cuuint64_t)p->v.ld * 8};                         // row pitch in bytes (ld is even: multiple of 16)
    const cuuint32_t box[2] = {32, 32};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)fn)((CUtensorMap*)p->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)p->v.dense_P, gdim, gstr, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(QCQP_ERR_CUDA, "cuTensorMapEncodeTiled failed for the objective matrix (code " + std::to_string((int)r) + ")");
    p->tmap_state = 1;
    return QCQP_OK;
}

// shared memory of one restart
size_t lpc2_smem_bytes(int n)
{
    const int npad = (n + 1) & ~1;
    return (size_t)2 * LPC2_BLK * 8 + 2 * 32 * 6 * 8 + (size_t)npad * 8 + 32 * 8 + 4 * 8 + LPC2_CQ * 8 + (2 + LPC2_CQ + 2) * 8 + LPC2_CQ * 4;
}

// rows of up to 32 * 32 sixteen-byte chunks (n <= 2048): g lives in the worker's registers
bool lpc2_supported(int n)
{
    const int n2 = ((n + 1) & ~1) >> 1;
    return n > 64 && n2 <= 32 * 32;
}

template <int CH, bool PROF>
static int lpc2_launch_k(qcqp_pack* p, const CdK& k, int R, qcqp_rng_state* drng, double* dX, const double* G, qcqp_cd_stats* dstats,
                         cudaStream_t stream, size_t smem, unsigned long long* dprof)
{
    QCQP_CUDA_TRY(cudaFuncSetAttribute(cd_lpc2_kernel<CH, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QCQP_CUDA_TRY(cudaFuncSetAttribute(cd_lpc2_kernel<CH, PROF>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CUtensorMap tm;
    memcpy(&tm, p->tmap, sizeof(tm));
    cd_lpc2_kernel<CH, PROF><<<R, 64, smem, stream>>>(tm, p->v, p->lpc, k, R, drng, dX, G, dstats, p->d_ctr, dprof, p->x_mirror);
    if (p->x_mirror) p->x_mirror_done = true;
    QCQP_CUDA_TRY(cudaGetLastError());
    return QCQP_OK;
}

template <int CH>
static int lpc2_launch_t(qcqp_pack* p, const CdK& k, int R, qcqp_rng_state* drng, double* dX, const double* G, qcqp_cd_stats* dstats,
                         cudaStream_t stream, size_t smem)
{
    // QCQP_LPC2_PROF=<file>: per-restart cycle breakdown (clock64) written to <file> as R x 8 uint64 after the launch -- development
    // aid, synchronises the stream
    const char* pf = getenv("QCQP_LPC2_PROF");
    unsigned long long* dprof = nullptr;
    if (pf && pf[0]) {
        QCQP_CUDA_TRY(cudaMalloc((void**)&dprof, (size_t)R * 64));
        QCQP_CUDA_TRY(cudaMemsetAsync(dprof, 0, (size_t)R * 64, stream));
    }
    int rc = dprof ? lpc2_launch_k<CH, true>(p, k, R, drng, dX, G, dstats, stream, smem, dprof)
                   : lpc2_launch_k<CH, false>(p, k, R, drng, dX, G, dstats, stream, smem, nullptr);
    if (rc != QCQP_OK) return rc;
    if (dprof) {
        std::vector<unsigned long long> hp((size_t)R * 8);
        QCQP_CUDA_TRY(cudaStreamSynchronize(stream));
        QCQP_CUDA_TRY(cudaMemcpy(hp.data(), dprof, (size_t)R * 64, cudaMemcpyDeviceToHost));
        cudaFree(dprof);
        if (FILE* fh = fopen(pf, "wb")) { fwrite(hp.data(), 8, hp.size(), fh); fclose(fh); }
    }
    return QCQP_OK;
}

// phase 2 of every restart from the output of stage 1 (X, rng, stats) with g = X P_0 supplied in G
int lpc2_launch(qcqp_pack* p, const CdK& k, int R, qcqp_rng_state* drng, double* dX, const double* G, qcqp_cd_stats* dstats,
                cudaStream_t stream)
{
    int rc = lpc2_make_tmap(p);
    if (rc != QCQP_OK) return rc;
    const int n = p->v.n;
    if (!lpc2_supported(n)) return fail(QCQP_ERR_CAPACITY, "qcqp_cd_improve: n too large for the resolver / worker kernel");
    const size_t smem = lpc2_smem_bytes(n);
    const int n2 = ((n + 1) & ~1) >> 1;
    if (n2 <= 32 * 16) return lpc2_launch_t<16>(p, k, R, drng, dX, G, dstats, stream, smem);
    return lpc2_launch_t<32>(p, k, R, drng, dX, G, dstats, stream, smem);
}

}  // namespace qcqp
