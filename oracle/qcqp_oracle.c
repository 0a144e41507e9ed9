/*
 * qcqp_oracle.c -- CPU restatement of the cvxgrp/qcqp hot path, in plain C.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the *checker* for the CUDA engine in qcqp_b200/csrc.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product (qcqp_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/golden/ holds vectors minted by executing the unmodified reference
 * (tests/golden/make_golden.py, through oracle/ref_harness.py); tests/test_oracle_golden.py checks
 * this file against every one of them (G1..G5, Q1..Q5 of SURVEY.md section 8c and more).
 *
 * All line citations are to /root/reference (qcqp 0.8.3).
 *
 * Two modes share one code path:
 *   faithful (fast=0)  every get_onevar_func recomputes t0 = (P z + q).z + r from scratch and loops
 *                      over all m constraints, exactly as utilities.py:99-105 / qcqp.py:115,164 do.
 *   fast     (fast=1)  same decisions, but t0 comes from a cached f_j(x) and only the structurally
 *                      incident forms of coordinate k are visited.  This is the strong CPU baseline
 *                      bench.py reports (kind "port").
 *
 * Compile:  gcc -O2 -fPIC -shared -pthread -ffp-contract=off -fno-fast-math
 *           (no FMA contraction: NumPy/SciPy round every multiply and add separately.)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>


/* ------------------------------------------------------------------------------------------
 * tiny pthread parallel-for (dynamic schedule) for the batch drivers; restarts are independent.
 * ---------------------------------------------------------------------------------------- */
typedef void (*pf_body)(int idx, void* ctx);
typedef struct { pf_body body; void* ctx; int count; volatile int next; } pf_job;

static void* pf_worker(void* arg)
{
    pf_job* job = (pf_job*)arg;
    for (;;) {
        int i = __sync_fetch_and_add(&job->next, 1);
        if (i >= job->count) break;
        job->body(i, job->ctx);
    }
    return NULL;
}

int orc_max_threads(void)
{
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    return (c > 0) ? (int)c : 1;
}

static void parallel_for(int count, int nthreads, pf_body body, void* ctx)
{
    if (nthreads <= 0) nthreads = orc_max_threads();
    if (nthreads > count) nthreads = count;
    pf_job job = { body, ctx, count, 0 };
    if (nthreads <= 1) { pf_worker(&job); return; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    int started = 0;
    for (int t = 0; t < nthreads - 1; t++) if (pthread_create(&th[started], NULL, pf_worker, &job) == 0) started++;
    pf_worker(&job);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    free(th);
}

/* ------------------------------------------------------------------------------------------
 * MT19937 + NumPy legacy transforms (np.random.seed / uniform / choice / standard_normal).
 * Verified bit-exact against numpy 2.3.5 RandomState (tests/test_oracle_golden.py).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t key[624];
    int32_t pos;
    int32_t has_gauss;
    double gauss;
} orc_rng;

void orc_rng_seed(orc_rng* st, uint32_t seed)
{
    for (int i = 0; i < 624; i++) {
        st->key[i] = seed;
        seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u;
    }
    st->pos = 624;
    st->has_gauss = 0;
    st->gauss = 0.0;
}

static void mt_refill(orc_rng* st)
{
    uint32_t* mt = st->key;
    int i;
    for (i = 0; i < 624 - 397; i++) {
        uint32_t y = (mt[i] & 0x80000000u) | (mt[i + 1] & 0x7fffffffu);
        mt[i] = mt[i + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; i < 623; i++) {
        uint32_t y = (mt[i] & 0x80000000u) | (mt[i + 1] & 0x7fffffffu);
        mt[i] = mt[i + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    {
        uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    st->pos = 0;
}

static uint32_t mt_next(orc_rng* st)
{
    if (st->pos >= 624) mt_refill(st);
    uint32_t y = st->key[st->pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

static double mt_double(orc_rng* st)
{
    uint32_t a = mt_next(st) >> 5, b = mt_next(st) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

/* np.random.uniform(lo, hi): lo + (hi - lo) * random_sample() */
double orc_rng_uniform(orc_rng* st, double lo, double hi)
{
    double range = hi - lo;
    return lo + range * mt_double(st);
}

/* np.random.choice(n) == legacy randint(0, n): masked rejection on 32-bit draws; n == 1 draws nothing */
int64_t orc_rng_choice(orc_rng* st, int64_t n)
{
    uint64_t rng = (uint64_t)(n - 1);
    if (rng == 0) return 0;
    uint64_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4;
    mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
    if (rng <= 0xffffffffull) {
        uint32_t v;
        do { v = mt_next(st) & (uint32_t)mask; } while (v > rng);
        return (int64_t)v;
    } else {
        uint64_t v;
        do {
            uint64_t hi = mt_next(st);
            uint64_t lo = mt_next(st);
            v = ((hi << 32) | lo) & mask;
        } while (v > rng);
        return (int64_t)v;
    }
}

/* np.random.standard_normal(): Marsaglia polar with the cached second variate */
double orc_rng_gauss(orc_rng* st)
{
    if (st->has_gauss) {
        double t = st->gauss;
        st->gauss = 0.0;
        st->has_gauss = 0;
        return t;
    }
    double f, x1, x2, r2;
    do {
        x1 = 2.0 * mt_double(st) - 1.0;
        x2 = 2.0 * mt_double(st) - 1.0;
        r2 = x1 * x1 + x2 * x2;
    } while (r2 >= 1.0 || r2 == 0.0);
    f = sqrt(-2.0 * log(r2) / r2);
    st->gauss = f * x1;
    st->has_gauss = 1;
    return f * x2;
}

uint32_t orc_rng_u32(orc_rng* st) { return mt_next(st); }

/* ------------------------------------------------------------------------------------------
 * Problem container: QuadraticFunction (utilities.py:41-46) x (m+1), QCQPForm (utilities.py:122-131)
 * ---------------------------------------------------------------------------------------- */
enum { RELOP_NONE = 0, RELOP_LE = 1, RELOP_EQ = 2 };

enum {
    ORC_OK = 0,
    ORC_ERR_EMPTY_MAX = 1,        /* qcqp.py:117 max() of an empty list (coordinate in no constraint, phase 1) */
    ORC_ERR_UNBOUNDED_UNIFORM = 2 /* utilities.py:267 np.random.uniform with an infinite bound -> OverflowError */
};

typedef struct {
    int n, m;
    /* stacked CSR, form-major: row (j*n + i) is row i of P_j; j = 0 is the objective. Columns sorted. */
    const int64_t* indptr;
    const int32_t* indices;
    const double* data;
    const double* q;      /* dense qarray, [(m+1)*n] */
    const double* r;      /* [m+1] */
    const int32_t* relop; /* [m+1], relop[0] = RELOP_NONE */
    /* coordinate incidence (built here, used by fast mode): forms with a stored entry in row k or q[k] != 0 */
    int64_t* inc_ptr; /* [n+1] */
    int32_t* inc_form;
    /* ADMM data for constraints 1..m (host-computed with the same NumPy calls as utilities.py:160-166) */
    const double* lmb;  /* [m][n] */
    const double* Q;    /* [m][n][n] row-major: Q[i][a][b], eigenvector b in column b */
    const double* qhat; /* [m][n]  Q^T q */
} orc_problem;

orc_problem* orc_problem_create(int n, int m, const int64_t* indptr, const int32_t* indices, const double* data,
                                const double* q, const double* r, const int32_t* relop)
{
    orc_problem* p = (orc_problem*)calloc(1, sizeof(orc_problem));
    p->n = n; p->m = m;
    p->indptr = indptr; p->indices = indices; p->data = data;
    p->q = q; p->r = r; p->relop = relop;
    p->inc_ptr = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
    int64_t* fill = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
    for (int pass = 0; pass < 2; pass++) {
        for (int j = 0; j <= m; j++) {
            for (int k = 0; k < n; k++) {
                int64_t a = indptr[(int64_t)j * n + k], b = indptr[(int64_t)j * n + k + 1];
                int inc = (q[(int64_t)j * n + k] != 0.0);
                for (int64_t e = a; e < b && !inc; e++) inc = (data[e] != 0.0);
                if (!inc) continue;
                if (pass == 0) fill[k]++;
                else p->inc_form[fill[k]++] = j;
            }
        }
        if (pass == 0) {
            for (int k = 0; k < n; k++) p->inc_ptr[k + 1] = p->inc_ptr[k] + fill[k];
            for (int k = 0; k < n; k++) fill[k] = p->inc_ptr[k];
            p->inc_form = (int32_t*)malloc(sizeof(int32_t) * (size_t)(p->inc_ptr[n] > 0 ? p->inc_ptr[n] : 1));
        }
    }
    free(fill);
    p->inc_ptr[0] = 0;
    return p;
}

void orc_problem_set_eig(orc_problem* p, const double* lmb, const double* Q, const double* qhat)
{
    p->lmb = lmb; p->Q = Q; p->qhat = qhat;
}

void orc_problem_destroy(orc_problem* p)
{
    if (!p) return;
    free(p->inc_ptr);
    free(p->inc_form);
    free(p);
}

/* SciPy csr_matvec for one row: sequential, multiply and add rounded separately (SURVEY notes on a-3) */
static double row_dot(const orc_problem* p, int j, int i, const double* x)
{
    int64_t a = p->indptr[(int64_t)j * p->n + i], b = p->indptr[(int64_t)j * p->n + i + 1];
    double s = 0.0;
    for (int64_t e = a; e < b; e++) s = s + p->data[e] * x[p->indices[e]];
    return s;
}

static double diag_entry(const orc_problem* p, int j, int k)
{
    int64_t a = p->indptr[(int64_t)j * p->n + k], b = p->indptr[(int64_t)j * p->n + k + 1];
    for (int64_t e = a; e < b; e++) if (p->indices[e] == k) return p->data[e];
    return 0.0;
}

/* QuadraticFunction.eval (utilities.py:49-50): (P.dot(x) + qarray).dot(x) + r.
 * The outer dot is OpenBLAS ddot in the reference (order not reproducible); here it is sequential. */
double orc_form_eval(const orc_problem* p, int j, const double* x)
{
    double acc = 0.0;
    const double* q = p->q + (int64_t)j * p->n;
    for (int i = 0; i < p->n; i++) {
        double y = row_dot(p, j, i, x) + q[i];
        acc = acc + y * x[i];
    }
    return acc + p->r[j];
}

static double violation_of(int relop, double v)
{
    /* utilities.py:56-62 */
    if (relop == RELOP_EQ) return fabs(v);
    return (v > 0.0) ? v : 0.0;
}

double orc_form_violation(const orc_problem* p, int j, const double* x)
{
    return violation_of(p->relop[j], orc_form_eval(p, j, x));
}

/* max(prob.violations(x)) (utilities.py:133-134) */
double orc_max_violation(const orc_problem* p, const double* x)
{
    double mv = 0.0;
    for (int j = 1; j <= p->m; j++) {
        double v = orc_form_violation(p, j, x);
        if (j == 1 || v > mv) mv = v;
    }
    return mv;
}

/* ------------------------------------------------------------------------------------------
 * One-variable machinery (utilities.py:99-120, 198-288)
 * ---------------------------------------------------------------------------------------- */
typedef struct { double p, q, r; int relop; } onevar_t;

/* OneVarQuadraticFunction.eval (utilities.py:115-120) */
static double onevar_eval(const onevar_t* f, double x)
{
    if (isinf(x)) {
        if (f->p != 0.0) return f->p * x * x;
        if (f->q != 0.0) return f->q * x;
        return f->r; /* the reference has a NameError here (utilities.py:119); unreachable on the path */
    }
    return x * (f->p * x + f->q) + f->r;
}

/* get_onevar_func, faithful (utilities.py:99-105). zbuf: scratch [n] */
static onevar_t get_onevar_faithful(const orc_problem* p, int j, const double* x, int k, double* zbuf)
{
    onevar_t f;
    int n = p->n;
    memcpy(zbuf, x, sizeof(double) * (size_t)n);
    zbuf[k] = 0.0;
    f.p = diag_entry(p, j, k);
    f.q = 2.0 * row_dot(p, j, k, zbuf) + p->q[(int64_t)j * n + k];
    f.r = orc_form_eval(p, j, zbuf);
    f.relop = p->relop[j];
    return f;
}

/* same t2, t1; t0 from the cached f_j(x): t0 = f_j(x) - x_k (t2 x_k + t1) */
static onevar_t get_onevar_cached(const orc_problem* p, int j, double* x, int k, double fval)
{
    onevar_t f;
    int n = p->n;
    double xk = x[k];
    x[k] = 0.0;
    f.p = diag_entry(p, j, k);
    f.q = 2.0 * row_dot(p, j, k, x) + p->q[(int64_t)j * n + k];
    x[k] = xk;
    f.r = fval - xk * (f.p * xk + f.q);
    f.relop = p->relop[j];
    return f;
}

typedef struct { double lo, hi; } ival_t;

static int intervals_le(double p, double q, double r, double s, double tol, ival_t* out)
{
    /* utilities.py:210-231, the '<=' branch: p x^2 + q x + r - s <= 0 */
    if (p > tol) {
        double D = q * q - 4 * p * (r - s);
        if (D >= 0) {
            double rD = sqrt(D);
            out[0].lo = (-q - rD) / (2 * p);
            out[0].hi = (-q + rD) / (2 * p);
            return 1;
        }
        return 0;
    } else if (p < -tol) {
        double D = q * q - 4 * p * (r - s);
        if (D >= 0) {
            double rD = sqrt(D);
            out[0].lo = -INFINITY; out[0].hi = (-q + rD) / (2 * p);
            out[1].lo = (-q - rD) / (2 * p); out[1].hi = INFINITY;
            return 2;
        }
        out[0].lo = -INFINITY; out[0].hi = INFINITY;
        return 1;
    } else {
        if (q > tol) { out[0].lo = -INFINITY; out[0].hi = (s - r) / q; return 1; }
        if (q < -tol) { out[0].lo = (s - r) / q; out[0].hi = INFINITY; return 1; }
        out[0].lo = -INFINITY; out[0].hi = INFINITY;
        return 1;
    }
}

/* get_feasible_intervals(f, s, tol=1e-4) (utilities.py:198-232). out has room for 4. */
int orc_feasible_intervals(double p, double q, double r, int relop, double s, ival_t* out)
{
    const double tol = 1e-4;
    if (relop == RELOP_EQ) {
        ival_t a[2], b[2];
        int na = intervals_le(p, q, r - s, 0.0, tol, a);
        int nb = intervals_le(-p, -q, -r - s, 0.0, tol, b);
        int c = 0;
        for (int i = 0; i < na; i++)
            for (int k = 0; k < nb; k++) {
                double lo = (b[k].lo > a[i].lo) ? b[k].lo : a[i].lo; /* max(I1[0], I2[0]) */
                double hi = (b[k].hi < a[i].hi) ? b[k].hi : a[i].hi; /* min(I1[1], I2[1]) */
                if (lo <= hi) { out[c].lo = lo; out[c].hi = hi; c++; }
            }
        return c;
    }
    return intervals_le(p, q, r, s, tol, out);
}

typedef struct { double key; long delta; } event_t;

static int event_cmp(const void* a, const void* b)
{
    double x = ((const event_t*)a)->key, y = ((const event_t*)b)->key;
    return (x < y) ? -1 : (x > y) ? 1 : 0;
}

typedef struct {
    event_t* ev;
    ival_t* C;
    double* bestxs;
    size_t cap;
} onevar_ws;

static void ws_reserve(onevar_ws* w, size_t m)
{
    size_t need = 4 * m + 8;
    if (need <= w->cap) return;
    free(w->ev); free(w->C); free(w->bestxs);
    w->ev = (event_t*)malloc(sizeof(event_t) * need);
    w->C = (ival_t*)malloc(sizeof(ival_t) * need);
    w->bestxs = (double*)malloc(sizeof(double) * 2 * need);
    w->cap = need;
}

static void ws_free(onevar_ws* w) { free(w->ev); free(w->C); free(w->bestxs); memset(w, 0, sizeof(*w)); }

/* onevar_qcqp(f0, fs, s) (utilities.py:241-288).  Returns 1 and *xout, or 0 for None.
 * *err is set when the reference would raise (np.random.uniform with an infinite bound). */
static int onevar_qcqp(const onevar_t* f0, const onevar_t* fs, long m, double s, orc_rng* rng, onevar_ws* w,
                       double* xout, int* err)
{
    ws_reserve(w, (size_t)m);
    long ne = 0;
    w->ev[ne].key = -INFINITY; w->ev[ne].delta = +1; ne++;
    w->ev[ne].key = INFINITY; w->ev[ne].delta = -1; ne++;
    for (long i = 0; i < m; i++) {
        ival_t I[4];
        int c = orc_feasible_intervals(fs[i].p, fs[i].q, fs[i].r, fs[i].relop, s, I);
        for (int t = 0; t < c; t++) {
            w->ev[ne].key = I[t].lo; w->ev[ne].delta = +1; ne++;
            w->ev[ne].key = I[t].hi; w->ev[ne].delta = -1; ne++;
        }
    }
    qsort(w->ev, (size_t)ne, sizeof(event_t), event_cmp);
    /* dict semantics: merge equal keys, drop zero-net entries (utilities.py:245-250) */
    long nk = 0;
    for (long i = 0; i < ne;) {
        long jx = i, d = 0;
        while (jx < ne && w->ev[jx].key == w->ev[i].key) { d += w->ev[jx].delta; jx++; }
        if (d != 0) { w->ev[nk].key = w->ev[i].key; w->ev[nk].delta = d; nk++; }
        i = jx;
    }
    long nC = 0, tot = 0;
    for (long i = 0; i < nk; i++) {
        tot += w->ev[i].delta;
        if (tot == m && w->ev[i].delta == -1) {
            long prev = (i == 0) ? nk - 1 : i - 1; /* python xs[i-1] */
            w->C[nC].lo = w->ev[prev].key; w->C[nC].hi = w->ev[i].key; nC++;
        }
    }
    if (nC == 0) return 0;

    double p = f0->p, q = f0->q;
    if (p == 0 && q == 0) {
        long idx = (long)orc_rng_choice(rng, nC);
        if (isinf(w->C[idx].lo) || isinf(w->C[idx].hi)) { *err = ORC_ERR_UNBOUNDED_UNIFORM; return 0; }
        *xout = orc_rng_uniform(rng, w->C[idx].lo, w->C[idx].hi);
        return 1;
    }
    double x0 = (p > 0) ? -q / (2. * p) : NAN;
    long nb = 0;
    double bestf = INFINITY;
    for (long i = 0; i < nC; i++) {
        if (w->C[i].lo <= x0 && x0 <= w->C[i].hi) { *xout = x0; return 1; }
        double fl = onevar_eval(f0, w->C[i].lo), fr = onevar_eval(f0, w->C[i].hi);
        if (bestf > fl) { nb = 0; w->bestxs[nb++] = w->C[i].lo; bestf = fl; }
        else if (bestf == fl) w->bestxs[nb++] = w->C[i].lo;
        if (bestf > fr) { nb = 0; w->bestxs[nb++] = w->C[i].hi; bestf = fr; }
        else if (bestf == fr) w->bestxs[nb++] = w->C[i].hi;
    }
    if (nb == 0) return 0;
    *xout = w->bestxs[orc_rng_choice(rng, nb)];
    return 1;
}

/* exported scalar entry points for the golden Q-cases */
int orc_onevar_qcqp(const double* f0, const double* fs /*[m][3]*/, const int32_t* relops, int m, double s,
                    orc_rng* rng, double* xout)
{
    onevar_ws w; memset(&w, 0, sizeof(w));
    onevar_t o = { f0[0], f0[1], f0[2], RELOP_NONE };
    onevar_t* c = (onevar_t*)malloc(sizeof(onevar_t) * (size_t)(m > 0 ? m : 1));
    for (int i = 0; i < m; i++) { c[i].p = fs[3 * i]; c[i].q = fs[3 * i + 1]; c[i].r = fs[3 * i + 2]; c[i].relop = relops[i]; }
    int err = 0;
    int ok = onevar_qcqp(&o, c, m, s, rng, &w, xout, &err);
    free(c); ws_free(&w);
    return err ? -err : ok;
}

int orc_get_feasible_intervals(double p, double q, double r, int relop, double s, double* out /*[8]*/)
{
    ival_t I[4];
    int c = orc_feasible_intervals(p, q, r, relop, s, I);
    for (int i = 0; i < c; i++) { out[2 * i] = I[i].lo; out[2 * i + 1] = I[i].hi; }
    return c;
}

void orc_get_onevar_func(const orc_problem* p, int j, const double* x, int k, double* out3)
{
    double* z = (double*)malloc(sizeof(double) * (size_t)p->n);
    onevar_t f = get_onevar_faithful(p, j, x, k, z);
    out3[0] = f.p; out3[1] = f.q; out3[2] = f.r;
    free(z);
}

/* ------------------------------------------------------------------------------------------
 * Coordinate descent (qcqp.py:101-192)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t num_iters;   /* 1000 */
    double viol_tol;     /* 1e-2 */
    double tol;          /* 1e-4 */
    int32_t phase1;      /* 1 */
    int32_t fast;        /* 0 faithful, 1 cached-f/incidence-list */
    int32_t refresh_every; /* fast mode: recompute the cached f_j every this many phase-2 sweeps (0 = 64) */
} orc_cd_params;

typedef struct {
    int64_t steps_p1;    /* coordinate steps executed in phase 1 */
    int64_t steps_p2;
    int64_t updates_p1;
    int64_t updates_p2;
    int32_t sweeps_p1;   /* outer iterations entered */
    int32_t sweeps_p2;
    int32_t status;      /* ORC_OK or an ORC_ERR_* */
    int32_t ran_phase2;
    int64_t steps_skipped; /* fast mode: phase-1 steps not executed after a no-op sweep (see below) */
} orc_cd_stats;

typedef struct {
    onevar_ws ws;
    onevar_t* nfs;
    double* zbuf;
    double* fval; /* fast mode: cached f_j(x), j = 0..m */
} cd_ctx;

static void ctx_init(cd_ctx* c, const orc_problem* p)
{
    memset(c, 0, sizeof(*c));
    c->nfs = (onevar_t*)malloc(sizeof(onevar_t) * (size_t)(p->m + 1));
    c->zbuf = (double*)malloc(sizeof(double) * (size_t)p->n);
    c->fval = (double*)malloc(sizeof(double) * (size_t)(p->m + 1));
}

static void ctx_free(cd_ctx* c) { ws_free(&c->ws); free(c->nfs); free(c->zbuf); free(c->fval); }

static void refresh_fval(cd_ctx* c, const orc_problem* p, const double* x)
{
    for (int j = 0; j <= p->m; j++) c->fval[j] = orc_form_eval(p, j, x);
}

/* nfs = [f.get_onevar_func(x, i) for f in prob.fs if P != 0 or q != 0]  (qcqp.py:115-116, 164-166) */
static long collect_onevar(cd_ctx* c, const orc_problem* p, double* x, int k, int fast, int want_obj, onevar_t* obj)
{
    long cnt = 0;
    if (!fast) {
        if (want_obj) *obj = get_onevar_faithful(p, 0, x, k, c->zbuf);
        for (int j = 1; j <= p->m; j++) {
            onevar_t f = get_onevar_faithful(p, j, x, k, c->zbuf);
            if (f.p != 0 || f.q != 0) c->nfs[cnt++] = f;
        }
    } else {
        if (want_obj) *obj = get_onevar_cached(p, 0, x, k, c->fval[0]);
        for (int64_t e = p->inc_ptr[k]; e < p->inc_ptr[k + 1]; e++) {
            int j = p->inc_form[e];
            if (j == 0) continue;
            onevar_t f = get_onevar_cached(p, j, x, k, c->fval[j]);
            if (f.p != 0 || f.q != 0) c->nfs[cnt++] = f;
        }
    }
    return cnt;
}

/* fast mode: after x_k: a -> b, f_j(x) = t0 + b (t2 b + t1) for every incident form */
static void apply_move_fast(cd_ctx* c, const orc_problem* p, double* x, int k, double b)
{
    for (int64_t e = p->inc_ptr[k]; e < p->inc_ptr[k + 1]; e++) {
        int j = p->inc_form[e];
        onevar_t f = get_onevar_cached(p, j, x, k, c->fval[j]);
        c->fval[j] = f.r + b * (f.p * b + f.q);
    }
    x[k] = b;
}

static double max_violation_ctx(cd_ctx* c, const orc_problem* p, const double* x, int fast)
{
    if (!fast) return orc_max_violation(p, x);
    double mv = 0.0;
    for (int j = 1; j <= p->m; j++) {
        double v = violation_of(p->relop[j], c->fval[j]);
        if (j == 1 || v > mv) mv = v;
    }
    return mv;
}

/* coord_descent_phase1 (qcqp.py:101-148) */
static void cd_phase1(cd_ctx* c, const orc_problem* p, double* x, const orc_cd_params* prm, orc_rng* rng, orc_cd_stats* st)
{
    const int n = p->n;
    const double tol = prm->tol, viol_tol = prm->viol_tol;
    long update_counter = 0;
    double viol_last = INFINITY;
    onevar_t obj = { 0, 0, 0, RELOP_NONE };
    if (prm->fast) refresh_fval(c, p, x);
    for (int t = 0; t < prm->num_iters; t++) {
        if (viol_last < viol_tol) break;
        st->sweeps_p1++;
        int64_t updates_before = st->updates_p1;
        int broke = 0;
        for (int i = 0; i < n; i++) {
            st->steps_p1++;
            long cnt = collect_onevar(c, p, x, i, prm->fast, 0, NULL);
            if (cnt == 0) { st->status = ORC_ERR_EMPTY_MAX; return; }
            double viol = violation_of(c->nfs[0].relop, onevar_eval(&c->nfs[0], x[i]));
            for (long a = 1; a < cnt; a++) {
                double v = violation_of(c->nfs[a].relop, onevar_eval(&c->nfs[a], x[i]));
                if (v > viol) viol = v;
            }
            double new_xi = x[i], new_viol = viol;
            double ss = -tol, es = viol - viol_tol;
            while (es - ss > tol) {
                double s = (ss + es) / 2;
                double xi; int err = 0;
                int ok = onevar_qcqp(&obj, c->nfs, cnt, s, rng, &c->ws, &xi, &err);
                if (err) { st->status = err; return; }
                if (!ok) ss = s;
                else { new_xi = xi; new_viol = s; es = s; }
            }
            if (new_viol < viol) {
                if (prm->fast) apply_move_fast(c, p, x, i, new_xi); else x[i] = new_xi;
                update_counter = 0;
                st->updates_p1++;
            } else {
                update_counter++;
                if (update_counter == n) { broke = 1; break; } /* 'failed': leaves the inner loop only (qcqp.py:138-141) */
            }
        }
        if (prm->fast) refresh_fval(c, p, x);
        viol_last = max_violation_ctx(c, p, x, prm->fast);
        if (prm->fast && !broke && st->updates_p1 == updates_before && !(viol_last < viol_tol)) {
            /* A full sweep that moved nothing drew no random number either (a feasible probe always moves), and
               update_counter is already past n: every remaining iteration of qcqp.py:110 repeats it exactly.
               Jump to the end of the loop; the steps not executed are reported, not counted as work. */
            st->steps_skipped += (int64_t)(prm->num_iters - (t + 1)) * n;
            break;
        }
    }
}

/* coord_descent_phase2 (qcqp.py:152-178) */
static void cd_phase2(cd_ctx* c, const orc_problem* p, double* x, const orc_cd_params* prm, orc_rng* rng, orc_cd_stats* st)
{
    const int n = p->n;
    if (prm->fast) refresh_fval(c, p, x);
    const double viol = max_violation_ctx(c, p, x, prm->fast);
    long update_counter = 0;
    int converged = 0;
    st->ran_phase2 = 1;
    for (int t = 0; t < prm->num_iters && !converged; t++) {
        st->sweeps_p2++;
        for (int i = 0; i < n; i++) {
            st->steps_p2++;
            onevar_t obj;
            long cnt = collect_onevar(c, p, x, i, prm->fast, 1, &obj);
            double xi; int err = 0;
            int ok = onevar_qcqp(&obj, c->nfs, cnt, viol, rng, &c->ws, &xi, &err);
            if (err) { st->status = err; return; }
            if (ok && fabs(xi - x[i]) > prm->tol) {
                if (prm->fast) apply_move_fast(c, p, x, i, xi); else x[i] = xi;
                update_counter = 0;
                st->updates_p2++;
            } else {
                update_counter++;
                if (update_counter == n) { converged = 1; break; }
            }
        }
        if (prm->fast && !converged) {
            int every = prm->refresh_every > 0 ? prm->refresh_every : 64;
            if ((t + 1) % every == 0) refresh_fval(c, p, x);
        }
    }
}

/* improve_coord_descent (qcqp.py:181-192). x: in/out [n]. */
void orc_improve_cd(const orc_problem* p, const orc_cd_params* prm, double* x, orc_rng* rng, orc_cd_stats* st)
{
    cd_ctx c;
    ctx_init(&c, p);
    memset(st, 0, sizeof(*st));
    if (prm->phase1) cd_phase1(&c, p, x, prm, rng, st);
    if (st->status == ORC_OK) {
        double mv = orc_max_violation(p, x);
        if (mv < prm->viol_tol) cd_phase2(&c, p, x, prm, rng, st);
    }
    ctx_free(&c);
}

/* phase selectors for the goldens: which = 1 or 2 runs coord_descent_phase{1,2} alone */
void orc_cd_phase(const orc_problem* p, const orc_cd_params* prm, int which, double* x, orc_rng* rng, orc_cd_stats* st)
{
    cd_ctx c;
    ctx_init(&c, p);
    memset(st, 0, sizeof(*st));
    if (which == 1) cd_phase1(&c, p, x, prm, rng, st); else cd_phase2(&c, p, x, prm, rng, st);
    ctx_free(&c);
}

/* R independent restarts, each with its own MT19937 stream; threads over restarts. */
typedef struct {
    const orc_problem* p; const orc_cd_params* prm; double* X; orc_rng* rngs; double* f0; double* maxviol; orc_cd_stats* stats;
} cd_batch_ctx;

static void cd_batch_body(int r, void* vctx)
{
    cd_batch_ctx* c = (cd_batch_ctx*)vctx;
    double* x = c->X + (int64_t)r * c->p->n;
    orc_improve_cd(c->p, c->prm, x, &c->rngs[r], &c->stats[r]);
    c->f0[r] = orc_form_eval(c->p, 0, x);
    c->maxviol[r] = orc_max_violation(c->p, x);
}

void orc_improve_cd_batch(const orc_problem* p, const orc_cd_params* prm, int R, double* X /*[R][n] in/out*/,
                          orc_rng* rngs /*[R]*/, double* f0 /*[R]*/, double* maxviol /*[R]*/, orc_cd_stats* stats /*[R]*/,
                          int nthreads)
{
    cd_batch_ctx c = { p, prm, X, rngs, f0, maxviol, stats };
    parallel_for(R, nthreads, cd_batch_body, &c);
}

/* (f0, max violation) of R points (qcqp.py:399-401, 415-417) */
typedef struct { const orc_problem* p; const double* X; double* f0; double* maxviol; double* viol; } eval_batch_ctx;

static void eval_batch_body(int r, void* vctx)
{
    eval_batch_ctx* c = (eval_batch_ctx*)vctx;
    const orc_problem* p = c->p;
    const double* x = c->X + (int64_t)r * p->n;
    c->f0[r] = orc_form_eval(p, 0, x);
    double mv = 0.0;
    for (int j = 1; j <= p->m; j++) {
        double v = orc_form_violation(p, j, x);
        if (c->viol) c->viol[(int64_t)r * p->m + (j - 1)] = v;
        if (j == 1 || v > mv) mv = v;
    }
    c->maxviol[r] = mv;
}

void orc_eval_batch(const orc_problem* p, int R, const double* X, double* f0, double* maxviol, double* viol /*[R][m] or NULL*/,
                    int nthreads)
{
    eval_batch_ctx c = { p, X, f0, maxviol, viol };
    parallel_for(R, nthreads, eval_batch_body, &c);
}

/* QCQPForm.better(x1, x2, tol=1e-4) (utilities.py:135-146): returns 1 if x1 is returned, 2 otherwise */
int orc_better(const orc_problem* p, const double* x1, const double* x2, double tol)
{
    long v1 = (long)(orc_max_violation(p, x1) / tol);
    long v2 = (long)(orc_max_violation(p, x2) / tol);
    double f1 = orc_form_eval(p, 0, x1), f2 = orc_form_eval(p, 0, x2);
    if (v1 < v2) return 1;
    if (v2 < v1) return 2;
    if (f1 < f2) return 1;
    return 2;
}

/* ------------------------------------------------------------------------------------------
 * One-constraint projection and consensus ADMM (utilities.py:149-196, qcqp.py:195-285)
 * ---------------------------------------------------------------------------------------- */
static double phi_of(const double* lmb, const double* qhat, double r, const double* xh, int n)
{
    /* lmb.dot(xhat**2) + qhat.dot(xhat) + r  (utilities.py:174) */
    double a = 0.0, b = 0.0;
    for (int i = 0; i < n; i++) a = a + lmb[i] * (xh[i] * xh[i]);
    for (int i = 0; i < n; i++) b = b + qhat[i] * xh[i];
    return a + b + r;
}

static void xhat_of(double nu, const double* lmb, const double* qhat, const double* zhat, double* xh, int n)
{
    /* -(nu*qhat - 2*zhat) / (2*(1 + nu*lmb))  (utilities.py:173) */
    for (int i = 0; i < n; i++) xh[i] = -((nu * qhat[i] - 2 * zhat[i]) / (2 * (1 + nu * lmb[i])));
}

/* onecons_qcqp(z, f, tol=1e-6) for constraint j (1-based).  out may alias z. Returns bisection iterations. */
int orc_onecons(const orc_problem* p, int j, const double* z, double tol, double* out)
{
    const int n = p->n;
    if (p->relop[j] == RELOP_LE && orc_form_eval(p, j, z) <= 0) {
        if (out != z) memcpy(out, z, sizeof(double) * (size_t)n);
        return 0;
    }
    const double* lmb = p->lmb + (int64_t)(j - 1) * n;
    const double* Q = p->Q + (int64_t)(j - 1) * n * n;
    const double* qhat = p->qhat + (int64_t)(j - 1) * n;
    double* zhat = (double*)malloc(sizeof(double) * (size_t)n * 2);
    double* xh = zhat + n;
    for (int b = 0; b < n; b++) {
        double s = 0.0;
        for (int a = 0; a < n; a++) s = s + Q[(int64_t)a * n + b] * z[a]; /* Q.T.dot(z) */
        zhat[b] = s;
    }
    double r = p->r[j];
    double s = -INFINITY, e = INFINITY;
    for (int i = 0; i < n; i++) {
        double l = lmb[i];
        if (l > 0) { double c = -1. / l; if (c > s) s = c; }
        if (l < 0) { double c = -1. / l; if (c < e) e = c; }
    }
    int it = 0;
    if (s == -INFINITY) {
        s = -1.;
        for (;;) { xhat_of(s, lmb, qhat, zhat, xh, n); if (!(phi_of(lmb, qhat, r, xh, n) <= 0)) break; s *= 2.; if (++it > 4096) break; }
    }
    if (e == INFINITY) {
        e = 1.;
        for (;;) { xhat_of(e, lmb, qhat, zhat, xh, n); if (!(phi_of(lmb, qhat, r, xh, n) >= 0)) break; e *= 2.; if (++it > 4096) break; }
    }
    while (e - s > tol) {
        double mid = (s + e) / 2.;
        xhat_of(mid, lmb, qhat, zhat, xh, n);
        double ph = phi_of(lmb, qhat, r, xh, n);
        it++;
        if (ph > 0) s = mid;
        else if (ph < 0) e = mid;
        else { s = e = mid; break; }
    }
    double nu = (s + e) / 2.;
    xhat_of(nu, lmb, qhat, zhat, xh, n);
    for (int a = 0; a < n; a++) {
        double acc = 0.0;
        for (int b = 0; b < n; b++) acc = acc + Q[(int64_t)a * n + b] * xh[b]; /* Q.dot(xhat) */
        out[a] = acc;
    }
    free(zhat);
    return it;
}

typedef struct {
    int32_t num_iters;  /* 1000 */
    double viol_lim;    /* 1e4 */
    double tol;         /* 1e-2 */
    double rho;
    int32_t phase1;     /* 1 */
} orc_admm_params;

typedef struct {
    int32_t iters_p1;
    int32_t iters_p2;
    int64_t onecons_calls;
    int32_t status;
} orc_admm_stats;

static void copy_better(const orc_problem* p, const double* a, const double* b, double* out)
{
    const double* w = (orc_better(p, a, b, 1e-4) == 1) ? a : b;
    if (out != w) memmove(out, w, sizeof(double) * (size_t)p->n);
}

/* admm_phase1 (qcqp.py:195-212); z: in x0, out z */
static void admm_phase1(const orc_problem* p, double* z, double tol, int num_iters, double* xs, double* us, orc_admm_stats* st)
{
    const int n = p->n, m = p->m;
    for (int i = 0; i < m; i++) { memcpy(xs + (int64_t)i * n, z, sizeof(double) * (size_t)n); memset(us + (int64_t)i * n, 0, sizeof(double) * (size_t)n); }
    double* tmp = (double*)malloc(sizeof(double) * (size_t)n);
    for (int t = 0; t < num_iters; t++) {
        if (orc_max_violation(p, z) < tol) break;
        st->iters_p1++;
        for (int a = 0; a < n; a++) {
            /* (sum(xs) - sum(us)) / m ; python sum() starts from 0 and adds in list order */
            double sx = 0.0, su = 0.0;
            for (int i = 0; i < m; i++) sx = sx + xs[(int64_t)i * n + a];
            for (int i = 0; i < m; i++) su = su + us[(int64_t)i * n + a];
            z[a] = (sx - su) / m;
        }
        for (int i = 0; i < m; i++) {
            for (int a = 0; a < n; a++) tmp[a] = z[a] + us[(int64_t)i * n + a];
            orc_onecons(p, i + 1, tmp, 1e-6, xs + (int64_t)i * n);
            st->onecons_calls++;
        }
        for (int i = 0; i < m; i++)
            for (int a = 0; a < n; a++) us[(int64_t)i * n + a] += z[a] - xs[(int64_t)i * n + a];
    }
    free(tmp);
}

/* admm_phase2 (qcqp.py:215-251).  Zinv = inverse of 2(P0 + rho m I) (the reference factorises it with SuperLU). */
static void admm_phase2(const orc_problem* p, const double* x0, double rho, const double* Zinv, double tol, int num_iters,
                        double viol_lim, double* bestx, double* xs, double* us, orc_admm_stats* st)
{
    const int n = p->n, m = p->m;
    double* z = (double*)malloc(sizeof(double) * (size_t)n * 4);
    double* last_z = z + n; double* rhs = z + 2 * n; double* tmp = z + 3 * n;
    int have_last = 0;
    memcpy(bestx, x0, sizeof(double) * (size_t)n);
    memcpy(z, x0, sizeof(double) * (size_t)n);
    for (int i = 0; i < m; i++) { memcpy(xs + (int64_t)i * n, x0, sizeof(double) * (size_t)n); memset(us + (int64_t)i * n, 0, sizeof(double) * (size_t)n); }
    for (int t = 0; t < num_iters; t++) {
        st->iters_p2++;
        for (int a = 0; a < n; a++) {
            double sx = 0.0, su = 0.0;
            for (int i = 0; i < m; i++) sx = sx + xs[(int64_t)i * n + a];
            for (int i = 0; i < m; i++) su = su + us[(int64_t)i * n + a];
            rhs[a] = 2 * rho * (sx - su) - p->q[a];
        }
        for (int a = 0; a < n; a++) {
            double acc = 0.0;
            for (int b = 0; b < n; b++) acc = acc + Zinv[(int64_t)a * n + b] * rhs[b];
            z[a] = acc;
        }
        for (int i = 0; i < m; i++) {
            for (int a = 0; a < n; a++) tmp[a] = z[a] + us[(int64_t)i * n + a];
            orc_onecons(p, i + 1, tmp, 1e-6, xs + (int64_t)i * n);
            st->onecons_calls++;
        }
        for (int i = 0; i < m; i++)
            for (int a = 0; a < n; a++) us[(int64_t)i * n + a] += z[a] - xs[(int64_t)i * n + a];
        if (have_last) {
            double nrm = 0.0;
            for (int a = 0; a < n; a++) { double d = last_z[a] - z[a]; nrm += d * d; }
            if (sqrt(nrm) < tol) break;
        }
        memcpy(last_z, z, sizeof(double) * (size_t)n); have_last = 1;
        double maxviol = orc_max_violation(p, z);
        if (maxviol > viol_lim) break;
        copy_better(p, z, bestx, bestx);
    }
    free(z);
}

/* improve_admm (qcqp.py:254-285) with rho already validated/chosen by the host. x: in x0, out result. */
void orc_improve_admm(const orc_problem* p, const orc_admm_params* prm, const double* Zinv, double* x, orc_admm_stats* st)
{
    const int n = p->n, m = p->m;
    memset(st, 0, sizeof(*st));
    double* xs = (double*)malloc(sizeof(double) * (size_t)n * (size_t)(m > 0 ? m : 1) * 2);
    double* us = xs + (int64_t)n * m;
    double* x1 = (double*)malloc(sizeof(double) * (size_t)n * 3);
    double* zz = x1 + n; double* x2 = x1 + 2 * n;
    if (prm->phase1) {
        memcpy(zz, x, sizeof(double) * (size_t)n);
        admm_phase1(p, zz, prm->tol, prm->num_iters, xs, us, st);
        copy_better(p, x, zz, x1); /* better(x0, phase1) */
    } else {
        memcpy(x1, x, sizeof(double) * (size_t)n);
    }
    admm_phase2(p, x1, prm->rho, Zinv, prm->tol, prm->num_iters, prm->viol_lim, x2, xs, us, st);
    copy_better(p, x1, x2, x);
    free(xs); free(x1);
}

/* K (rho, Zinv) settings x R starts; out X[K][R][n] */
typedef struct {
    const orc_problem* p; const orc_admm_params* prm; const double* rhos; const double* Zinv; int R; const double* X0;
    double* X; double* f0; double* maxviol; orc_admm_stats* stats;
} admm_batch_ctx;

static void admm_batch_body(int kr, void* vctx)
{
    admm_batch_ctx* c = (admm_batch_ctx*)vctx;
    const int n = c->p->n;
    int k = kr / c->R, r = kr % c->R;
    orc_admm_params q = *c->prm;
    q.rho = c->rhos[k];
    double* x = c->X + (int64_t)kr * n;
    memcpy(x, c->X0 + (int64_t)r * n, sizeof(double) * (size_t)n);
    orc_improve_admm(c->p, &q, c->Zinv + (int64_t)k * n * n, x, &c->stats[kr]);
    c->f0[kr] = orc_form_eval(c->p, 0, x);
    c->maxviol[kr] = orc_max_violation(c->p, x);
}

void orc_improve_admm_batch(const orc_problem* p, const orc_admm_params* prm, int K, const double* rhos, const double* Zinv /*[K][n][n]*/,
                            int R, const double* X0 /*[R][n]*/, double* X /*[K][R][n]*/, double* f0, double* maxviol,
                            orc_admm_stats* stats, int nthreads)
{
    admm_batch_ctx c = { p, prm, rhos, Zinv, R, X0, X, f0, maxviol, stats };
    parallel_for(K * R, nthreads, admm_batch_body, &c);
}

/* ------------------------------------------------------------------------------------------
 * SDR randomized rounding (qcqp.py:394-401): x = mu + z @ F with z = standard_normal(n),
 * F = sqrt(s)[:,None] * Vt from the SVD NumPy's multivariate_normal takes of Sigma (host-side).
 * ---------------------------------------------------------------------------------------- */
void orc_sdr_sample(int n, const double* mu, const double* F /*[n][n]*/, orc_rng* rng, double* z /*[n] out*/, double* x /*[n] out*/)
{
    for (int i = 0; i < n; i++) z[i] = orc_rng_gauss(rng);
    for (int j = 0; j < n; j++) x[j] = 0.0;
    for (int i = 0; i < n; i++) {
        double zi = z[i];
        const double* Fi = F + (int64_t)i * n;
        for (int j = 0; j < n; j++) x[j] = x[j] + zi * Fi[j];
    }
    for (int j = 0; j < n; j++) x[j] = x[j] + mu[j];
}

/* X = mu + Z F for S given normal vectors, then (f0, maxviol) */
typedef struct { const orc_problem* p; const double* mu; const double* F; const double* Z; double* X; double* f0; double* maxviol; } sdr_batch_ctx;

static void sdr_batch_body(int s, void* vctx)
{
    sdr_batch_ctx* c = (sdr_batch_ctx*)vctx;
    const int n = c->p->n;
    double* x = c->X + (int64_t)s * n;
    const double* z = c->Z + (int64_t)s * n;
    for (int j = 0; j < n; j++) x[j] = 0.0;
    for (int i = 0; i < n; i++) {
        double zi = z[i];
        const double* Fi = c->F + (int64_t)i * n;
        for (int j = 0; j < n; j++) x[j] = x[j] + zi * Fi[j];
    }
    for (int j = 0; j < n; j++) x[j] = x[j] + c->mu[j];
    c->f0[s] = orc_form_eval(c->p, 0, x);
    c->maxviol[s] = orc_max_violation(c->p, x);
}

void orc_sdr_sample_eval(const orc_problem* p, const double* mu, const double* F, const double* Z /*[S][n]*/, int S,
                         double* X /*[S][n]*/, double* f0, double* maxviol, int nthreads)
{
    sdr_batch_ctx c = { p, mu, F, Z, X, f0, maxviol };
    parallel_for(S, nthreads, sdr_batch_body, &c);
}
