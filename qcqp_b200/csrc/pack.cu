// pack.cu -- builds the HBM layouts of a QCQP (the reference's QCQPForm / QuadraticFunction containers,
// utilities.py:41-46,122-131) and owns the device memory behind the opaque qcqp_pack handle.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace qcqp {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg)
{
    g_last_error = msg;
    return code;
}

int num_sms(int device)
{
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
    return v > 0 ? v : 148;
}
int max_smem_optin(int device)
{
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    return v > 0 ? v : 227 * 1024;
}

int ensure_workspace(qcqp_pack* p, size_t bytes)
{
    if (bytes <= p->ws_bytes) return QCQP_OK;
    if (p->ws) cudaFree(p->ws);
    p->ws = nullptr;
    p->ws_bytes = 0;
    size_t want = bytes + bytes / 4;
    cudaError_t e = cudaMalloc(&p->ws, want);
    if (e != cudaSuccess) return fail(QCQP_ERR_NOMEM, std::string("workspace cudaMalloc: ") + cudaGetErrorString(e));
    p->ws_bytes = want;
    return QCQP_OK;
}

int ensure_workspace2(qcqp_pack* p, size_t bytes)
{
    if (bytes <= p->ws2_bytes) return QCQP_OK;
    if (p->ws2) cudaFree(p->ws2);
    p->ws2 = nullptr;
    p->ws2_bytes = 0;
    size_t want = bytes + bytes / 4;
    cudaError_t e = cudaMalloc(&p->ws2, want);
    if (e != cudaSuccess) return fail(QCQP_ERR_NOMEM, std::string("eval workspace cudaMalloc: ") + cudaGetErrorString(e));
    p->ws2_bytes = want;
    return QCQP_OK;
}

int ensure_io(qcqp_pack* p, size_t bytes)
{
    if (bytes <= p->io_bytes) return QCQP_OK;
    if (p->io) cudaFree(p->io);
    p->io = nullptr;
    p->io_bytes = 0;
    size_t want = bytes + bytes / 4;
    cudaError_t e = cudaMalloc(&p->io, want);
    if (e != cudaSuccess) return fail(QCQP_ERR_NOMEM, std::string("staging cudaMalloc: ") + cudaGetErrorString(e));
    p->io_bytes = want;
    return QCQP_OK;
}

template <class T>
static int upload(qcqp_pack* p, const std::vector<T>& h, const T** dptr)
{
    size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, bytes);
    if (e != cudaSuccess) return fail(QCQP_ERR_NOMEM, std::string("pack cudaMalloc: ") + cudaGetErrorString(e));
    p->allocs.push_back(d);
    p->info.device_bytes += (int64_t)bytes;
    if (!h.empty()) {
        e = cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return fail(QCQP_ERR_CUDA, std::string("pack cudaMemcpy: ") + cudaGetErrorString(e));
    }
    *dptr = (const T*)d;
    return QCQP_OK;
}

}  // namespace qcqp

using namespace qcqp;

extern "C" const char* qcqp_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* qcqp_version(void) { return "qcqp_b200 0.1 (sm_100a)"; }

extern "C" int qcqp_device_count(void)
{
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
    return c;
}

extern "C" void qcqp_pack_destroy(qcqp_pack* p)
{
    if (!p) return;
    for (void* d : p->allocs) cudaFree(d);
    if (p->ws) cudaFree(p->ws);
    if (p->io) cudaFree(p->io);
    if (p->ws2) cudaFree(p->ws2);
    if (p->sdr_mu) cudaFree(p->sdr_mu);
    if (p->sdr_F) cudaFree(p->sdr_F);
    for (int i = 0; i < 2; i++) { if (p->zpre[i]) cudaFree(p->zpre[i]); if (p->zpre_ev[i]) cudaEventDestroy(p->zpre_ev[i]); }
    if (p->zpre_stream) cudaStreamDestroy(p->zpre_stream);
    if (p->h_small) cudaFreeHost(p->h_small);
    if (p->ev_ok) for (int i = 0; i < 6; i++) cudaEventDestroy(p->ev[i]);
    delete p;
}

extern "C" int qcqp_pack_get_info(const qcqp_pack* p, qcqp_pack_info* info)
{
    if (!p || !info) return fail(QCQP_ERR_INVALID, "qcqp_pack_get_info: null argument");
    *info = p->info;
    return QCQP_OK;
}

extern "C" int qcqp_pack_create(const qcqp_pack_desc* d, qcqp_pack** out)
{
    if (!d || !out) return fail(QCQP_ERR_INVALID, "qcqp_pack_create: null argument");
    *out = nullptr;
    const int n = d->n, m = d->m;
    if (n <= 0 || m < 0) return fail(QCQP_ERR_INVALID, "qcqp_pack_create: need n > 0 and m >= 0");
    if ((uint32_t)m >= INC_FORM_MASK) return fail(QCQP_ERR_INVALID, "qcqp_pack_create: too many constraints");
    if (!d->p_ptr || !d->q_ptr || !d->r || !d->relop) return fail(QCQP_ERR_INVALID, "qcqp_pack_create: null array");
    if (qcqp_device_count() <= 0)
        return fail(QCQP_ERR_NO_DEVICE, "qcqp_pack_create: no CUDA device visible; this engine has no CPU fallback");
    const int nf = m + 1;
    if (d->relop[0] != QCQP_RELOP_NONE) return fail(QCQP_ERR_INVALID, "qcqp_pack_create: relop[0] must be NONE (objective)");
    for (int j = 1; j < nf; j++)
        if (d->relop[j] != QCQP_RELOP_LE && d->relop[j] != QCQP_RELOP_EQ)
            return fail(QCQP_ERR_INVALID, "qcqp_pack_create: constraint relop must be LE or EQ");

    // ---- validate entry order, pick dense forms --------------------------------------------------------
    const double fill = (d->dense_min_fill > 0) ? d->dense_min_fill : 0.25;
    std::vector<int> dense_slot(nf, -1), dense_form;
    for (int j = 0; j < nf; j++) {
        int64_t a = d->p_ptr[j], b = d->p_ptr[j + 1];
        if (b < a) return fail(QCQP_ERR_INVALID, "qcqp_pack_create: p_ptr not monotone");
        for (int64_t e = a; e < b; e++) {
            int i = d->p_row[e], c = d->p_col[e];
            if (i < 0 || i >= n || c < 0 || c >= n) return fail(QCQP_ERR_INVALID, "qcqp_pack_create: P index out of range");
            if (e > a) {
                int pi = d->p_row[e - 1], pc = d->p_col[e - 1];
                if (pi > i || (pi == i && pc >= c))
                    return fail(QCQP_ERR_INVALID, "qcqp_pack_create: entries of a form must be sorted by (row, col) without duplicates");
            }
        }
        int64_t qa = d->q_ptr[j], qb = d->q_ptr[j + 1];
        for (int64_t e = qa; e < qb; e++) {
            if (d->q_idx[e] < 0 || d->q_idx[e] >= n) return fail(QCQP_ERR_INVALID, "qcqp_pack_create: q index out of range");
            if (e > qa && d->q_idx[e - 1] >= d->q_idx[e]) return fail(QCQP_ERR_INVALID, "qcqp_pack_create: q indices must be sorted");
        }
        if (n >= 16 && (double)(b - a) >= fill * (double)n * (double)n) {
            dense_slot[j] = (int)dense_form.size();
            dense_form.push_back(j);
        }
    }
    const int nd = (int)dense_form.size();
    const int ld = (n + 1) & ~1;

    // ---- coordinate-major incidence ----------------------------------------------------------------------
    // (j, k) is an incidence when row k of P_j has an entry or q_j[k] != 0: exactly the forms for which the
    // reference's get_onevar_func can return (t2, t1) != (0, 0)  (utilities.py:99-105, qcqp.py:116,166).
    std::vector<int> inc_cnt(n, 0);
    auto for_each_incidence = [&](auto&& fn) {
        for (int j = 0; j < nf; j++) {
            int64_t e = d->p_ptr[j], pe = d->p_ptr[j + 1];
            int64_t qe = d->q_ptr[j], qend = d->q_ptr[j + 1];
            while (e < pe || qe < qend) {
                int kp = (e < pe) ? d->p_row[e] : n;
                int kq = (qe < qend) ? d->q_idx[qe] : n;
                int k = std::min(kp, kq);
                int64_t e0 = e;
                while (e < pe && d->p_row[e] == k) e++;
                double qk = 0.0;
                if (kq == k) { qk = d->q_val[qe]; qe++; }
                fn(j, k, e0, e, qk);
            }
        }
    };
    for_each_incidence([&](int, int k, int64_t, int64_t, double) { inc_cnt[k]++; });
    std::vector<int> inc_ptr(n + 1, 0);
    for (int k = 0; k < n; k++) {
        if ((int64_t)inc_ptr[k] + inc_cnt[k] > 0x7fffffff) return fail(QCQP_ERR_CAPACITY, "qcqp_pack_create: incidence count overflows int32");
        inc_ptr[k + 1] = inc_ptr[k] + inc_cnt[k];
    }
    const int64_t INC = inc_ptr[n];
    std::vector<uint32_t> inc_form((size_t)INC);
    std::vector<double> inc_t2((size_t)INC, 0.0), inc_qk((size_t)INC, 0.0);
    std::vector<int> row_len((size_t)INC, 0);
    std::vector<int> fillp(inc_ptr.begin(), inc_ptr.end() - 1);
    std::vector<int64_t> inc_e0((size_t)INC), inc_e1((size_t)INC);
    for_each_incidence([&](int j, int k, int64_t e0, int64_t e1, double qk) {
        int e = fillp[k]++;
        bool dn = dense_slot[j] >= 0;
        inc_form[e] = (uint32_t)j | ((uint32_t)d->relop[j] << INC_RELOP_SHIFT) | (dn ? INC_DENSE_BIT : 0u);
        inc_qk[e] = qk;
        int off = 0;
        for (int64_t t = e0; t < e1; t++) {
            if (d->p_col[t] == k) inc_t2[e] = d->p_val[t];
            else off++;
        }
        row_len[e] = dn ? 0 : off;
        inc_e0[e] = e0;
        inc_e1[e] = e1;
    });
    std::vector<int> row_ptr((size_t)INC + 1, 0);
    int64_t nnz_off = 0;
    for (int64_t e = 0; e < INC; e++) {
        nnz_off += row_len[e];
        if (nnz_off > 0x7fffffff) return fail(QCQP_ERR_CAPACITY, "qcqp_pack_create: off-diagonal nnz overflows int32");
        row_ptr[e + 1] = (int)nnz_off;
    }
    std::vector<int> row_col((size_t)nnz_off);
    std::vector<double> row_val((size_t)nnz_off);
    for (int k = 0; k < n; k++)
        for (int e = inc_ptr[k]; e < inc_ptr[k + 1]; e++) {
            if (inc_form[e] & INC_DENSE_BIT) continue;
            int w = row_ptr[e];
            for (int64_t t = inc_e0[e]; t < inc_e1[e]; t++)
                if (d->p_col[t] != k) { row_col[w] = d->p_col[t]; row_val[w] = d->p_val[t]; w++; }
        }

    // capacity of the sweep line's hole list: only incidences that can yield two intervals need a slot
    // (get_feasible_intervals, utilities.py:198-232: p < -tol, or '==' with |p| > tol; tol = 1e-4 is fixed there)
    int max_inc = 0, max_inc_small = 0, max_two = 0;
    for (int k = 0; k < n; k++) {
        const int cnt = inc_ptr[k + 1] - inc_ptr[k];
        max_inc = std::max(max_inc, cnt);
        if (cnt <= 1024) max_inc_small = std::max(max_inc_small, cnt);
        int two = 0;
        for (int e = inc_ptr[k]; e < inc_ptr[k + 1]; e++) {
            int rel = (inc_form[e] >> INC_RELOP_SHIFT) & 3;
            double t2 = inc_t2[e];
            if (rel == QCQP_RELOP_NONE) continue;
            if (t2 < -1e-4 || (rel == QCQP_RELOP_EQ && std::fabs(t2) > 1e-4)) two++;
        }
        max_two = std::max(max_two, two);
    }

    // ---- form-major COO (sparse forms only) and dense matrices ----------------------------------------------
    std::vector<long long> f_ptr(nf + 1, 0), q_ptr(nf + 1, 0);
    std::vector<int> f_row, f_col, q_idx;
    std::vector<double> f_val, q_val;
    for (int j = 0; j < nf; j++) {
        if (dense_slot[j] < 0)
            for (int64_t e = d->p_ptr[j]; e < d->p_ptr[j + 1]; e++) {
                f_row.push_back(d->p_row[e]); f_col.push_back(d->p_col[e]); f_val.push_back(d->p_val[e]);
            }
        f_ptr[j + 1] = (long long)f_row.size();
        for (int64_t e = d->q_ptr[j]; e < d->q_ptr[j + 1]; e++) { q_idx.push_back(d->q_idx[e]); q_val.push_back(d->q_val[e]); }
        q_ptr[j + 1] = (long long)q_idx.size();
    }
    // evaluation programs (forms_eval.cuh: eval_sparse_form_seq): rows ascending over the union of P's rows and q's indices; inside a
    // row P's entries by ascending column, then q_i, then the row is closed
    std::vector<long long> ev_ptr(nf + 1, 0);
    std::vector<EvOp> ev_op;
    for (int j = 0; j < nf; j++) {
        if (dense_slot[j] < 0) {
            int64_t e = d->p_ptr[j], pe = d->p_ptr[j + 1], qe = d->q_ptr[j], qend = d->q_ptr[j + 1];
            while (e < pe || qe < qend) {
                const int ip = (e < pe) ? d->p_row[e] : 0x7fffffff, iq = (qe < qend) ? d->q_idx[qe] : 0x7fffffff;
                const int i = ip < iq ? ip : iq;
                while (e < pe && d->p_row[e] == i) { EvOp o; o.val = d->p_val[e]; o.col = d->p_col[e]; o.row = -1; ev_op.push_back(o); e++; }
                if (iq == i) { EvOp o; o.val = d->q_val[qe]; o.col = -1; o.row = -1; ev_op.push_back(o); qe++; }
                ev_op.back().row = i;
            }
        }
        ev_ptr[j + 1] = (long long)ev_op.size();
    }
    std::vector<double> dense_P((size_t)nd * n * ld, 0.0);
    for (int s = 0; s < nd; s++) {
        int j = dense_form[s];
        double* M = dense_P.data() + (size_t)s * n * ld;
        for (int64_t e = d->p_ptr[j]; e < d->p_ptr[j + 1]; e++) M[(size_t)d->p_row[e] * ld + d->p_col[e]] = d->p_val[e];
    }
    std::vector<double> r(d->r, d->r + nf);
    std::vector<int> relop(d->relop, d->relop + nf);

    // ---- upload ------------------------------------------------------------------------------------------------
    qcqp_pack* p = new qcqp_pack();
    std::memset(&p->v, 0, sizeof(p->v));
    std::memset(&p->info, 0, sizeof(p->info));
    p->ws = nullptr; p->ws_bytes = 0; p->io = nullptr; p->io_bytes = 0; p->ws2 = nullptr; p->ws2_bytes = 0; p->has_eig = false; p->lpc_ok = false; p->sdr_mu = nullptr; p->sdr_F = nullptr; p->sdr_ok = false;
    p->ev_ok = false; p->ev_count = 0; p->tmap_state = 0; p->d_ctr = nullptr; p->x_mirror = nullptr; p->x_mirror_done = false;
    for (int i = 0; i < 2; i++) { p->zpre[i] = nullptr; p->zpre_cap[i] = 0; p->zpre_src[i] = nullptr; p->zpre_S[i] = 0; p->zpre_ev[i] = nullptr; }
    p->zpre_stream = nullptr; p->zpre_next = 0; p->zpre_seq[0] = p->zpre_seq[1] = 0; p->zpre_count = 0; p->h_small = nullptr; p->h_small_cap = 0;
    std::memset(&p->lpc, 0, sizeof(p->lpc));
    cudaGetDevice(&p->device);
    p->objective_dense = dense_slot[0] >= 0;
    PackView& v = p->v;
    v.n = n; v.m = m; v.n_dense = nd; v.ld = ld; v.max_inc = max_inc; v.max_inc_small = max_inc_small; v.max_two = max_two;
    int rc = QCQP_OK;
#define UP(vec, field) if (rc == QCQP_OK) rc = upload(p, vec, &v.field)
    UP(inc_ptr, inc_ptr); UP(inc_form, inc_form); UP(inc_t2, inc_t2); UP(inc_qk, inc_qk);
    std::vector<int> inc_rbeg((size_t)INC), inc_rlen((size_t)INC);
    for (int64_t e = 0; e < INC; e++) {
        if (inc_form[e] & INC_DENSE_BIT) { inc_rbeg[e] = dense_slot[inc_form[e] & INC_FORM_MASK]; inc_rlen[e] = -1; }
        else { inc_rbeg[e] = row_ptr[e]; inc_rlen[e] = row_len[e]; }
    }
    UP(inc_rbeg, inc_rbeg); UP(inc_rlen, inc_rlen); UP(row_col, row_col); UP(row_val, row_val);
    UP(f_ptr, f_ptr); UP(f_row, f_row); UP(f_col, f_col); UP(f_val, f_val);
    UP(q_ptr, q_ptr); UP(q_idx, q_idx); UP(q_val, q_val);
    UP(ev_ptr, ev_ptr); UP(ev_op, ev_op);
    UP(r, r); UP(relop, relop); UP(dense_slot, dense_slot); UP(dense_form, dense_form); UP(dense_P, dense_P);
#undef UP
    if (rc == QCQP_OK) {
        std::vector<unsigned long long> zero8(8, 0ull);
        const unsigned long long* dc = nullptr;
        rc = upload(p, zero8, &dc);
        p->d_ctr = const_cast<unsigned long long*>(dc);
    }
    if (rc != QCQP_OK) { qcqp_pack_destroy(p); return rc; }

    // ---- separable view for the lane-per-coordinate kernel (cd_lpc.cu) ----------------------------------------
    {
        bool ok = (m == n);
        std::vector<double> c_p(n, 0.0), c_q(n, 0.0), c_r(n, 0.0), o_diag(n, 0.0), o_q(n, 0.0);
        std::vector<int> c_rel(n, 0), c_seen(n, 0), o_rbeg(n, 0), o_rlen(n, 0);
        std::vector<unsigned char> o_inc(n, 0);
        for (int j = 1; j < nf && ok; j++) {
            int64_t a = d->p_ptr[j], b = d->p_ptr[j + 1], qa = d->q_ptr[j], qb = d->q_ptr[j + 1];
            if (dense_slot[j] >= 0 || b - a > 1 || qb - qa > 1 || (b - a == 0 && qb - qa == 0)) { ok = false; break; }
            int k = (b > a) ? d->p_row[a] : d->q_idx[qa];
            if (b > a && d->p_col[a] != k) { ok = false; break; }
            if (qb > qa && d->q_idx[qa] != k) { ok = false; break; }
            if (c_seen[k]) { ok = false; break; }
            c_seen[k] = 1;
            c_p[k] = (b > a) ? d->p_val[a] : 0.0;
            c_q[k] = (qb > qa) ? d->q_val[qa] : 0.0;
            c_r[k] = d->r[j];
            c_rel[k] = d->relop[j];
            if (c_p[k] == 0.0 && c_q[k] == 0.0) ok = false;
        }
        for (int k = 0; k < n && ok; k++) if (!c_seen[k]) ok = false;
        if (ok) {
            for (int k = 0; k < n; k++) {
                int e = inc_ptr[k];
                if (e < inc_ptr[k + 1] && (inc_form[e] & INC_FORM_MASK) == 0) {
                    o_inc[k] = 1; o_diag[k] = inc_t2[e]; o_q[k] = inc_qk[e];
                    o_rbeg[k] = row_ptr[e]; o_rlen[k] = row_len[e];
                }
            }
            LpcView& L = p->lpc;
            L.o_r = d->r[0];
            L.obj_dense = dense_slot[0] >= 0;
            if (rc == QCQP_OK) rc = upload(p, c_p, &L.c_p);
            if (rc == QCQP_OK) rc = upload(p, c_q, &L.c_q);
            if (rc == QCQP_OK) rc = upload(p, c_r, &L.c_r);
            if (rc == QCQP_OK) rc = upload(p, c_rel, &L.c_rel);
            if (rc == QCQP_OK) rc = upload(p, o_diag, &L.o_diag);
            if (rc == QCQP_OK) rc = upload(p, o_q, &L.o_q);
            if (rc == QCQP_OK) rc = upload(p, o_rbeg, &L.o_rbeg);
            if (rc == QCQP_OK) rc = upload(p, o_rlen, &L.o_rlen);
            if (rc == QCQP_OK) rc = upload(p, o_inc, &L.o_inc);
            {
                // the six per-coordinate constants side by side (48 B per coordinate): one bulk copy stages a pass's worth
                std::vector<double> cst6((size_t)n * 6);
                for (int k = 0; k < n; k++) {
                    double* c = &cst6[(size_t)k * 6];
                    c[0] = c_p[k]; c[1] = c_q[k]; c[2] = c_r[k]; c[3] = (double)c_rel[k]; c[4] = o_diag[k]; c[5] = o_q[k];
                }
                if (rc == QCQP_OK) rc = upload(p, cst6, &L.cst6);
            }
            if (rc != QCQP_OK) { qcqp_pack_destroy(p); return rc; }
        }
        p->lpc_ok = ok;
        p->info.separable = ok ? 1 : 0;
    }

    // ---- info: the algorithmic bytes of one restart-sweep, streaming model of SURVEY.md 8d --------------------
    //   sum over dense forms (8 n^2 + 8 n) + 12 nnz_offdiag(sparse) + 20 INC + 16 (m+1) + 16 n
    p->info.n = n; p->info.m = m; p->info.n_dense = nd; p->info.max_incidence = max_inc;
    p->info.incidences = INC; p->info.nnz_offdiag = nnz_off;
    double dense_all = (double)nd * (8.0 * n * n + 8.0 * n);
    double dense_obj = p->objective_dense ? (8.0 * n * n + 8.0 * n) : 0.0;
    int64_t nnz_off_obj = 0, inc_obj = 0;
    for (int64_t e = 0; e < INC; e++)
        if ((inc_form[e] & INC_FORM_MASK) == 0) { nnz_off_obj += row_len[e]; inc_obj++; }
    p->info.bytes_per_sweep_phase2 = dense_all + 12.0 * nnz_off + 20.0 * INC + 16.0 * (m + 1) + 16.0 * n;
    p->info.bytes_per_sweep_phase1 = (dense_all - dense_obj) + 12.0 * (nnz_off - nnz_off_obj) + 20.0 * (INC - inc_obj) + 16.0 * m + 16.0 * n;
    *out = p;
    return QCQP_OK;
}

extern "C" int qcqp_admm_pack_eig(qcqp_pack* p, const double* lambda, const double* Q, const double* qhat)
{
    if (!p || !lambda || !Q || !qhat) return fail(QCQP_ERR_INVALID, "qcqp_admm_pack_eig: null argument");
    const size_t n = p->v.n, m = p->v.m;
    cudaSetDevice(p->device);
    std::vector<double> hl(lambda, lambda + m * n), hq(Q, Q + m * n * n), hh(qhat, qhat + m * n);
    std::vector<double> ht(m * n * n);
    for (size_t i = 0; i < m; i++)
        for (size_t a = 0; a < n; a++)
            for (size_t b = 0; b < n; b++) ht[(i * n + b) * n + a] = hq[(i * n + a) * n + b];
    int rc = upload(p, hl, &p->v.eig_lambda);
    if (rc == QCQP_OK) rc = upload(p, hq, &p->v.eig_Q);
    if (rc == QCQP_OK) rc = upload(p, ht, &p->v.eig_Qt);
    if (rc == QCQP_OK) rc = upload(p, hh, &p->v.eig_qhat);
    if (rc == QCQP_OK) p->has_eig = true;
    return rc;
}
