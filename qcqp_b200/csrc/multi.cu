// multi.cu -- the best pick across GPUs INSIDE the library ("qcqp_best ... across GPUs"; QCQPForm.better,
// utilities.py:135-146): one process per GPU, restarts sharded with no collective on the data path, and ONE NCCL all-gather at
// the end carrying every rank's best (f0, maxviol, global index, x).  A C caller of the ABI can shard without torch.distributed.
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy torch ships is already mapped in a torch process; QCQP_NCCL_LIB
// overrides the name), so the library itself has no link-time dependency on it.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace qcqp {

int best_launch(const double* df0, const double* dmv, int R, double tol, int* dbest, long long* dbucket, double* dbf, cudaStream_t stream);

// the few NCCL entry points used, declared here so that nccl.h is not needed to build
typedef struct { char internal[128]; } nccl_unique_id;
typedef void* nccl_comm_t;
typedef int (*fn_get_unique_id)(nccl_unique_id*);
typedef int (*fn_comm_init_rank)(nccl_comm_t*, int, nccl_unique_id, int);
typedef int (*fn_comm_destroy)(nccl_comm_t);
typedef int (*fn_all_gather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t);
typedef const char* (*fn_error_string)(int);
constexpr int NCCL_FLOAT64 = 8;      // ncclFloat64 / ncclDouble

struct Nccl {
    void* h = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_all_gather all_gather = nullptr;
    fn_error_string error_string = nullptr;
};
static Nccl g_nccl;
static std::once_flag g_nccl_once;

static int nccl_load()
{
    std::call_once(g_nccl_once, [] {
        const char* name = getenv("QCQP_NCCL_LIB");
        g_nccl.h = dlopen(name && name[0] ? name : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!g_nccl.h) return;
        g_nccl.get_unique_id = (fn_get_unique_id)dlsym(g_nccl.h, "ncclGetUniqueId");
        g_nccl.comm_init_rank = (fn_comm_init_rank)dlsym(g_nccl.h, "ncclCommInitRank");
        g_nccl.comm_destroy = (fn_comm_destroy)dlsym(g_nccl.h, "ncclCommDestroy");
        g_nccl.all_gather = (fn_all_gather)dlsym(g_nccl.h, "ncclAllGather");
        g_nccl.error_string = (fn_error_string)dlsym(g_nccl.h, "ncclGetErrorString");
    });
    if (!g_nccl.h || !g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.comm_destroy || !g_nccl.all_gather)
        return fail(QCQP_ERR_NCCL, "libnccl.so.2 could not be loaded (set QCQP_NCCL_LIB to its path): the multi-GPU best pick needs NCCL");
    return QCQP_OK;
}
static int nccl_fail(const char* what, int rc)
{
    return fail(QCQP_ERR_NCCL, std::string(what) + ": " + (g_nccl.error_string ? g_nccl.error_string(rc) : "NCCL error") + " (" + std::to_string(rc) + ")");
}

// send[0] = f0[b], [1] = maxviol[b], [2] = index_offset + b, [3 .. 3 + n) = X[b]   (b = the local best; nothing usable: +inf, +inf, -1)
__global__ void multi_pack_kernel(const double* __restrict__ f0, const double* __restrict__ mv, const double* __restrict__ X, const int* __restrict__ best, int n,
                                  long long offset, double* __restrict__ send)
{
    const int b = *best;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        send[0] = (b >= 0) ? f0[b] : (__builtin_huge_val());
        send[1] = (b >= 0) ? mv[b] : (__builtin_huge_val());
        send[2] = (b >= 0) ? (double)(offset + b) : -1.0;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) send[3 + i] = (b >= 0 && X) ? X[(size_t)b * n + i] : 0.0;
}
// columns 0 and 1 of the gathered table as contiguous arrays for best_kernel (a rank with nothing usable gets NaN f0: never picked)
__global__ void multi_cols_kernel(const double* __restrict__ recv, int N, int stride, double* __restrict__ fcol, double* __restrict__ vcol)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < N) {
        const bool ok = recv[(size_t)r * stride + 2] >= 0.0;
        fcol[r] = ok ? recv[(size_t)r * stride] : nan("");
        vcol[r] = ok ? recv[(size_t)r * stride + 1] : 0.0;
    }
}
// out[0] = f0, [1] = maxviol, [2] = global index, [3] = rank of the winner; x_best = its point
__global__ void multi_unpack_kernel(const double* __restrict__ recv, const int* __restrict__ win, int n, int stride, double* __restrict__ out, double* __restrict__ xbest)
{
    const int w = *win;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        out[0] = (w >= 0) ? recv[(size_t)w * stride] : (__builtin_huge_val());
        out[1] = (w >= 0) ? recv[(size_t)w * stride + 1] : (__builtin_huge_val());
        out[2] = (w >= 0) ? recv[(size_t)w * stride + 2] : -1.0;
        out[3] = (double)w;
    }
    if (xbest && w >= 0)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) xbest[i] = recv[(size_t)w * stride + 3 + i];
}

}  // namespace qcqp

struct qcqp_comm {
    qcqp::nccl_comm_t comm;
    int rank, nranks, device;
    double* buf;          // send [n + 3] | recv [nranks][n + 3] | fcol [nranks] | vcol [nranks] | out [4]
    size_t buf_doubles;
    int* ibuf;            // local best, winner
};

using namespace qcqp;

extern "C" int qcqp_comm_unique_id(void* id128)
{
    if (!id128) return fail(QCQP_ERR_INVALID, "qcqp_comm_unique_id: null argument");
    int rc = nccl_load();
    if (rc != QCQP_OK) return rc;
    nccl_unique_id id;
    const int nr = g_nccl.get_unique_id(&id);
    if (nr != 0) return nccl_fail("ncclGetUniqueId", nr);
    memcpy(id128, id.internal, 128);
    return QCQP_OK;
}

extern "C" int qcqp_comm_create(int32_t rank, int32_t nranks, const void* id128, qcqp_comm** out)
{
    if (!out || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(QCQP_ERR_INVALID, "qcqp_comm_create: bad argument");
    if (qcqp_device_count() <= 0) return fail(QCQP_ERR_NO_DEVICE, "qcqp_comm_create: no CUDA device visible");
    int rc = nccl_load();
    if (rc != QCQP_OK) return rc;
    qcqp_comm* c = new qcqp_comm();
    c->rank = rank; c->nranks = nranks; c->buf = nullptr; c->buf_doubles = 0; c->ibuf = nullptr; c->comm = nullptr;
    QCQP_CUDA_TRY(cudaGetDevice(&c->device));
    nccl_unique_id id;
    memcpy(id.internal, id128, 128);
    const int nr = g_nccl.comm_init_rank(&c->comm, nranks, id, rank);
    if (nr != 0) { delete c; return nccl_fail("ncclCommInitRank", nr); }
    cudaError_t e = cudaMalloc((void**)&c->ibuf, 2 * sizeof(int));
    if (e != cudaSuccess) { g_nccl.comm_destroy(c->comm); delete c; return fail(QCQP_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
    *out = c;
    return QCQP_OK;
}

extern "C" void qcqp_comm_destroy(qcqp_comm* c)
{
    if (!c) return;
    if (c->comm && g_nccl.comm_destroy) g_nccl.comm_destroy(c->comm);
    if (c->buf) cudaFree(c->buf);
    if (c->ibuf) cudaFree(c->ibuf);
    delete c;
}

extern "C" int qcqp_best_multi(qcqp_comm* c, const double* df0, const double* dmaxviol, const double* dX, int32_t R, int32_t n, double tol,
                               int64_t index_offset, int64_t* best_index, int32_t* best_rank, double* best_f0, double* best_maxviol,
                               double* dx_best, void* stream_)
{
    if (!c || R < 0 || n < 0 || !(tol > 0) || (R > 0 && (!df0 || !dmaxviol))) return fail(QCQP_ERR_INVALID, "qcqp_best_multi: bad argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    QCQP_CUDA_TRY(cudaSetDevice(c->device));
    const int N = c->nranks, stride = n + 3;
    const size_t need = (size_t)stride * (1 + N) + 2 * (size_t)N + 4;
    if (need > c->buf_doubles) {
        if (c->buf) cudaFree(c->buf);
        c->buf = nullptr; c->buf_doubles = 0;
        QCQP_CUDA_TRY(cudaMalloc((void**)&c->buf, need * 8));
        c->buf_doubles = need;
    }
    double* send = c->buf; double* recv = send + stride; double* fcol = recv + (size_t)stride * N; double* vcol = fcol + N; double* outd = vcol + N;
    int* dbest = c->ibuf; int* dwin = c->ibuf + 1;
    if (R > 0) {
        int rc = best_launch(df0, dmaxviol, R, tol, dbest, nullptr, nullptr, stream);
        if (rc != QCQP_OK) return rc;
    } else {
        const int minus1 = -1;
        QCQP_CUDA_TRY(cudaMemcpyAsync(dbest, &minus1, sizeof(int), cudaMemcpyHostToDevice, stream));
    }
    const int blocks = n > 0 ? ((n + 255) / 256 < 64 ? (n + 255) / 256 : 64) : 1;
    multi_pack_kernel<<<blocks, 256, 0, stream>>>(df0, dmaxviol, dX, dbest, n, (long long)index_offset, send);
    QCQP_CUDA_TRY(cudaGetLastError());
    const int nr = g_nccl.all_gather(send, recv, (size_t)stride, NCCL_FLOAT64, c->comm, stream);      // the one collective
    if (nr != 0) return nccl_fail("ncclAllGather", nr);
    multi_cols_kernel<<<(N + 127) / 128, 128, 0, stream>>>(recv, N, stride, fcol, vcol);
    QCQP_CUDA_TRY(cudaGetLastError());
    // ranks own increasing index ranges, so "later index wins an exact tie" is "later rank wins" in the table
    int rc = best_launch(fcol, vcol, N, tol, dwin, nullptr, nullptr, stream);
    if (rc != QCQP_OK) return rc;
    multi_unpack_kernel<<<blocks, 256, 0, stream>>>(recv, dwin, n, stride, outd, dx_best);
    QCQP_CUDA_TRY(cudaGetLastError());
    double h[4];
    QCQP_CUDA_TRY(cudaMemcpyAsync(h, outd, sizeof(h), cudaMemcpyDeviceToHost, stream));
    QCQP_CUDA_TRY(cudaStreamSynchronize(stream));
    if (best_f0) *best_f0 = h[0];
    if (best_maxviol) *best_maxviol = h[1];
    if (best_index) *best_index = (int64_t)h[2];
    if (best_rank) *best_rank = (int32_t)h[3];
    return QCQP_OK;
}
