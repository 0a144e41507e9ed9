// api.cu -- extern "C" entry points of libqcqp_b200.so (include/qcqp_b200.h).  Host-buffer variants stage through
// device memory and synchronise; _device variants only enqueue on the caller's stream.
#include <cstring>

#include "common.cuh"

namespace qcqp {
int cd_launch(qcqp_pack* p, const qcqp_cd_params* prm, const double* dX0, int R, qcqp_rng_state* drng, double* dX, double* df0,
              double* dmv, qcqp_cd_stats* dstats, cudaStream_t stream);
int eval_launch(qcqp_pack* p, const double* dX, int R, double* df0, double* dmv, double* dviol, cudaStream_t stream);
int best_launch(const double* df0, const double* dmv, int R, double tol, int* dbest, long long* dbucket, double* dbf, cudaStream_t stream);
int admm_launch(qcqp_pack* p, const qcqp_admm_params* prm, const double* drhos, const double* dZinv, int K, const double* dX0, int R,
                double* dX, double* df0, double* dmv, qcqp_admm_stats* dstats, cudaStream_t stream);
int sdr_launch(qcqp_pack* p, const double* dmu, const double* dF, const double* dZ, uint64_t seed, int S, double* dX, double* df0,
               double* dmv, cudaStream_t stream);

// scoped device buffers for the host-buffer entry points
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes)
    {
        cudaError_t e = cudaMalloc(&p, bytes > 0 ? bytes : 1);
        if (e != cudaSuccess) { p = nullptr; return fail(QCQP_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
        return QCQP_OK;
    }
    template <class T> T* as() { return (T*)p; }
};

// carves 256-byte aligned pieces out of the pack's staging arena
struct Arena {
    char* base = nullptr;
    size_t off = 0;
    static size_t pad(size_t b) { return (b + 255) & ~(size_t)255; }
    template <class T> T* take(size_t bytes) { T* r = (T*)(base + off); off += pad(bytes); return r; }
};

static int check_pack(qcqp_pack* p, const char* who)
{
    if (!p) return fail(QCQP_ERR_INVALID, std::string(who) + ": null pack");
    cudaError_t e = cudaSetDevice(p->device);
    if (e != cudaSuccess) return fail(QCQP_ERR_CUDA, std::string(who) + ": cudaSetDevice: " + cudaGetErrorString(e));
    return QCQP_OK;
}
}  // namespace qcqp

// np.random.seed(int): MT19937 init_genrand, pos = 624, no cached gaussian (SURVEY a-7); one thread per stream
__global__ void mt_seed_kernel(const uint32_t* __restrict__ seeds, qcqp_rng_state* __restrict__ out, int R)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    uint32_t s = seeds[r];
    for (int i = 0; i < 624; i++) {
        out[r].key[i] = s;
        s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i + 1u;
    }
    out[r].pos = 624; out[r].has_gauss = 0; out[r].gauss = 0.0;
}

using namespace qcqp;

#define TRY(x) do { int rc__ = (x); if (rc__ != QCQP_OK) return rc__; } while (0)

// ---------------------------------------------------------------------------------------------------------
extern "C" int qcqp_eval_device(qcqp_pack* pack, const double* dX, int32_t R, double* df0, double* dmaxviol, double* dviol, void* stream)
{
    TRY(check_pack(pack, "qcqp_eval_device"));
    if (R < 0 || (R > 0 && (!dX || !df0 || !dmaxviol))) return fail(QCQP_ERR_INVALID, "qcqp_eval_device: bad argument");
    return eval_launch(pack, dX, R, df0, dmaxviol, dviol, (cudaStream_t)stream);
}

extern "C" int qcqp_eval(qcqp_pack* pack, const double* X, int32_t R, double* f0, double* maxviol, double* viol)
{
    TRY(check_pack(pack, "qcqp_eval"));
    if (R < 0 || (R > 0 && (!X || !f0 || !maxviol))) return fail(QCQP_ERR_INVALID, "qcqp_eval: bad argument");
    if (R == 0) return QCQP_OK;
    const size_t n = pack->v.n, m = pack->v.m;
    TRY(ensure_io(pack, Arena::pad(R * n * 8) + 2 * Arena::pad(R * 8) + Arena::pad(viol ? R * m * 8 : 0)));
    Arena ar; ar.base = (char*)pack->io;
    double* dX = ar.take<double>(R * n * 8); double* dF = ar.take<double>(R * 8); double* dM = ar.take<double>(R * 8);
    double* dV = viol ? ar.take<double>(R * m * 8) : nullptr;
    QCQP_CUDA_TRY(cudaMemcpyAsync(dX, X, R * n * 8, cudaMemcpyHostToDevice, 0));
    TRY(eval_launch(pack, dX, R, dF, dM, dV, 0));
    QCQP_CUDA_TRY(cudaMemcpyAsync(f0, dF, R * 8, cudaMemcpyDeviceToHost, 0));
    QCQP_CUDA_TRY(cudaMemcpyAsync(maxviol, dM, R * 8, cudaMemcpyDeviceToHost, 0));
    if (viol) QCQP_CUDA_TRY(cudaMemcpyAsync(viol, dV, R * m * 8, cudaMemcpyDeviceToHost, 0));
    QCQP_CUDA_TRY(cudaStreamSynchronize(0));
    return QCQP_OK;
}

// ---------------------------------------------------------------------------------------------------------
// device-visible alias of a PINNED host array (cudaHostAlloc / cudaHostRegister / torch pin_memory), or null
static double* pinned_alias(const void* host)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
    return (double*)at.devicePointer;
}

static int check_cd_params(const qcqp_cd_params* p)
{
    if (!p) return fail(QCQP_ERR_INVALID, "qcqp_cd_improve: null params");
    if (p->num_iters < 0) return fail(QCQP_ERR_INVALID, "qcqp_cd_improve: num_iters < 0");
    return QCQP_OK;
}

extern "C" int qcqp_cd_improve_device(qcqp_pack* pack, const qcqp_cd_params* params, const double* dX0, int32_t R, qcqp_rng_state* drng,
                                      double* dX, double* df0, double* dmaxviol, qcqp_cd_stats* dstats, void* stream)
{
    TRY(check_pack(pack, "qcqp_cd_improve_device"));
    TRY(check_cd_params(params));
    if (R < 0 || (R > 0 && (!dX0 || !drng || !dX || !df0 || !dmaxviol))) return fail(QCQP_ERR_INVALID, "qcqp_cd_improve_device: bad argument");
    return cd_launch(pack, params, dX0, R, drng, dX, df0, dmaxviol, dstats, (cudaStream_t)stream);
}

extern "C" int qcqp_cd_improve(qcqp_pack* pack, const qcqp_cd_params* params, const double* X0, int32_t R, qcqp_rng_state* rng,
                               double* X, double* f0, double* maxviol, qcqp_cd_stats* stats)
{
    TRY(check_pack(pack, "qcqp_cd_improve"));
    TRY(check_cd_params(params));
    if (R < 0 || (R > 0 && (!X0 || !rng || !X || !f0 || !maxviol))) return fail(QCQP_ERR_INVALID, "qcqp_cd_improve: bad argument");
    if (R == 0) return QCQP_OK;
    const size_t n = pack->v.n;
    TRY(ensure_io(pack, 2 * Arena::pad(R * n * 8) + 2 * Arena::pad(R * 8) + Arena::pad(R * sizeof(qcqp_rng_state)) +
                            Arena::pad(R * sizeof(qcqp_cd_stats))));
    Arena ar; ar.base = (char*)pack->io;
    double* dX0 = ar.take<double>(R * n * 8); double* dX = ar.take<double>(R * n * 8);
    double* dF = ar.take<double>(R * 8); double* dM = ar.take<double>(R * 8);
    qcqp_rng_state* dR = ar.take<qcqp_rng_state>(R * sizeof(qcqp_rng_state));
    qcqp_cd_stats* dS = ar.take<qcqp_cd_stats>(R * sizeof(qcqp_cd_stats));
    QCQP_CUDA_TRY(cudaMemcpyAsync(dX0, X0, R * n * 8, cudaMemcpyHostToDevice, 0));
    QCQP_CUDA_TRY(cudaMemcpyAsync(dR, rng, R * sizeof(qcqp_rng_state), cudaMemcpyHostToDevice, 0));
    pack->x_mirror = pinned_alias(X); pack->x_mirror_done = false;
    const int rc_cd = cd_launch(pack, params, dX0, R, dR, dX, dF, dM, dS, 0);
    const bool delivered = pack->x_mirror_done;
    pack->x_mirror = nullptr; pack->x_mirror_done = false;
    TRY(rc_cd);
    if (!delivered) QCQP_CUDA_TRY(cudaMemcpyAsync(X, dX, R * n * 8, cudaMemcpyDeviceToHost, 0));
    QCQP_CUDA_TRY(cudaMemcpyAsync(f0, dF, R * 8, cudaMemcpyDeviceToHost, 0));
    QCQP_CUDA_TRY(cudaMemcpyAsync(maxviol, dM, R * 8, cudaMemcpyDeviceToHost, 0));
    QCQP_CUDA_TRY(cudaMemcpyAsync(rng, dR, R * sizeof(qcqp_rng_state), cudaMemcpyDeviceToHost, 0));
    if (stats) QCQP_CUDA_TRY(cudaMemcpyAsync(stats, dS, R * sizeof(qcqp_cd_stats), cudaMemcpyDeviceToHost, 0));
    QCQP_CUDA_TRY(cudaStreamSynchronize(0));
    return QCQP_OK;
}

// Device time of the launches behind the last qcqp_cd_improve(_device) call on this pack, from CUDA events recorded on the
// launching stream: ms[0] phase-1 kernel, ms[1] G = X P0 GEMM, ms[2] phase-2 kernel, ms[3] batched (f0, maxviol).
// Only the separable dense-objective path is split into launches; otherwise *count is 0.  Call after synchronising.
extern "C" int qcqp_cd_get_timing(qcqp_pack* pack, double* ms, int32_t* count)
{
    TRY(check_pack(pack, "qcqp_cd_get_timing"));
    if (!ms || !count) return fail(QCQP_ERR_INVALID, "qcqp_cd_get_timing: null argument");
    *count = 0;
    if (!pack->ev_ok || pack->ev_count < 5) return QCQP_OK;
    for (int i = 0; i < 4; i++) {
        float t = 0.f;
        QCQP_CUDA_TRY(cudaEventElapsedTime(&t, pack->ev[i], pack->ev[i + 1]));
        ms[i] = (double)t;
    }
    *count = 4;
    return QCQP_OK;
}

// Sizes every grow-on-demand device buffer of the pack for batches of up to R restarts / draws (and K rho values), so that the
// `_device` entry points that follow never allocate: no implicit device-wide synchronisation from cudaFree / cudaMalloc inside a
// stream of work, and they become legal inside CUDA-graph capture.  (A pack stays single-stream: its workspaces are shared.)
extern "C" int qcqp_pack_reserve(qcqp_pack* pack, int32_t R, int32_t K)
{
    TRY(check_pack(pack, "qcqp_pack_reserve"));
    if (R < 0 || K < 0) return fail(QCQP_ERR_INVALID, "qcqp_pack_reserve: negative size");
    const size_t n = pack->v.n, m = pack->v.m, Rz = (size_t)R, Kz = (size_t)(K > 0 ? K : 1);
    const size_t npad = (n + 1) & ~(size_t)1;
    // coordinate descent: G = X P0 + stats (separable dense path); cached f_j + coefficient scratch (general / CTA-per-restart)
    size_t ws = Rz * npad * 8 + 512 + Rz * sizeof(qcqp_cd_stats) + Rz * n * 40;
    const size_t gen = Rz * (m + 1) * 8 + 512 + (pack->v.max_inc > 1024 ? Rz * (size_t)pack->v.max_inc * 28 : 0);
    if (gen > ws) ws = gen;
    // ADMM (run-per-CTA kernel): xs / us per run
    const size_t adm = Kz * Rz * 2 * m * n * 8 + 4096;
    if (pack->has_eig && adm > ws) ws = adm;
    TRY(ensure_workspace(pack, ws));
    // batched eval / SDR: row-dot partials of the dense forms and device-generated normals
    TRY(ensure_workspace2(pack, Rz * (size_t)(pack->v.n_dense + 1) * ((n + 63) / 64 + 1) * 8 + Rz * n * 8 + 4096));
    // host-buffer entry points: staging arena of the largest call (the SDR -> CD pipeline)
    TRY(ensure_io(pack, 3 * Arena::pad(Rz * n * 8) + 4 * Arena::pad(Rz * 8) + Arena::pad(Rz * sizeof(qcqp_rng_state)) +
                            Arena::pad(Rz * sizeof(qcqp_cd_stats)) + Arena::pad(Rz * 4) + Arena::pad(64) + Arena::pad(n * n * 8) + Arena::pad(n * 8)));
    if (pack->objective_dense && !pack->sdr_mu) {
        QCQP_CUDA_TRY(cudaMalloc((void**)&pack->sdr_mu, n * 8));
        QCQP_CUDA_TRY(cudaMalloc((void**)&pack->sdr_F, n * n * 8));
    }
    return QCQP_OK;
}

// Device counters of the last qcqp_cd_improve* call on this pack (zeroed at its start; kernels that do not count leave zeros):
// out[0] rows of P_0 applied to g (accepted phase-2 moves), out[1] 32 x 32 diagonal blocks fetched, out[2] rows read by
// from-scratch refreshes of g, out[3] bytes requested from L2 by the phase-2 kernel.  Call after synchronising the stream.
extern "C" int qcqp_cd_get_counters(qcqp_pack* pack, uint64_t* out, int32_t* count)
{
    TRY(check_pack(pack, "qcqp_cd_get_counters"));
    if (!out || !count) return fail(QCQP_ERR_INVALID, "qcqp_cd_get_counters: null argument");
    *count = 0;
    if (!pack->d_ctr) return QCQP_OK;
    QCQP_CUDA_TRY(cudaMemcpy(out, pack->d_ctr, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    *count = 4;
    return QCQP_OK;
}

// ---------------------------------------------------------------------------------------------------------
extern "C" int qcqp_admm_improve_device(qcqp_pack* pack, const qcqp_admm_params* params, const double* drhos, const double* dZinv,
                                        int32_t K, const double* dX0, int32_t R, double* dX, double* df0, double* dmaxviol,
                                        qcqp_admm_stats* dstats, void* stream)
{
    TRY(check_pack(pack, "qcqp_admm_improve_device"));
    if (!params || K < 0 || R < 0) return fail(QCQP_ERR_INVALID, "qcqp_admm_improve_device: bad argument");
    if (!pack->has_eig) return fail(QCQP_ERR_INVALID, "qcqp_admm_improve: call qcqp_admm_pack_eig first");
    return admm_launch(pack, params, drhos, dZinv, K, dX0, R, dX, df0, dmaxviol, dstats, (cudaStream_t)stream);
}

extern "C" int qcqp_admm_improve(qcqp_pack* pack, const qcqp_admm_params* params, const double* rhos, const double* Zinv, int32_t K,
                                 const double* X0, int32_t R, double* X, double* f0, double* maxviol, qcqp_admm_stats* stats)
{
    TRY(check_pack(pack, "qcqp_admm_improve"));
    if (!params || K < 0 || R < 0 || !rhos || !Zinv || !X0 || !X || !f0 || !maxviol) return fail(QCQP_ERR_INVALID, "qcqp_admm_improve: bad argument");
    if (!pack->has_eig) return fail(QCQP_ERR_INVALID, "qcqp_admm_improve: call qcqp_admm_pack_eig first");
    if (K == 0 || R == 0) return QCQP_OK;
    const size_t n = pack->v.n, KR = (size_t)K * R;
    DevBuf dRho, dZ, dX0, dX, dF, dM, dS;
    TRY(dRho.alloc(K * 8)); TRY(dZ.alloc(K * n * n * 8)); TRY(dX0.alloc(R * n * 8)); TRY(dX.alloc(KR * n * 8));
    TRY(dF.alloc(KR * 8)); TRY(dM.alloc(KR * 8)); TRY(dS.alloc(KR * sizeof(qcqp_admm_stats)));
    QCQP_CUDA_TRY(cudaMemcpy(dRho.p, rhos, K * 8, cudaMemcpyHostToDevice));
    QCQP_CUDA_TRY(cudaMemcpy(dZ.p, Zinv, K * n * n * 8, cudaMemcpyHostToDevice));
    QCQP_CUDA_TRY(cudaMemcpy(dX0.p, X0, R * n * 8, cudaMemcpyHostToDevice));
    TRY(admm_launch(pack, params, dRho.as<double>(), dZ.as<double>(), K, dX0.as<double>(), R, dX.as<double>(), dF.as<double>(),
                    dM.as<double>(), dS.as<qcqp_admm_stats>(), 0));
    QCQP_CUDA_TRY(cudaStreamSynchronize(0));
    QCQP_CUDA_TRY(cudaMemcpy(X, dX.p, KR * n * 8, cudaMemcpyDeviceToHost));
    QCQP_CUDA_TRY(cudaMemcpy(f0, dF.p, KR * 8, cudaMemcpyDeviceToHost));
    QCQP_CUDA_TRY(cudaMemcpy(maxviol, dM.p, KR * 8, cudaMemcpyDeviceToHost));
    if (stats) QCQP_CUDA_TRY(cudaMemcpy(stats, dS.p, KR * sizeof(qcqp_admm_stats), cudaMemcpyDeviceToHost));
    return QCQP_OK;
}

// ---------------------------------------------------------------------------------------------------------
extern "C" int qcqp_sdr_sample_eval_device(qcqp_pack* pack, const double* dmu, const double* dF, const double* dZ, uint64_t seed, int32_t S,
                                           double* dX, double* df0, double* dmaxviol, void* stream)
{
    TRY(check_pack(pack, "qcqp_sdr_sample_eval_device"));
    if (S < 0 || (S > 0 && (!dmu || !dF || !dX || !df0 || !dmaxviol))) return fail(QCQP_ERR_INVALID, "qcqp_sdr_sample_eval_device: bad argument");
    return sdr_launch(pack, dmu, dF, dZ, seed, S, dX, df0, dmaxviol, (cudaStream_t)stream);
}

extern "C" int qcqp_sdr_sample_eval(qcqp_pack* pack, const double* mu, const double* F, const double* Z, uint64_t seed, int32_t S,
                                    double* X, double* f0, double* maxviol)
{
    TRY(check_pack(pack, "qcqp_sdr_sample_eval"));
    if (S < 0 || (S > 0 && (!mu || !F || !X || !f0 || !maxviol))) return fail(QCQP_ERR_INVALID, "qcqp_sdr_sample_eval: bad argument");
    if (S == 0) return QCQP_OK;
    const size_t n = pack->v.n;
    TRY(ensure_io(pack, Arena::pad(n * 8) + Arena::pad(n * n * 8) + 2 * Arena::pad(S * n * 8) + 2 * Arena::pad(S * 8)));
    Arena ar; ar.base = (char*)pack->io;
    double* dMu = ar.take<double>(n * 8); double* dFm = ar.take<double>(n * n * 8);
    double* dZ = ar.take<double>(S * n * 8); double* dX = ar.take<double>(S * n * 8);
    double* dF0 = ar.take<double>(S * 8); double* dM = ar.take<double>(S * 8);
    if (Z) QCQP_CUDA_TRY(cudaMemcpyAsync(dZ, Z, S * n * 8, cudaMemcpyHostToDevice, 0));
    QCQP_CUDA_TRY(cudaMemcpyAsync(dMu, mu, n * 8, cudaMemcpyHostToDevice, 0));
    QCQP_CUDA_TRY(cudaMemcpyAsync(dFm, F, n * n * 8, cudaMemcpyHostToDevice, 0));
    TRY(sdr_launch(pack, dMu, dFm, Z ? dZ : nullptr, seed, S, dX, dF0, dM, 0));
    QCQP_CUDA_TRY(cudaMemcpyAsync(X, dX, S * n * 8, cudaMemcpyDeviceToHost, 0));
    QCQP_CUDA_TRY(cudaMemcpyAsync(f0, dF0, S * 8, cudaMemcpyDeviceToHost, 0));
    QCQP_CUDA_TRY(cudaMemcpyAsync(maxviol, dM, S * 8, cudaMemcpyDeviceToHost, 0));
    QCQP_CUDA_TRY(cudaStreamSynchronize(0));
    return QCQP_OK;
}

// ---------------------------------------------------------------------------------------------------------
extern "C" int qcqp_sdr_cd_pipeline(qcqp_pack* pack, const qcqp_cd_params* params, const double* mu, const double* F, const double* Z,
                                    uint64_t seed, int32_t S, const uint32_t* seeds, double* X0, double* f0_draw, double* maxviol_draw,
                                    double* X, double* f0, double* maxviol, qcqp_cd_stats* stats, qcqp_rng_state* rng_out, int32_t* best_idx)
{
    TRY(check_pack(pack, "qcqp_sdr_cd_pipeline"));
    TRY(check_cd_params(params));
    if (S < 0 || (S > 0 && (!seeds || !X || !f0 || !maxviol))) return fail(QCQP_ERR_INVALID, "qcqp_sdr_cd_pipeline: bad argument");
    if ((mu == nullptr) != (F == nullptr)) return fail(QCQP_ERR_INVALID, "qcqp_sdr_cd_pipeline: mu and F must be given together");
    if (!mu && !pack->sdr_ok) return fail(QCQP_ERR_INVALID, "qcqp_sdr_cd_pipeline: no cached SDR factor; pass mu and F once");
    if (S == 0) return QCQP_OK;
    const size_t n = pack->v.n;
    if (mu) {
        if (!pack->sdr_mu) {
            QCQP_CUDA_TRY(cudaMalloc((void**)&pack->sdr_mu, n * 8));
            QCQP_CUDA_TRY(cudaMalloc((void**)&pack->sdr_F, n * n * 8));
        }
        QCQP_CUDA_TRY(cudaMemcpyAsync(pack->sdr_mu, mu, n * 8, cudaMemcpyHostToDevice, 0));
        QCQP_CUDA_TRY(cudaMemcpyAsync(pack->sdr_F, F, n * n * 8, cudaMemcpyHostToDevice, 0));
        pack->sdr_ok = true;
    }
    TRY(ensure_io(pack, 3 * Arena::pad(S * n * 8) + 4 * Arena::pad(S * 8) + Arena::pad(S * sizeof(qcqp_rng_state)) +
                            Arena::pad(S * sizeof(qcqp_cd_stats)) + Arena::pad(S * 4) + Arena::pad(64)));
    Arena ar; ar.base = (char*)pack->io;
    double* dZ = ar.take<double>(S * n * 8); double* dX0 = ar.take<double>(S * n * 8); double* dX = ar.take<double>(S * n * 8);
    double* dFs = ar.take<double>(S * 8); double* dMs = ar.take<double>(S * 8);
    qcqp_rng_state* dR = ar.take<qcqp_rng_state>(S * sizeof(qcqp_rng_state));
    uint32_t* dSeeds = ar.take<uint32_t>(S * 4);
    // the small results lie back to back: ONE read-back into the pack's pinned staging block, then host copies (instead of four
    // staged copies into the caller's possibly pageable arrays)
    const size_t small0 = ar.off;
    double* dF0 = ar.take<double>(S * 8); double* dM = ar.take<double>(S * 8);
    qcqp_cd_stats* dS = ar.take<qcqp_cd_stats>(S * sizeof(qcqp_cd_stats));
    int* dBest = ar.take<int>(64);
    const size_t small_bytes = ar.off - small0;
    if (pack->h_small_cap < small_bytes) {
        if (pack->h_small) QCQP_CUDA_TRY(cudaFreeHost(pack->h_small));
        pack->h_small = nullptr; pack->h_small_cap = 0;
        QCQP_CUDA_TRY(cudaHostAlloc(&pack->h_small, small_bytes, cudaHostAllocDefault));
        pack->h_small_cap = small_bytes;
    }
    // standard normals: already on the device if qcqp_sdr_prefetch was given this very array (its upload ran beside the previous call)
    const double* dZuse = dZ;
    bool prefetched = false;
    if (Z) {
        int pick = -1;
        for (int j = 0; j < 2; j++)
            if (pack->zpre_src[j] == Z && pack->zpre_S[j] == S && (pick < 0 || pack->zpre_seq[j] < pack->zpre_seq[pick])) pick = j;
        if (pick >= 0) {
            QCQP_CUDA_TRY(cudaStreamWaitEvent(0, pack->zpre_ev[pick], 0));
            dZuse = pack->zpre[pick]; pack->zpre_src[pick] = nullptr; prefetched = true;
        }
    }
    if (Z && !prefetched) QCQP_CUDA_TRY(cudaMemcpyAsync(dZ, Z, S * n * 8, cudaMemcpyHostToDevice, 0));
    QCQP_CUDA_TRY(cudaMemcpyAsync(dSeeds, seeds, S * 4, cudaMemcpyHostToDevice, 0));
    mt_seed_kernel<<<(S + 127) / 128, 128, 0, 0>>>(dSeeds, dR, S);
    QCQP_CUDA_TRY(cudaGetLastError());
    TRY(sdr_launch(pack, pack->sdr_mu, pack->sdr_F, Z ? dZuse : nullptr, seed, S, dX0, dFs, dMs, 0));
    pack->x_mirror = pinned_alias(X); pack->x_mirror_done = false;
    const int rc_cd = cd_launch(pack, params, dX0, S, dR, dX, dF0, dM, dS, 0);
    const bool delivered = pack->x_mirror_done;
    pack->x_mirror = nullptr; pack->x_mirror_done = false;
    TRY(rc_cd);
    if (best_idx) TRY(best_launch(dF0, dM, S, 1e-4, dBest, nullptr, nullptr, 0));
    if (!delivered) QCQP_CUDA_TRY(cudaMemcpyAsync(X, dX, S * n * 8, cudaMemcpyDeviceToHost, 0));
    QCQP_CUDA_TRY(cudaMemcpyAsync(pack->h_small, (char*)pack->io + small0, small_bytes, cudaMemcpyDeviceToHost, 0));
    if (X0) QCQP_CUDA_TRY(cudaMemcpyAsync(X0, dX0, S * n * 8, cudaMemcpyDeviceToHost, 0));
    if (f0_draw) QCQP_CUDA_TRY(cudaMemcpyAsync(f0_draw, dFs, S * 8, cudaMemcpyDeviceToHost, 0));
    if (maxviol_draw) QCQP_CUDA_TRY(cudaMemcpyAsync(maxviol_draw, dMs, S * 8, cudaMemcpyDeviceToHost, 0));
    if (rng_out) QCQP_CUDA_TRY(cudaMemcpyAsync(rng_out, dR, S * sizeof(qcqp_rng_state), cudaMemcpyDeviceToHost, 0));
    QCQP_CUDA_TRY(cudaStreamSynchronize(0));
    const char* hs = (const char*)pack->h_small;
    memcpy(f0, hs + ((char*)dF0 - ((char*)pack->io + small0)), S * 8);
    memcpy(maxviol, hs + ((char*)dM - ((char*)pack->io + small0)), S * 8);
    if (stats) memcpy(stats, hs + ((char*)dS - ((char*)pack->io + small0)), S * sizeof(qcqp_cd_stats));
    if (best_idx) memcpy(best_idx, hs + ((char*)dBest - ((char*)pack->io + small0)), 4);
    return QCQP_OK;
}

// Upload of the standard normals of a LATER qcqp_sdr_cd_pipeline call, asynchronous, on a private stream: a caller that runs batch
// after batch issues it before the call on the current batch, and the copy hides behind that call's kernels.
extern "C" int qcqp_sdr_prefetch(qcqp_pack* pack, const double* Z, int32_t S)
{
    TRY(check_pack(pack, "qcqp_sdr_prefetch"));
    if (!Z || S <= 0) return fail(QCQP_ERR_INVALID, "qcqp_sdr_prefetch: bad argument");
    const size_t bytes = (size_t)S * pack->v.n * 8;
    // an empty slot if there is one, else the older of the two
    int i = !pack->zpre_src[0] ? 0 : (!pack->zpre_src[1] ? 1 : (pack->zpre_seq[0] < pack->zpre_seq[1] ? 0 : 1));
    if (!pack->zpre_stream) QCQP_CUDA_TRY(cudaStreamCreateWithFlags(&pack->zpre_stream, cudaStreamNonBlocking));
    if (!pack->zpre_ev[i]) QCQP_CUDA_TRY(cudaEventCreateWithFlags(&pack->zpre_ev[i], cudaEventDisableTiming));
    if (pack->zpre_cap[i] < bytes) {
        if (pack->zpre[i]) QCQP_CUDA_TRY(cudaFree(pack->zpre[i]));
        pack->zpre[i] = nullptr; pack->zpre_cap[i] = 0; pack->zpre_src[i] = nullptr;
        QCQP_CUDA_TRY(cudaMalloc((void**)&pack->zpre[i], bytes));
        pack->zpre_cap[i] = bytes;
    }
    QCQP_CUDA_TRY(cudaMemcpyAsync(pack->zpre[i], Z, bytes, cudaMemcpyHostToDevice, pack->zpre_stream));
    QCQP_CUDA_TRY(cudaEventRecord(pack->zpre_ev[i], pack->zpre_stream));
    pack->zpre_src[i] = Z; pack->zpre_S[i] = S; pack->zpre_seq[i] = ++pack->zpre_count;
    return QCQP_OK;
}

// ---------------------------------------------------------------------------------------------------------
extern "C" int qcqp_best_device(const double* df0, const double* dmaxviol, int32_t R, double tol, int32_t* dbest_idx, int64_t* dbest_bucket,
                                double* dbest_f0, void* stream)
{
    if (R <= 0 || !df0 || !dmaxviol || !dbest_idx || !(tol > 0)) return fail(QCQP_ERR_INVALID, "qcqp_best_device: bad argument");
    return best_launch(df0, dmaxviol, R, tol, dbest_idx, (long long*)dbest_bucket, dbest_f0, (cudaStream_t)stream);
}

extern "C" int qcqp_best(const double* f0, const double* maxviol, int32_t R, double tol, int32_t* best_idx)
{
    if (R <= 0 || !f0 || !maxviol || !best_idx || !(tol > 0)) return fail(QCQP_ERR_INVALID, "qcqp_best: bad argument");
    if (qcqp_device_count() <= 0) return fail(QCQP_ERR_NO_DEVICE, "qcqp_best: no CUDA device visible; this engine has no CPU fallback");
    DevBuf dF, dM, dI;
    TRY(dF.alloc((size_t)R * 8)); TRY(dM.alloc((size_t)R * 8)); TRY(dI.alloc(4));
    QCQP_CUDA_TRY(cudaMemcpy(dF.p, f0, (size_t)R * 8, cudaMemcpyHostToDevice));
    QCQP_CUDA_TRY(cudaMemcpy(dM.p, maxviol, (size_t)R * 8, cudaMemcpyHostToDevice));
    TRY(best_launch(dF.as<double>(), dM.as<double>(), R, tol, dI.as<int>(), nullptr, nullptr, 0));
    QCQP_CUDA_TRY(cudaStreamSynchronize(0));
    QCQP_CUDA_TRY(cudaMemcpy(best_idx, dI.p, 4, cudaMemcpyDeviceToHost));
    return QCQP_OK;
}
