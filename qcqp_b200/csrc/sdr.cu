// sdr.cu -- SDR randomized rounding (the sampler lines of QCQP.suggest, qcqp.py:394-401), batched over S draws:
//   x_s = mu + z_s F     (F = sqrt(s)[:,None] * Vt: exactly the factor np.random.multivariate_normal builds from Sigma)
//   f0(x_s), max_i violation_i(x_s)
// z_s comes from the caller (NumPy standard_normal stream: parity mode) or from a device Philox4x32-10 + Box-Muller
// stream (throughput mode).  One warp per draw; lanes own columns of F so rows of F stream coalesced.
#include "common.cuh"
#include "forms_eval.cuh"
#include "onevar.cuh"

namespace qcqp {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// two standard normals from one Philox block (Box-Muller on 53-bit uniforms in (0, 1])
__device__ __forceinline__ void philox_normal2(uint64_t seed, uint64_t sample, uint32_t pair, double* a, double* b)
{
    uint32_t o[4];
    philox4x32_10(pair, (uint32_t)sample, (uint32_t)(sample >> 32), 0x5d2u, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    double u1 = ((double)(((uint64_t)(o[0] >> 5) << 26) | (o[1] >> 6)) + 1.0) / 9007199254740992.0;
    double u2 = ((double)(((uint64_t)(o[2] >> 5) << 26) | (o[3] >> 6))) / 9007199254740992.0;
    double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    *a = rad * c;
    *b = rad * s;
}

__global__ void sdr_kernel(PackView P, const double* __restrict__ mu, const double* __restrict__ F, const double* __restrict__ Z,
                           uint64_t seed, int S, double* __restrict__ X, double* __restrict__ f0, double* __restrict__ maxviol)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int n = P.n, m = P.m;
    const int npad = (n + 1) & ~1;
    double* z = reinterpret_cast<double*>(smem) + (size_t)warp * 2 * npad;
    double* x = z + npad;
    for (int s = blockIdx.x * wpb + warp; s < S; s += gridDim.x * wpb) {
        __syncwarp();
        if (Z) {
            for (int i = lane; i < n; i += 32) z[i] = Z[(size_t)s * n + i];
        } else {
            for (int pr = lane; 2 * pr < n; pr += 32) {
                double a, b;
                philox_normal2(seed, (uint64_t)s, (uint32_t)pr, &a, &b);
                z[2 * pr] = a;
                if (2 * pr + 1 < n) z[2 * pr + 1] = b;
            }
        }
        __syncwarp();
        // x = z @ F + mu : lanes own columns j, rows of F are read coalesced
        for (int j0 = 0; j0 < n; j0 += 128) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            const int j = j0 + lane;
            const bool v0 = j < n, v1 = j + 32 < n, v2 = j + 64 < n, v3 = j + 96 < n;
            for (int i = 0; i < n; i++) {
                const double zi = z[i];
                const double* Fi = F + (size_t)i * n;
                if (v0) a0 = fma(zi, Fi[j], a0);
                if (v1) a1 = fma(zi, Fi[j + 32], a1);
                if (v2) a2 = fma(zi, Fi[j + 64], a2);
                if (v3) a3 = fma(zi, Fi[j + 96], a3);
            }
            if (v0) x[j] = a0 + mu[j];
            if (v1) x[j + 32] = a1 + mu[j + 32];
            if (v2) x[j + 64] = a2 + mu[j + 64];
            if (v3) x[j + 96] = a3 + mu[j + 96];
        }
        __syncwarp();
        for (int i = lane; i < n; i += 32) X[(size_t)s * n + i] = x[i];
        double mv = -QCQP_INF, fobj = 0.0;
        const PackView& Pr = P;
        eval_forms(P, x, 0, m, false, lane, [&](int j, double v) {
            if (j == 0) fobj = v;
            else {
                double vv = violation_of(Pr.relop[j], v);
                mv = (vv > mv) ? vv : mv;
            }
        });
        mv = warp_max(mv);
        fobj = warp_sum(fobj);
        if (lane == 0) { f0[s] = fobj; maxviol[s] = (m > 0) ? mv : 0.0; }
    }
}

int gemm_sdr_launch(int S, int n, const double* dZ, const double* dF, const double* dmu, double* dX, cudaStream_t stream);
int eval_launch(qcqp_pack* p, const double* dX, int R, double* df0, double* dmv, double* dviol, cudaStream_t stream);

// standard normals for the batched path when the caller supplies none: the same Philox stream as sdr_kernel
__global__ void philox_normals_kernel(uint64_t seed, int S, int n, double* __restrict__ Z)
{
    const int pairs = (n + 1) / 2;
    const long long total = (long long)S * pairs;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int s = (int)(i / pairs), pr = (int)(i % pairs);
        double a, b;
        philox_normal2(seed, (uint64_t)s, (uint32_t)pr, &a, &b);
        Z[(size_t)s * n + 2 * pr] = a;
        if (2 * pr + 1 < n) Z[(size_t)s * n + 2 * pr + 1] = b;
    }
}

int sdr_launch(qcqp_pack* p, const double* dmu, const double* dF, const double* dZ, uint64_t seed, int S, double* dX, double* df0,
               double* dmv, cudaStream_t stream)
{
    if (S <= 0) return QCQP_OK;
    const int n = p->v.n;
    if (S >= 32) {
        // batched path: X = Z F + mu as a tiled FP64 GEMM (gemm.cu), then the batched evaluation
        if (!dZ) {
            // dX doubles as the buffer of the normals? No: the GEMM reads Z while writing X.  Stage Z behind the eval scratch.
            double* zbuf = nullptr;
            cudaError_t e = cudaMallocAsync((void**)&zbuf, (size_t)S * n * 8, stream);
            if (e != cudaSuccess) return fail(QCQP_ERR_NOMEM, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
            philox_normals_kernel<<<296, 256, 0, stream>>>(seed, S, n, zbuf);
            int rc = gemm_sdr_launch(S, n, zbuf, dF, dmu, dX, stream);
            cudaFreeAsync(zbuf, stream);
            if (rc != QCQP_OK) return rc;
        } else {
            int rc = gemm_sdr_launch(S, n, dZ, dF, dmu, dX, stream);
            if (rc != QCQP_OK) return rc;
        }
        return eval_launch(p, dX, S, df0, dmv, nullptr, stream);
    }
    const int wpb = 4;
    size_t smem = (size_t)wpb * 2 * ((n + 1) & ~1) * 8;
    if (smem > (size_t)max_smem_optin(p->device)) return fail(QCQP_ERR_CAPACITY, "qcqp_sdr_sample_eval: n too large for shared-memory staging");
    QCQP_CUDA_TRY(cudaFuncSetAttribute(sdr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = (S + wpb - 1) / wpb;
    sdr_kernel<<<blocks, wpb * 32, smem, stream>>>(p->v, dmu, dF, dZ, seed, S, dX, df0, dmv);
    QCQP_CUDA_TRY(cudaGetLastError());
    return QCQP_OK;
}

}  // namespace qcqp
