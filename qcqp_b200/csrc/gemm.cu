// gemm.cu -- the dense FP64 contractions of the path as one shared-memory-tiled kernel:
//   * SDR sampler      X = Z F + mu                      (np.random.multivariate_normal's  mean + z @ factor, qcqp.py:396)
//   * dense quadratic forms of a batch   f_j(x_s) = x_s' P_j x_s   as  rowdot(X P_j, X)   (QuadraticFunction.eval, utilities.py:49-50)
// Two kernels: dgemm_mma_kernel (default) runs on the FP64 tensor pipe (mma.sync.m8n8k4.f64; tcgen05 has no f64 kind);
// dgemm_tile_kernel is the FMA-pipe version (64x64 CTA tile, K-tile 16, 4x4 register tile), kept for A/B runs (QCQP_GEMM_FMA=1).
// Row dots are reduced in a fixed order (per-column-block partials, then a fixed-order sum) so results are reproducible.
#include <cstdlib>

#include "common.cuh"

namespace qcqp {

constexpr int GB_M = 64, GB_N = 64, GB_K = 16;

enum { EPI_BIAS_STORE = 0, EPI_ROWDOT = 1 };

// C[M x N] = A[M x K] B[K x N]; A row-major (lda), B row-major (ldb).
//   EPI_BIAS_STORE: C[r][c] = acc + bias[c]           (C row-major, ldc)
//   EPI_ROWDOT:     part[blockIdx.x][r] = sum over this CTA's 64 columns of acc[r][c] * D[r][c]   (D row-major, ldd)
template <int EPI>
__global__ void __launch_bounds__(256) dgemm_tile_kernel(int M, int N, int K, const double* __restrict__ A, int lda,
                                                         const double* __restrict__ B, int ldb, const double* __restrict__ bias,
                                                         double* __restrict__ C, int ldc, const double* __restrict__ D, int ldd,
                                                         double* __restrict__ part)
{
    __shared__ double As[GB_K][GB_M + 1];
    __shared__ double Bs[GB_K][GB_N];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;            // 16 x 16 threads; thread tile rows ty*4.., cols tx*4..
    const int row0 = blockIdx.y * GB_M, col0 = blockIdx.x * GB_N;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.0;

    // loaders: A tile 64 x 16 (each thread 4 elements along k), B tile 16 x 64 (each thread 4 elements along n)
    const int a_r = tid >> 2, a_k = (tid & 3) * 4;      // 64 rows x 4 groups of 4 k
    const int b_k = tid >> 4, b_c = (tid & 15) * 4;     // 16 k x 16 groups of 4 cols
    for (int k0 = 0; k0 < K; k0 += GB_K) {
        {
            const int gr = row0 + a_r;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int gk = k0 + a_k + u;
                As[a_k + u][a_r] = (gr < M && gk < K) ? A[(size_t)gr * lda + gk] : 0.0;
            }
            const int gk = k0 + b_k;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int gc = col0 + b_c + u;
                Bs[b_k][b_c + u] = (gk < K && gc < N) ? B[(size_t)gk * ldb + gc] : 0.0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GB_K; kk++) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    if (EPI == EPI_BIAS_STORE) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int gr = row0 + ty * 4 + i;
            if (gr >= M) continue;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int gc = col0 + tx * 4 + j;
                if (gc < N) C[(size_t)gr * ldc + gc] = acc[i][j] + (bias ? bias[gc] : 0.0);
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int gr = row0 + ty * 4 + i;
            double s = 0.0;
            if (gr < M) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int gc = col0 + tx * 4 + j;
                    if (gc < N) s = fma(acc[i][j], D[(size_t)gr * ldd + gc], s);
                }
            }
            // the 16 threads of a row group are 16 consecutive lanes: fixed-order butterfly
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, 16);
            if (tx == 0 && gr < M) part[(size_t)blockIdx.x * M + gr] = s;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// FP64 tensor-core variant (mma.sync.m8n8k4.f64, "DMMA"): the dense batched contraction of the path on the tensor pipe.
// CTA tile 128 x 64, K-tile 16, 8 warps as 4 (M) x 2 (N), warp tile 32 x 32 = 4 x 4 DMMA tiles (32 accumulators per lane).
// Fragments (PTX ISA): a = A[lane/4][lane%4], b = B[lane%4][lane/4], c0/c1 = C[lane/4][2 (lane%4) + {0,1}].
// Shared tiles are padded so every fragment load is 2 wavefronts: A row stride 20 (= 4 mod 16), B row stride 72 (= 8 mod 16).
// The next K-tile is fetched into registers while the current one is multiplied.  Per k-step of 4 a warp issues 8 LDS for 16
// DMMAs (4096 FMAs) -- the FMA-pipe kernel above needs 8 LDS per 16 FMAs per lane and is shared-memory bound at 20 % of peak.
// ---------------------------------------------------------------------------------------------------------
constexpr int TB_N = 64, TB_K = 16, TA_LD = TB_K + 4, TBB_LD = TB_N + 8;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// WN = warps along N (2: CTA tile 128 x 64, warp tile 32 x 32;  4: CTA tile 64 x 64, warp tile 32 x 16 -- twice the CTAs, for
// problems too small to give every SM two of the large tiles)
template <int EPI, int WN>
__global__ void __launch_bounds__(256) dgemm_mma_kernel(int M, int N, int K, const double* __restrict__ A, int lda,
                                                        const double* __restrict__ B, int ldb, const double* __restrict__ bias,
                                                        double* __restrict__ C, int ldc, const double* __restrict__ D, int ldd,
                                                        double* __restrict__ part)
{
    constexpr int WM = 8 / WN, TBM = WM * 32, WCOLS = TB_N / WN, JT = WCOLS / 8, AK = (TBM * TB_K) / 256;   // AK doubles of A per thread
    __shared__ __align__(16) double As[TBM * TA_LD];
    __shared__ __align__(16) double Bs[TB_K * TBB_LD];
    __shared__ double red[WN][TBM];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tq = lane & 3;
    const int wm = warp / WN, wn = warp % WN;           // warp tile origin: rows wm*32, cols wn*WCOLS
    const int row0 = blockIdx.y * TBM, col0 = blockIdx.x * TB_N;
    double c0[4][JT], c1[4][JT];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < JT; j++) { c0[i][j] = 0.0; c1[i][j] = 0.0; }

    // loaders: A tile TBM x 16 (thread: AK consecutive k of one row), B tile 16 x 64 (thread: k tid/16, 4 consecutive columns)
    constexpr int TPR = TB_K / AK;                       // threads per A row
    const int a_r = tid / TPR, a_k = (tid % TPR) * AK;
    const int b_k = tid >> 4, b_c = (tid & 15) * 4;
    double ra[AK], rb[4];
    // interior tiles with 16-byte aligned rows take 16-byte loads without bounds tests
    const bool vec_ok = ((lda | ldb) & 1) == 0 && ((reinterpret_cast<size_t>(A) | reinterpret_cast<size_t>(B)) & 15) == 0 &&
                        row0 + TBM <= M && col0 + TB_N <= N;
    auto fetch = [&](int k0) {
        if (vec_ok && k0 + TB_K <= K) {
            const double2* ap = reinterpret_cast<const double2*>(A + (size_t)(row0 + a_r) * lda + k0 + a_k);
#pragma unroll
            for (int u = 0; u < AK / 2; u++) { const double2 t = __ldg(ap + u); ra[2 * u] = t.x; ra[2 * u + 1] = t.y; }
            const double2* bp = reinterpret_cast<const double2*>(B + (size_t)(k0 + b_k) * ldb + col0 + b_c);
#pragma unroll
            for (int u = 0; u < 2; u++) { const double2 t = __ldg(bp + u); rb[2 * u] = t.x; rb[2 * u + 1] = t.y; }
            return;
        }
        const int gr = row0 + a_r;
#pragma unroll
        for (int u = 0; u < AK; u++) {
            const int gk = k0 + a_k + u;
            ra[u] = (gr < M && gk < K) ? A[(size_t)gr * lda + gk] : 0.0;
        }
        const int gk = k0 + b_k;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int gc = col0 + b_c + u;
            rb[u] = (gk < K && gc < N) ? B[(size_t)gk * ldb + gc] : 0.0;
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < K; k0 += TB_K) {
#pragma unroll
        for (int u = 0; u < AK; u++) As[a_r * TA_LD + a_k + u] = ra[u];
#pragma unroll
        for (int u = 0; u < 4; u++) Bs[b_k * TBB_LD + b_c + u] = rb[u];
        __syncthreads();
        if (k0 + TB_K < K) fetch(k0 + TB_K);
#pragma unroll
        for (int kk = 0; kk < TB_K; kk += 4) {
            double af[4], bf[JT];
#pragma unroll
            for (int i = 0; i < 4; i++) af[i] = As[(wm * 32 + i * 8 + g) * TA_LD + kk + tq];
#pragma unroll
            for (int j = 0; j < JT; j++) bf[j] = Bs[(kk + tq) * TBB_LD + wn * WCOLS + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < JT; j++) dmma884(c0[i][j], c1[i][j], af[i], bf[j]);
        }
        __syncthreads();
    }
    if (EPI == EPI_BIAS_STORE) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int gr = row0 + wm * 32 + i * 8 + g;
            if (gr >= M) continue;
#pragma unroll
            for (int j = 0; j < JT; j++) {
                const int gc = col0 + wn * WCOLS + j * 8 + 2 * tq;
                if (gc < N) C[(size_t)gr * ldc + gc] = c0[i][j] + (bias ? bias[gc] : 0.0);
                if (gc + 1 < N) C[(size_t)gr * ldc + gc + 1] = c1[i][j] + (bias ? bias[gc + 1] : 0.0);
            }
        }
    } else {
        // row dot with D over this CTA's 64 columns, reduced in a fixed order: lane's columns, the 4 lanes of a row, the N-warps
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int lr = wm * 32 + i * 8 + g, gr = row0 + lr;
            double s = 0.0;
            if (gr < M) {
#pragma unroll
                for (int j = 0; j < JT; j++) {
                    const int gc = col0 + wn * WCOLS + j * 8 + 2 * tq;
                    if (gc < N) s = fma(c0[i][j], D[(size_t)gr * ldd + gc], s);
                    if (gc + 1 < N) s = fma(c1[i][j], D[(size_t)gr * ldd + gc + 1], s);
                }
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (tq == 0) red[wn][lr] = s;
        }
        __syncthreads();
        if (tid < TBM && row0 + tid < M) {
            double s = red[0][tid];
#pragma unroll
            for (int w = 1; w < WN; w++) s += red[w][tid];
            part[(size_t)blockIdx.x * M + row0 + tid] = s;
        }
    }
}

// tile choice: the 128 x 64 tile when it gives every SM at least two CTAs, else 64 x 64 (QCQP_GEMM_TILE=128|64 forces one)
template <int EPI>
static void launch_mma(int M, int N, int K, const double* A, int lda, const double* B, int ldb, const double* bias, double* C, int ldc,
                       const double* D, int ldd, double* part, cudaStream_t stream)
{
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("QCQP_GEMM_TILE"); forced = e ? atoi(e) : 0; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long big = (long long)((N + TB_N - 1) / TB_N) * ((M + 127) / 128);
    const bool small_tile = forced == 64 || (forced != 128 && big < 2LL * sms);
    if (small_tile) {
        dim3 grid((N + TB_N - 1) / TB_N, (M + 63) / 64);
        dgemm_mma_kernel<EPI, 4><<<grid, 256, 0, stream>>>(M, N, K, A, lda, B, ldb, bias, C, ldc, D, ldd, part);
    } else {
        dim3 grid((N + TB_N - 1) / TB_N, (M + 127) / 128);
        dgemm_mma_kernel<EPI, 2><<<grid, 256, 0, stream>>>(M, N, K, A, lda, B, ldb, bias, C, ldc, D, ldd, part);
    }
}

static bool use_mma_gemm()
{
    static int v = -1;
    if (v < 0) v = getenv("QCQP_GEMM_FMA") ? 0 : 1;     // QCQP_GEMM_FMA=1: the FMA-pipe kernel (A/B runs)
    return v != 0;
}

// X = Z F + mu
int gemm_sdr_launch(int S, int n, const double* dZ, const double* dF, const double* dmu, double* dX, cudaStream_t stream)
{
    if (use_mma_gemm()) {
        launch_mma<EPI_BIAS_STORE>(S, n, n, dZ, n, dF, n, dmu, dX, n, nullptr, 0, nullptr, stream);
        QCQP_CUDA_TRY(cudaGetLastError());
        return QCQP_OK;
    }
    dim3 grid((n + GB_N - 1) / GB_N, (S + GB_M - 1) / GB_M);
    dgemm_tile_kernel<EPI_BIAS_STORE><<<grid, 256, 0, stream>>>(S, n, n, dZ, n, dF, n, dmu, dX, n, nullptr, 0, nullptr);
    QCQP_CUDA_TRY(cudaGetLastError());
    return QCQP_OK;
}

// part[cb][s] = sum over column block cb of (X P)[s][c] X[s][c];  returns the number of column blocks
int gemm_quadform_launch(int S, int n, const double* dX, const double* dPj, int ld, double* dpart, cudaStream_t stream)
{
    if (use_mma_gemm()) {
        launch_mma<EPI_ROWDOT>(S, n, n, dX, n, dPj, ld, nullptr, nullptr, 0, dX, n, dpart, stream);
        QCQP_CUDA_TRY(cudaGetLastError());
        return QCQP_OK;
    }
    dim3 grid((n + GB_N - 1) / GB_N, (S + GB_M - 1) / GB_M);
    dgemm_tile_kernel<EPI_ROWDOT><<<grid, 256, 0, stream>>>(S, n, n, dX, n, dPj, ld, nullptr, nullptr, 0, dX, n, dpart);
    QCQP_CUDA_TRY(cudaGetLastError());
    return QCQP_OK;
}

// C = A B, no epilogue extras
int gemm_plain_launch(int M, int N, int K, const double* dA, int lda, const double* dB, int ldb, double* dC, int ldc, cudaStream_t stream)
{
    if (use_mma_gemm()) {
        launch_mma<EPI_BIAS_STORE>(M, N, K, dA, lda, dB, ldb, nullptr, dC, ldc, nullptr, 0, nullptr, stream);
        QCQP_CUDA_TRY(cudaGetLastError());
        return QCQP_OK;
    }
    dim3 grid((N + GB_N - 1) / GB_N, (M + GB_M - 1) / GB_M);
    dgemm_tile_kernel<EPI_BIAS_STORE><<<grid, 256, 0, stream>>>(M, N, K, dA, lda, dB, ldb, nullptr, dC, ldc, nullptr, 0, nullptr);
    QCQP_CUDA_TRY(cudaGetLastError());
    return QCQP_OK;
}

int gemm_col_blocks(int n) { return (n + GB_N - 1) / GB_N; }

}  // namespace qcqp
