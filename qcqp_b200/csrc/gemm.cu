// gemm.cu -- the dense FP64 contractions of the path as one shared-memory-tiled kernel:
//   * SDR sampler      X = Z F + mu                      (np.random.multivariate_normal's  mean + z @ factor, qcqp.py:396)
//   * dense quadratic forms of a batch   f_j(x_s) = x_s' P_j x_s   as  rowdot(X P_j, X)   (QuadraticFunction.eval, utilities.py:49-50)
// 64x64 CTA tile, K-tile 16, 256 threads x (4x4) register tile, FP64 FMA pipe.  tcgen05 has no f64 kind and on B200 the DMMA
// path has the same 40 TFLOP/s as the FMA pipe, so there is no tensor-core variant of this kernel.
// Row dots are reduced in a fixed order (per-column-block partials, then a fixed-order sum) so results are reproducible.
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

// X = Z F + mu
int gemm_sdr_launch(int S, int n, const double* dZ, const double* dF, const double* dmu, double* dX, cudaStream_t stream)
{
    dim3 grid((n + GB_N - 1) / GB_N, (S + GB_M - 1) / GB_M);
    dgemm_tile_kernel<EPI_BIAS_STORE><<<grid, 256, 0, stream>>>(S, n, n, dZ, n, dF, n, dmu, dX, n, nullptr, 0, nullptr);
    QCQP_CUDA_TRY(cudaGetLastError());
    return QCQP_OK;
}

// part[cb][s] = sum over column block cb of (X P)[s][c] X[s][c];  returns the number of column blocks
int gemm_quadform_launch(int S, int n, const double* dX, const double* dPj, int ld, double* dpart, cudaStream_t stream)
{
    dim3 grid((n + GB_N - 1) / GB_N, (S + GB_M - 1) / GB_M);
    dgemm_tile_kernel<EPI_ROWDOT><<<grid, 256, 0, stream>>>(S, n, n, dX, n, dPj, ld, nullptr, nullptr, 0, dX, n, dpart);
    QCQP_CUDA_TRY(cudaGetLastError());
    return QCQP_OK;
}

// C = A B, no epilogue extras
int gemm_plain_launch(int M, int N, int K, const double* dA, int lda, const double* dB, int ldb, double* dC, int ldc, cudaStream_t stream)
{
    dim3 grid((N + GB_N - 1) / GB_N, (M + GB_M - 1) / GB_M);
    dgemm_tile_kernel<EPI_BIAS_STORE><<<grid, 256, 0, stream>>>(M, N, K, dA, lda, dB, ldb, nullptr, dC, ldc, nullptr, 0, nullptr);
    QCQP_CUDA_TRY(cudaGetLastError());
    return QCQP_OK;
}

int gemm_col_blocks(int n) { return (n + GB_N - 1) / GB_N; }

}  // namespace qcqp
