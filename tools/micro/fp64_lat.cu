// Dependent-chain latencies of the FP64 building blocks of the 1-D solver on sm_100a (one warp, one CTA), in cycles per op.
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
template <int OP>
__global__ void k(double* out, double a, double b, long long* cyc)
{
    double x = a + threadIdx.x * 1e-9, y = b;
    __shared__ double sm[64];
    sm[threadIdx.x & 63] = x;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) {
        if (OP == 0) x = __dadd_rn(x, y);
        if (OP == 1) x = __dmul_rn(x, y);
        if (OP == 2) x = __fma_rn(x, y, y);
        if (OP == 3) x = x / y;
        if (OP == 4) x = sqrt(x) + y;
        if (OP == 5) x = (x > y) ? x * 0.5 : y + x;          // compare + select chain
        if (OP == 6) x = __shfl_xor_sync(0xffffffffu, x, 1);
        if (OP == 7) { sm[threadIdx.x & 63] = x; __syncwarp(); x = sm[(threadIdx.x + 1) & 63]; __syncwarp(); }
        if (OP == 8) { __syncthreads(); }
        if (OP == 9) { int v = __reduce_max_sync(0xffffffffu, (int)__double2hiint(x)); x = __hiloint2double(v, __double2loint(x)); }
        if (OP == 10) x = y / x;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    const char* names[] = {"DADD", "DMUL", "DFMA", "x/y", "sqrt+add", "cmp+sel", "SHFL.64", "smem st+ld", "bar.sync(128)", "redux.max(hi)", "y/x"};
    for (int threads : {32, 128, 256, 512, 1024}) {
        printf("threads per CTA = %d\n", threads);
#define RUN(OP) { k<OP><<<1, threads>>>(out, 1.000001, 1.0000003, cyc); k<OP><<<1, threads>>>(out, 1.000001, 1.0000003, cyc); long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("  %-14s %.1f cycles/op\n", names[OP], (double)h / N); }
        RUN(0) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6)
    }
    return 0;
}
