// fp64lat.cu -- dependent-issue latency (one warp) and throughput (W warps per SM sub-partition) of the FP64 instructions the
// coordinate-descent kernels are made of, on the GPU at hand: DADD, DMUL, DFMA, DSETP+FSEL (a running max), sqrt (IEEE), division (IEEE),
// LDS.128 + compare.  Diagnostic only (tools/README.md); build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64lat fp64lat.cu
#include <cstdio>
#include <cuda_runtime.h>

#define N 4096
template <int OP>
__global__ void chain(double* out, long long* cyc, double a, double b)
{
    __shared__ double2 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_double2(a + threadIdx.x, b - threadIdx.x);
    __syncthreads();
    double x = a + threadIdx.x * 1e-9, y = b, m = -1e300;
    const long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (OP == 0) x = x + y;
        if (OP == 1) x = x * y;
        if (OP == 2) x = fma(x, y, y);
        if (OP == 3) { if (x > m) m = x; x = m + y; }                 // DSETP -> FSEL x2 -> DADD
        if (OP == 4) x = sqrt(x) + y;
        if (OP == 5) x = y / x + y;
        if (OP == 6) { const double2 u = sm[(i + (int)x) & 63]; if (u.x < y && u.y > m) m = u.y; x = m * 1e-300; }
        if (OP == 7) { const double2 u = sm[i & 63]; if (u.x < y && u.y > m) m = u.y; }   // independent loads, dependent max
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + m;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(const char* name, int extra_dep)
{
    double* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8 * 1024);
    for (int warps = 1; warps <= 16; warps *= 4) {          // 1 warp (latency), 4 and 16 warps per CTA on one SM (throughput)
        chain<OP><<<1, 32 * warps>>>(out, cyc, 1.000001, 0.999999);
        cudaDeviceSynchronize();
        chain<OP><<<1, 32 * warps>>>(out, cyc, 1.000001, 0.999999);
        cudaDeviceSynchronize();
        long long c = 0;
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-28s warps/SM %2d: %7.1f cycles per iteration%s\n", name, warps, (double)c / N, warps == 1 ? "  (dependent latency)" : "");
    }
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<0>("DADD chain", 0);
    run<1>("DMUL chain", 0);
    run<2>("DFMA chain", 0);
    run<3>("DSETP+FSEL+DADD chain", 0);
    run<4>("sqrt (IEEE) + DADD chain", 0);
    run<5>("div (IEEE) + DADD chain", 0);
    run<6>("LDS.128 (dependent addr) + max", 0);
    run<7>("LDS.128 (indep.) + running max", 0);
    return 0;
}
