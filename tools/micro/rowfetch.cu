// rowfetch.cu -- how long does one warp need to pull a row of 8000 B from an L2-resident 8 MB matrix and fold it into a vector in
// shared memory (the helper's job in cd_lpc2.cu)?  Variants: loads in flight per lane (16 / 32 / 48 sixteen-byte loads = 1 / 2 / 3
// rows), load flavour (__ldg = ld.global.nc through L1, __ldcg = ld.global.cg, cp.async.bulk into shared memory), warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rowfetch rowfetch.cu && ./rowfetch
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int ROWS, int MODE>   // MODE 0: __ldg, 1: __ldcg, 2: no consume (loads only, xor-reduce), 3: bulk copy to smem
__global__ void __launch_bounds__(32) k(const double* __restrict__ P, int n, int ld, int iters, unsigned seed, long long* out, double* sink)
{
    extern __shared__ __align__(128) unsigned char sm[];
    double* g = (double*)sm;                       // [1000]
    double* stage = g + 1024;                      // [ROWS][1000] for MODE 3
    uint64_t* bar = (uint64_t*)(stage + ROWS * 1000);
    const int lane = threadIdx.x;
    for (int i = lane; i < 1024; i += 32) g[i] = 0.0;
    if (MODE == 3 && lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned s = seed + blockIdx.x * 7919u;
    const int n2 = 500;
    double2* g2 = (double2*)g;
    long long t0 = clock64();
    unsigned par = 0;
    for (int it = 0; it < iters; it++) {
        int rows[ROWS];
#pragma unroll
        for (int j = 0; j < ROWS; j++) { s = s * 1664525u + 1013904223u; rows[j] = (s >> 8) % n; }
        if (MODE == 3) {
            if (lane == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(ROWS * 8000) : "memory");
#pragma unroll
                for (int j = 0; j < ROWS; j++)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     smem_u32(stage + j * 1000)), "l"(P + (size_t)rows[j] * ld), "r"(8000), "r"(smem_u32(bar)) : "memory");
            }
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(bar)), "r"(par) : "memory");
            par ^= 1;
#pragma unroll
            for (int u = 0; u < 16; u++) {
                const int c = lane + 32 * u;
                if (c < n2) {
                    double2 gv = g2[c];
#pragma unroll
                    for (int j = 0; j < ROWS; j++) { const double2 r = ((double2*)(stage + j * 1000))[c]; gv.x = fma(r.x, 0.5, gv.x); gv.y = fma(r.y, 0.5, gv.y); }
                    g2[c] = gv;
                }
            }
        } else {
            double2 rv[ROWS][16];
#pragma unroll
            for (int j = 0; j < ROWS; j++) {
                const double2* row = (const double2*)(P + (size_t)rows[j] * ld);
#pragma unroll
                for (int u = 0; u < 16; u++) {
                    const int c = lane + 32 * u;
                    rv[j][u] = (c < n2) ? (MODE == 1 ? __ldcg(&row[c]) : __ldg(&row[c])) : make_double2(0, 0);
                }
            }
            if (MODE == 2) {
                double a = 0;
#pragma unroll
                for (int j = 0; j < ROWS; j++)
#pragma unroll
                    for (int u = 0; u < 16; u++) a += rv[j][u].x + rv[j][u].y;
                g[lane] += a;
            } else {
#pragma unroll
                for (int u = 0; u < 16; u++) {
                    const int c = lane + 32 * u;
                    if (c < n2) {
                        double2 gv = g2[c];
#pragma unroll
                        for (int j = 0; j < ROWS; j++) { gv.x = fma(rv[j][u].x, 0.5, gv.x); gv.y = fma(rv[j][u].y, 0.5, gv.y); }
                        g2[c] = gv;
                    }
                }
            }
        }
        __syncwarp();
    }
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
    if (g[lane] == 1.2345) *sink = g[lane];
}

template <int ROWS, int MODE>
void run(const double* P, int n, int ld, int ctas_per_sm, long long* dout, double* sink, int sms)
{
    const int iters = 400;
    const int grid = sms * ctas_per_sm;
    size_t smem = 8192 + (MODE == 3 ? ROWS * 8000 : 0) + 64;
    cudaFuncSetAttribute(k<ROWS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<ROWS, MODE><<<grid, 32, smem>>>(P, n, ld, 20, 1, dout, sink);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<ROWS, MODE><<<grid, 32, smem>>>(P, n, ld, iters, 2, dout, sink);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long* h = new long long[grid];
    cudaMemcpy(h, dout, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; i++) avg += h[i]; avg /= grid;
    const char* names[] = {"ldg.nc", "ld.cg", "loads only", "bulk->smem"};
    printf("%-10s rows in flight %d  warps/SM %2d : %7.0f cycles per row  (%.0f per batch)  aggregate %.2f TB/s  %s\n", names[MODE], ROWS, ctas_per_sm,
           avg / iters / ROWS, avg / iters, (double)grid * iters * ROWS * 8000.0 / (ms * 1e-3) / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
    delete[] h;
}

int main()
{
    const int n = 1000, ld = 1000;
    double* P; cudaMalloc(&P, (size_t)n * ld * 8); cudaMemset(P, 0, (size_t)n * ld * 8);
    long long* dout; cudaMalloc(&dout, 148 * 32 * 8);
    double* sink; cudaMalloc(&sink, 8);
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount;
    for (int w : {1, 3, 7, 14}) {
        run<1, 0>(P, n, ld, w, dout, sink, sms);
        run<2, 0>(P, n, ld, w, dout, sink, sms);
        run<3, 0>(P, n, ld, w, dout, sink, sms);
        run<1, 1>(P, n, ld, w, dout, sink, sms);
        run<1, 2>(P, n, ld, w, dout, sink, sms);
        run<3, 2>(P, n, ld, w, dout, sink, sms);
        run<1, 3>(P, n, ld, w, dout, sink, sms);
        run<2, 3>(P, n, ld, w, dout, sink, sms);
        run<4, 3>(P, n, ld, w, dout, sink, sms);
    }
    return 0;
}
