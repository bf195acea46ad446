// Do non-FP64 instructions issued between FP64 instructions cost FP64 throughput?
// Per trip: 8 independent DFMA chains plus N independent integer (or FP32) operations; 8 warps per sub-partition.
// If the scheduler can issue to another pipe while the FP64 pipe takes its second cycle, time stays flat up to
// N = 8 (16 FP64 pipe cycles per trip >= 8 + N issue slots).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix fp64_mix.cu && ./fp64_mix
#include <cstdio>
#include <cuda_runtime.h>

template <int N, int KIND, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k(double *out, int iters, double seed)
{
    double a[8];
    unsigned u[16];
    float f[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < 16; ++i) { u[i] = threadIdx.x * 17u + i; f[i] = threadIdx.x + 0.5f * i; }
    const double B = 1.0000000001, C = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            a[i] = __fma_rn(a[i], B, C);
#pragma unroll
            for (int j = 0; j < N / 8; ++j) {
                const int q = (i * (N / 8) + j) & 15;
                if (KIND == 0) u[q] = (u[q] ^ (u[q] >> 3)) + 0x9E3779B9u;          // ALU ops
                else f[q] = fmaf(f[q], 1.0001f, 0.25f);                              // one FMA-pipe op
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += u[i] + f[i];
    if (s == 12345.678) out[0] = s;
}

template <int N, int KIND, int WARPS>
void run(int sms)
{
    double *d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 14, grid = sms * 4;      // 4 CTAs per SM of WARPS warps: WARPS warps per sub-partition
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<N, KIND, WARPS><<<grid, WARPS * 32>>>(d, iters, 1.0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double ops = (double)grid * WARPS * 32 * iters * 8;
    printf("%d warps/sub-partition, 8 DFMA + %2d x %s per trip: %.3f ms  %.2f T DFMA/s\n", WARPS, N,
           KIND ? "FFMA" : "ALU pair", best, ops / (best * 1e-3) / 1e12);
    cudaFree(d);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int s = p.multiProcessorCount;
    run<0, 0, 8>(s); run<8, 0, 8>(s); run<16, 0, 8>(s); run<32, 0, 8>(s);
    run<8, 1, 8>(s); run<16, 1, 8>(s); run<32, 1, 8>(s);
    run<0, 0, 3>(s); run<8, 0, 3>(s); run<8, 1, 3>(s); run<16, 1, 3>(s);
    run<0, 0, 1>(s); run<8, 1, 1>(s);
    return 0;
}
