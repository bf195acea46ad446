// Does the FP64 pipe's issue rate depend on how many DISTINCT register operands an instruction reads?
// Eight independent chains per thread, 8 warps per SM sub-partition; variants differ only in operands:
//   0: a = fma(a, B, C)      B, C the same registers for all chains (operand reuse possible)
//   1: a_i = fma(a_i, b_i, c_i)   three distinct register pairs per instruction, no two alike in a row
//   2: a_i = a_i * b_i       two distinct
//   3: a_i = a_i + b_i       two distinct
//   4: a_i = fma(a_i, b_i, a_i)   two distinct (accumulator twice)
// nvcc -arch=sm_100a -O3 -o fp64_operands fp64_operands.cu && ./fp64_operands
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, double seed)
{
    double a[8], b[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x + i; b[i] = 1.0 + 1e-9 * (i + 1) + 1e-12 * threadIdx.x; c[i] = 1e-9 * (i + 2); }
    const double B = b[0], C = c[0];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = __fma_rn(a[i], B, C);
            if (MODE == 1) a[i] = __fma_rn(a[i], b[i], c[i]);
            if (MODE == 2) a[i] = __dmul_rn(a[i], b[i]);
            if (MODE == 3) a[i] = __dadd_rn(a[i], c[i]);
            if (MODE == 4) a[i] = __fma_rn(a[i], b[i], a[i]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + b[i] + c[i];
    if (s == 12345.678) out[0] = s;
}

template <int MODE>
void run(const char *name, int sms)
{
    double *d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 14, grid = sms * 8;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<grid, 256>>>(d, iters, 1.0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double ops = (double)grid * 256 * iters * 8;
    printf("%-40s %.3f ms  %.2f T op/s\n", name, best, ops / (best * 1e-3) / 1e12);
    cudaFree(d);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
    run<0>("DFMA a,B,C (shared operands)", p.multiProcessorCount);
    run<1>("DFMA a_i,b_i,c_i (3 distinct)", p.multiProcessorCount);
    run<2>("DMUL a_i,b_i (2 distinct)", p.multiProcessorCount);
    run<3>("DADD a_i,c_i (2 distinct)", p.multiProcessorCount);
    run<4>("DFMA a_i,b_i,a_i (2 distinct)", p.multiProcessorCount);
    return 0;
}
