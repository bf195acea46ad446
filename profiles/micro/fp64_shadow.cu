// Which instruction classes issue in the shadow of the FP64 pipe?  A DFMA occupies the FP64 pipe of a sub-partition for two
// cycles; per trip: 8 independent DFMA chains plus N instructions of ONE other class (independent of each other and of
// the DFMAs), W warps per sub-partition.  If the class hides behind the DFMAs, time stays flat up to N = 8.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_shadow fp64_shadow.cu && ./fp64_shadow
#include <cstdio>
#include <cuda_runtime.h>

enum { K_NONE, K_IMAD, K_IADD, K_LOP3, K_SHF, K_FMNMX, K_ISETP_SEL, K_FFMA, K_FMUL, K_LDS, K_MUFU, K_MOVSEL, K_COUNT };
static const char *NAMES[] = {"none", "IMAD", "IADD3", "LOP3", "SHF", "FMNMX", "ISETP+SEL", "FFMA", "FMUL", "LDS", "MUFU.RSQ", "PRMT"};

template <int N, int KIND, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k(double *out, int iters, double seed, unsigned one)
{
    __shared__ unsigned tab[256];
    tab[threadIdx.x & 255] = threadIdx.x;
    __syncthreads();
    double a[8];
    unsigned u[16];
    float f[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < 16; ++i) { u[i] = threadIdx.x * 17u + i; f[i] = threadIdx.x + 0.5f * i + 1.0f; }
    const double B = 1.0000000001, C = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            a[i] = __fma_rn(a[i], B, C);
#pragma unroll
            for (int j = 0; j < N / 8; ++j) {
                const int q = (i * (N / 8) + j) & 15;
                if (KIND == K_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, 12345;" : "+r"(u[q]) : "r"(one));
                else if (KIND == K_IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[q]) : "r"(u[(q + 5) & 15]));
                else if (KIND == K_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[q]) : "r"(u[(q + 5) & 15]), "r"(one));
                else if (KIND == K_SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(u[q]) : "r"(u[(q + 5) & 15]));
                else if (KIND == K_FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(f[q]) : "f"(f[(q + 5) & 15]));
                else if (KIND == K_ISETP_SEL) asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %1, %0, p; }" : "+r"(u[q]) : "r"(u[(q + 5) & 15]));
                else if (KIND == K_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[q]) : "f"(f[(q + 5) & 15]), "f"(f[(q + 9) & 15]));
                else if (KIND == K_FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[q]) : "f"(f[(q + 5) & 15]));
                else if (KIND == K_LDS) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u[q]) : "r"((unsigned)__cvta_generic_to_shared(tab) + ((u[(q + 5) & 15] & 255u) << 2)));
                else if (KIND == K_MUFU) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(f[q]));
                else if (KIND == K_MOVSEL) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(u[q]) : "r"(u[(q + 5) & 15]));
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
float run(int sms)
{
    double *d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 13, grid = sms * 4;      // 4 CTAs per SM of WARPS warps: WARPS warps per sub-partition
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<N, KIND, WARPS><<<grid, WARPS * 32>>>(d, iters, 1.0, 1u);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    cudaFree(d);
    return best;
}

template <int KIND, int WARPS>
void row(int sms)
{
    const float t0 = run<0, K_NONE, WARPS>(sms);
    const float t8 = run<8, KIND, WARPS>(sms), t16 = run<16, KIND, WARPS>(sms), t32 = run<32, KIND, WARPS>(sms);
    // cycles per trip and sub-partition relative to the 16 cycles of the 8 DFMAs alone
    printf("%d warps/sub-partition, 8 DFMA + N x %-10s  N=0: 16.0 cycles   N=8: %5.1f   N=16: %5.1f   N=32: %5.1f\n", WARPS, NAMES[KIND],
           16.f * t8 / t0, 16.f * t16 / t0, 16.f * t32 / t0);
}

template <int WARPS>
void all(int s)
{
    row<K_IMAD, WARPS>(s); row<K_IADD, WARPS>(s); row<K_LOP3, WARPS>(s); row<K_SHF, WARPS>(s); row<K_FMNMX, WARPS>(s);
    row<K_ISETP_SEL, WARPS>(s); row<K_FFMA, WARPS>(s); row<K_FMUL, WARPS>(s); row<K_LDS, WARPS>(s); row<K_MUFU, WARPS>(s);
    row<K_MOVSEL, WARPS>(s);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    all<4>(p.multiProcessorCount);
    all<2>(p.multiProcessorCount);
    return 0;
}
