// Where does the hardware place the CTAs of a one-wave grid?  Prints, for grids of 3 and 4 CTAs per SM,
// how the first k CTAs (by blockIdx) spread over the SMs, and the dependent-DFMA latency of one warp.
//   nvcc -arch=sm_100a -o cta_placement cta_placement.cu && ./cta_placement
#include <cstdio>
#include <vector>
#include <algorithm>
__global__ void __launch_bounds__(128, 5) where(int *smid, long long *t0, int spin)
{
    extern __shared__ double pad[];
    unsigned s;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    if (threadIdx.x == 0) { smid[blockIdx.x] = (int)s; t0[blockIdx.x] = clock64(); }
    double a = threadIdx.x;
    for (int i = 0; i < spin; ++i) a = fma(a, 1.0000001, 1e-9);   // keep the CTA resident for a while
    if (a == 12345.0) pad[0] = a;
}
__global__ void dfma_latency(long long *out, double *sink, int iters)
{
    double a = threadIdx.x, b = 1.0000001, c = 1e-9;
    long long t = clock64();
    for (int i = 0; i < iters; ++i) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
    t = clock64() - t;
    if (threadIdx.x == 0) out[0] = t;
    sink[threadIdx.x] = a;
}
int main()
{
    int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    for (int per : {3, 4, 5}) {
        int grid = nsm * per;
        int *d; long long *t; cudaMalloc(&d, grid * sizeof(int)); cudaMalloc(&t, grid * sizeof(long long));
        where<<<grid, 128, 7104>>>(d, t, 20000);
        cudaDeviceSynchronize();
        std::vector<int> h(grid); cudaMemcpy(h.data(), d, grid * sizeof(int), cudaMemcpyDeviceToHost);
        printf("grid %d (%d per SM): first 16 CTAs on SMs:", grid, per);
        for (int i = 0; i < 16; ++i) printf(" %d", h[i]);
        printf("\n");
        for (double frac : {0.25, 0.5, 0.65, 0.72, 0.95, 1.0}) {
            int k = (int)(grid * frac);
            std::vector<int> cnt(256, 0);
            for (int i = 0; i < k; ++i) cnt[h[i]]++;
            int hist[8] = {0};
            int used = 0;
            for (int s = 0; s < 256; ++s) if (cnt[s]) { hist[std::min(cnt[s], 7)]++; used++; }
            printf("  first %4d CTAs: %3d SMs used; SMs with 1/2/3/4/5 CTAs: %d/%d/%d/%d/%d\n", k, used, hist[1], hist[2], hist[3], hist[4], hist[5]);
        }
        cudaFree(d); cudaFree(t);
    }
    long long *o; double *sink; cudaMalloc(&o, 8); cudaMalloc(&sink, 32 * 8);
    for (int rep = 0; rep < 2; ++rep) dfma_latency<<<1, 32>>>(o, sink, 10000);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, o, 8, cudaMemcpyDeviceToHost);
    printf("dependent DFMA latency: %.2f cycles per operation (one warp)\n", (double)h / 40000.0);
    return 0;
}
