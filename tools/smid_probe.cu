// Which SM does CTA i of a one-wave grid land on?  576 CTAs x 128 threads, 47 KB dynamic smem, ~128 registers budget
// (4 CTAs per SM like k_sauvola_fused).  Prints blockIdx -> smid for the first CTAs and a histogram of (i - j) for CTA pairs
// sharing an SM.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <map>
__global__ void __launch_bounds__(128, 4) probe(int *smid, long long *t0)
{
    extern __shared__ char sm[];
    unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (threadIdx.x == 0) { smid[blockIdx.x] = (int)s; t0[blockIdx.x] = t; sm[0] = 1; }
    long long e = t + 200000;   // 200 us
    while (true) { long long n; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n)); if (n > e) break; }
}
int main(int argc, char **argv)
{
    int n = argc > 1 ? atoi(argv[1]) : 576;
    int *d; long long *t; cudaMalloc(&d, n * 4); cudaMalloc(&t, n * 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 47 * 1024);
    probe<<<n, 128, 47 * 1024>>>(d, t);
    cudaDeviceSynchronize();
    std::vector<int> h(n); std::vector<long long> ht(n);
    cudaMemcpy(h.data(), d, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(ht.data(), t, n * 8, cudaMemcpyDeviceToHost);
    printf("first 40: "); for (int i = 0; i < 40 && i < n; i++) printf("%d ", h[i]); printf("\n");
    std::map<int, std::vector<int>> by;
    for (int i = 0; i < n; i++) by[h[i]].push_back(i);
    int shown = 0;
    for (auto &kv : by) { if (shown++ >= 12) break; printf("sm %3d:", kv.first); for (int i : kv.second) printf(" %d", i); printf("\n"); }
    long long mn = ht[0], mx = ht[0]; for (auto v : ht) { if (v < mn) mn = v; if (v > mx) mx = v; }
    printf("sms used %zu, start spread %lld ns\n", by.size(), mx - mn);
    return 0;
}
