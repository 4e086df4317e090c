// Microbenchmark: does a generic-address prefetch (CCTL.E.PF1 / PF2) of LOCAL memory hide DRAM latency for a
// thread-per-plant kernel whose 10 KB per-thread state streams through L1/L2 every substep?
// Mimics nps_step_kernel: 64 threads x 7 blocks/SM, 1304 doubles per thread, chunks of 16 fields, each chunk feeds a
// dependent FP64 chain of CHAIN operations, every field is rewritten.  MODE 0: no prefetch; 1: prefetch.L1 of the next
// chunk (8-byte granularity); 2: same plus the +4 byte word; 3: prefetch.L2; 4: prefetch.L1 two chunks ahead.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
constexpr int NF = 1304, CH = 16;
template <int MODE, int CHAIN>
__global__ void __launch_bounds__(64, 7) k(double* __restrict__ slab, int64_t n, int rounds) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    double s[NF];
#pragma unroll 8
    for (int f = 0; f < NF; ++f) s[f] = slab[(int64_t)f * n + p];
    for (int r = 0; r < rounds; ++r) {
#pragma unroll 1
        for (int c = 0; c + CH <= NF; c += CH) {
            if (MODE != 0) {
                int ahead = (MODE == 4) ? 2 * CH : CH;
                int c2 = c + ahead; if (c2 + CH > NF) c2 = 0;
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    const char* a = reinterpret_cast<const char*>(&s[c2 + j]);
                    if (MODE == 3) asm volatile("prefetch.L2 [%0];" ::"l"(a));
                    else asm volatile("prefetch.L1 [%0];" ::"l"(a));
                    if (MODE == 2) asm volatile("prefetch.L1 [%0];" ::"l"(a + 4));
                }
            }
            double acc = 1.0;
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                double x = s[c + j];
#pragma unroll
                for (int q = 0; q < CHAIN / CH; ++q) acc = acc * 0.999999 + x;
                s[c + j] = acc * 1e-3;
            }
        }
    }
#pragma unroll 8
    for (int f = 0; f < NF; ++f) slab[(int64_t)f * n + p] = s[f];
}
template <int MODE, int CHAIN>
float run(double* d, int64_t n, int rounds) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE, CHAIN><<<(n + 63) / 64, 64>>>(d, n, rounds);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<MODE, CHAIN><<<(n + 63) / 64, 64>>>(d, n, rounds);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms;
}
int main() {
    const int64_t n = 65536; const int rounds = 8;
    double* d; cudaMalloc(&d, sizeof(double) * NF * n); cudaMemset(d, 0, sizeof(double) * NF * n);
    printf("per-thread state %d doubles, %lld threads, %d rounds; ms per launch\n", NF, (long long)n, rounds);
#define ROW(CHAIN) printf("chain %5d/chunk: none %.3f  pfL1 %.3f  pfL1+4 %.3f  pfL2 %.3f  pfL1x2 %.3f\n", CHAIN, \
    run<0, CHAIN>(d, n, rounds), run<1, CHAIN>(d, n, rounds), run<2, CHAIN>(d, n, rounds), run<3, CHAIN>(d, n, rounds), run<4, CHAIN>(d, n, rounds));
    ROW(64) ROW(256) ROW(1024)
    cudaError_t e = cudaGetLastError(); if (e) printf("error %s\n", cudaGetErrorString(e));
    return 0;
}
