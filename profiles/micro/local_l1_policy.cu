// Microbenchmark: how does the B200 L1 treat local-memory stores and reloads?
// One warp per block, `blocks_per_sm` resident blocks, per-thread local array of M doubles, dynamically indexed so it stays
// in local memory.  Pattern per element: load a[i], FMA, store a[i] (read-modify-write), rounds x M times, dependent chain.
// Reports cycles per element visit for footprints that fit L1 and footprints that do not, and for a load-only pattern.
#include <cstdio>
#include <cuda_runtime.h>
template <int M, bool STORE>
__global__ void k(double* out, int rounds, const int* __restrict__ perm, long long* cycles) {
    double a[M];
    for (int i = 0; i < M; ++i) a[i] = threadIdx.x + i;
    double acc = 0.0;
    int idx = perm[threadIdx.x & 1];   // opaque start index: keeps the array in local memory
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
#pragma unroll 1
        for (int i = 0; i < M; ++i) {
            double v = a[idx];
            acc = acc * 0.999 + v;
            if (STORE) a[idx] = acc * 1e-3;
            idx = idx + 1; if (idx >= M) idx = 0;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + a[idx];
}
template <int M, bool STORE>
void run(int blocks, int rounds, double* out, int* perm, long long* cyc) {
    k<M, STORE><<<blocks, 32>>>(out, 2, perm, cyc);
    k<M, STORE><<<blocks, 32>>>(out, rounds, perm, cyc);
    long long h; cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("  M=%4d doubles/thread (%7.1f KB per SM at %2d warps/SM)  %s: %7.1f cycles per element visit\n", M,
           M * 8.0 * 32 * (blocks / 148) / 1024.0, blocks / 148, STORE ? "load+store" : "load only ", (double)h / ((double)rounds * M));
}
int main() {
    double* out; int* perm; long long* cyc;
    cudaMalloc(&out, 148 * 16 * 32 * sizeof(double)); cudaMalloc(&perm, 8); cudaMemset(perm, 0, 8); cudaMalloc(&cyc, 8);
    for (int wps : {1, 14}) {
        int blocks = 148 * wps;
        printf("%d warp(s) per SM\n", wps);
        run<16, true>(blocks, 64, out, perm, cyc);   run<16, false>(blocks, 64, out, perm, cyc);
        run<64, true>(blocks, 32, out, perm, cyc);   run<64, false>(blocks, 32, out, perm, cyc);
        run<256, true>(blocks, 8, out, perm, cyc);   run<256, false>(blocks, 8, out, perm, cyc);
        run<1304, true>(blocks, 4, out, perm, cyc);  run<1304, false>(blocks, 4, out, perm, cyc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
