// Does a local-memory STORE allocate the line in L1 (so that the reload hits), or is it write-through/no-allocate
// (reload pays an L2 round trip)?  Per round: store M fresh values (lines last touched long ago, evicted by a 1 MB/SM
// sweep in between), then reload them in a dependent chain.  Compare with reload-after-LOAD (lines brought in by loads).
#include <cstdio>
#include <cuda_runtime.h>
template <int M, int BIG, bool BY_STORE>
__global__ void k(double* out, int rounds, const int* __restrict__ perm, long long* cycles) {
    double a[M];
    double big[BIG];
    int o = perm[threadIdx.x & 1];
    for (int i = 0; i < BIG; ++i) big[(i + o) % BIG] = i;
    for (int i = 0; i < M; ++i) a[(i + o) % M] = i;
    double acc = 0.0;
    long long total = 0;
    for (int r = 0; r < rounds; ++r) {
        for (int i = 0; i < BIG; ++i) acc += big[(i + o) % BIG];        // evict a[] from L1 (and keep big[] streaming)
        double tmp = 0.0;
        if (BY_STORE) { for (int i = 0; i < M; ++i) a[(i + o) % M] = acc + i; }          // bring a[] "in" by storing
        else { for (int i = 0; i < M; ++i) tmp += a[(i + o) % M]; }                      // ... or by loading
        acc += tmp;
        int idx = o;
        long long t0 = clock64();
#pragma unroll 1
        for (int i = 0; i < M; ++i) { acc = acc * 0.999 + a[idx]; idx = (idx + 1 == M) ? 0 : idx + 1; }   // dependent reloads
        total += clock64() - t0;
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = total;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int M, int BIG, bool S>
void run(double* out, int* perm, long long* cyc) {
    const int rounds = 8, blocks = 148 * 14;
    k<M, BIG, S><<<blocks, 32>>>(out, 1, perm, cyc);
    k<M, BIG, S><<<blocks, 32>>>(out, rounds, perm, cyc);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("  M=%3d  evicting sweep %4d doubles/thread  reload after %s: %7.1f cycles per dependent reload\n", M, BIG,
           S ? "STORE" : "LOAD ", (double)h / (rounds * (double)M));
}
int main() {
    double* out; int* perm; long long* cyc;
    cudaMalloc(&out, 148 * 14 * 32 * 8); cudaMalloc(&perm, 8); cudaMemset(perm, 0, 8); cudaMalloc(&cyc, 8);
    printf("14 warps per SM\n");
    run<16, 512, true>(out, perm, cyc);  run<16, 512, false>(out, perm, cyc);
    run<16, 16, true>(out, perm, cyc);   run<16, 16, false>(out, perm, cyc);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
