/* The C ABI used from plain C (no Python, no torch): what a non-Python binding of the reference's step loop would do.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/step_from_c.c -o step_from_c \
 *       -L nuclear-sim_b200/_lib -lnps_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/nuclear-sim_b200/_lib
 *   ./step_from_c plant.bin n_plants launches k_substeps out.bin
 *
 * plant.bin : nps_n_state() doubles (one plant's PlantState) followed by nps_n_params() doubles (PlantParams)
 * out.bin   : [n_state] doubles of plant 0, then [22] observation, then reward, after launches x k_substeps steps with
 *             NO_ACTION and no noise (the state of every plant is identical, which the program checks).
 * tests/test_gpu_parity.py builds and runs this and compares out.bin bit for bit with the Python host API. */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "nps_b200.h"

#define CHECK(call) do { if ((call) != 0) { fprintf(stderr, "%s failed: %s\n", #call, nps_last_error()); return 2; } } while (0)
#define CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 3; } } while (0)

int main(int argc, char** argv) {
    if (argc != 6) { fprintf(stderr, "usage: %s plant.bin n_plants launches k_substeps out.bin\n", argv[0]); return 1; }
    const long n = atol(argv[2]);
    const int launches = atoi(argv[3]), k = atoi(argv[4]);
    const int ns = nps_n_state(), np = nps_n_params();
    double* plant = (double*)malloc(sizeof(double) * (size_t)(ns + np));
    FILE* fh = fopen(argv[1], "rb");
    if (!fh || fread(plant, sizeof(double), (size_t)(ns + np), fh) != (size_t)(ns + np)) { fprintf(stderr, "cannot read %s\n", argv[1]); return 1; }
    fclose(fh);

    /* structure-of-arrays slab: field f of plant p at slab[f * n + p] */
    double* h_slab = (double*)malloc(sizeof(double) * (size_t)ns * (size_t)n);
    for (int f = 0; f < ns; ++f)
        for (long p = 0; p < n; ++p) h_slab[(size_t)f * n + p] = plant[f];
    double *d_slab, *d_obs, *d_reward;
    CUDA(cudaSetDevice(0));
    CUDA(cudaMalloc((void**)&d_slab, sizeof(double) * (size_t)ns * (size_t)n));
    CUDA(cudaMalloc((void**)&d_obs, sizeof(double) * NPS_OBS_DIM * (size_t)n));
    CUDA(cudaMalloc((void**)&d_reward, sizeof(double) * (size_t)n));
    CUDA(cudaMemcpy(d_slab, h_slab, sizeof(double) * (size_t)ns * (size_t)n, cudaMemcpyHostToDevice));

    nps_handle* h = NULL;
    CHECK(nps_create(n, 0, &h));
    CHECK(nps_set_params(h, plant + ns, np));
    for (int i = 0; i < launches; ++i)
        CHECK(nps_step(h, d_slab, NULL, NULL, NULL, NULL, k, d_obs, d_reward, NULL, NULL));   /* default stream */
    CUDA(cudaDeviceSynchronize());

    double* h_obs = (double*)malloc(sizeof(double) * NPS_OBS_DIM * (size_t)n);
    double* h_reward = (double*)malloc(sizeof(double) * (size_t)n);
    CUDA(cudaMemcpy(h_slab, d_slab, sizeof(double) * (size_t)ns * (size_t)n, cudaMemcpyDeviceToHost));
    CUDA(cudaMemcpy(h_obs, d_obs, sizeof(double) * NPS_OBS_DIM * (size_t)n, cudaMemcpyDeviceToHost));
    CUDA(cudaMemcpy(h_reward, d_reward, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    for (int f = 0; f < ns; ++f)
        for (long p = 1; p < n; ++p)
            if (memcmp(&h_slab[(size_t)f * n + p], &h_slab[(size_t)f * n], sizeof(double)) != 0) {
                fprintf(stderr, "identical plants diverged: field %s plant %ld\n", nps_field_name(f), p);
                return 4;
            }
    fh = fopen(argv[5], "wb");
    if (!fh) return 1;
    for (int f = 0; f < ns; ++f) fwrite(&h_slab[(size_t)f * n], sizeof(double), 1, fh);
    for (int j = 0; j < NPS_OBS_DIM; ++j) fwrite(&h_obs[(size_t)j * n], sizeof(double), 1, fh);
    fwrite(&h_reward[0], sizeof(double), 1, fh);
    fclose(fh);
    printf("%ld plants x %d steps through the C ABI (%d state fields, abi %d): power %.6f %%\n", n, launches * k, ns,
           nps_abi_version(), h_obs[10 * (size_t)n] * 100.0);
    nps_destroy(h);
    cudaFree(d_slab); cudaFree(d_obs); cudaFree(d_reward);
    free(h_slab); free(h_obs); free(h_reward); free(plant);
    return 0;
}
