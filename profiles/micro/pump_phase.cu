// Round-2 design probe (NOT part of the product library): the feedwater-pump part of a plant step as a UNIT-PARALLEL
// phase kernel working directly on the SoA slab - one thread per (plant, pump), the pump's 56 fields loaded in one
// coalesced burst into registers / local memory, fwp_update_lubrication + fwp_update from the product's own headers,
// results stored straight back.  Timing question: how long does the pump phase of one substep take for 65,536 plants
// in this form, against its share of the monolithic thread-per-plant kernel (19 % of ~380 us)?
// The per-plant couplings (flow demand from the level controller, running-pump bookkeeping, diagnostics) are left out:
// this is a timing probe of the dominant per-unit work, checked against the same functions run on the host.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false --expt-relaxed-constexpr -Xcompiler -fPIC
//        -shared -I nuclear-sim_b200/csrc/plant -o profiles/micro/libpump_phase.so profiles/micro/pump_phase.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>
#include "plant_step.h"
using namespace nps;

constexpr int kPumpFields = sizeof(FWPumpState) / sizeof(double);

__host__ __device__ inline int pump_field0(int pump) {
    PlantState* z = nullptr;
    return (int)(reinterpret_cast<double*>(&z->fw.pump[pump]) - reinterpret_cast<double*>(z));
}
__host__ __device__ inline int prev_levels_field0() {
    PlantState* z = nullptr;
    return (int)(reinterpret_cast<double*>(&z->sec.prev_sg_levels[0]) - reinterpret_cast<double*>(z));
}

template <int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) pump_phase_kernel(double* __restrict__ slab, const __grid_constant__ PlantParams prm,
                                                            int64_t n, int repeats) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 4 * n) return;
    const int pump = (int)(t / n);          // pump index is the slow dimension: a warp reads one field of 32 plants
    const int64_t p = t - (int64_t)pump * n;
    FWPumpState u;
    double* uv = reinterpret_cast<double*>(&u);
    const int f0 = pump_field0(pump);
#pragma unroll
    for (int f = 0; f < kPumpFields; ++f) uv[f] = slab[(int64_t)(f0 + f) * n + p];
    PumpSysCond sc;
    sc.feedwater_temperature = 40.0; sc.suction_pressure = 0.5; sc.discharge_pressure = 7.4;
    const int l0 = prev_levels_field0();
    for (int i = 0; i < 3; ++i) sc.sg_levels[i] = slab[(int64_t)(l0 + i) * n + p];
    for (int r = 0; r < repeats; ++r) {
        fwp_update_lubrication(u, prm, sc, prm.dt);
        fwp_update(u, prm, sc, prm.dt);
    }
#pragma unroll
    for (int f = 0; f < kPumpFields; ++f) slab[(int64_t)(f0 + f) * n + p] = uv[f];
}

extern "C" int pump_phase_launch(double* d_slab, const double* h_params, int64_t n, int repeats, int block, void* stream) {
    PlantParams prm; std::memcpy(&prm, h_params, sizeof(prm));
    const int64_t threads = 4 * n;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((threads + 255) / 256);
    // `block` selects the register budget: resident 256-thread blocks per SM (1: <=255 regs, 2: 128, 3: 80, 4: 64)
    if (block == 1) pump_phase_kernel<256, 1><<<grid, 256, 0, s>>>(d_slab, prm, n, repeats);
    else if (block == 2) pump_phase_kernel<256, 2><<<grid, 256, 0, s>>>(d_slab, prm, n, repeats);
    else if (block == 3) pump_phase_kernel<256, 3><<<grid, 256, 0, s>>>(d_slab, prm, n, repeats);
    else pump_phase_kernel<256, 4><<<grid, 256, 0, s>>>(d_slab, prm, n, repeats);
    return (int)cudaGetLastError();
}

// the same work on the host, plant-major state [n][n_state] (checker for the kernel above)
extern "C" void pump_phase_host(double* state, const double* h_params, int64_t n, int repeats) {
    PlantParams prm; std::memcpy(&prm, h_params, sizeof(prm));
    const int ns = sizeof(PlantState) / sizeof(double);
    for (int64_t p = 0; p < n; ++p) {
        PlantState* st = reinterpret_cast<PlantState*>(state + p * ns);
        PumpSysCond sc;
        sc.feedwater_temperature = 40.0; sc.suction_pressure = 0.5; sc.discharge_pressure = 7.4;
        for (int i = 0; i < 3; ++i) sc.sg_levels[i] = st->sec.prev_sg_levels[i];
        for (int k = 0; k < 4; ++k)
            for (int r = 0; r < repeats; ++r) {
                fwp_update_lubrication(st->fw.pump[k], prm, sc, prm.dt);
                fwp_update(st->fw.pump[k], prm, sc, prm.dt);
            }
    }
}
extern "C" int pump_phase_fields(void) { return kPumpFields; }
