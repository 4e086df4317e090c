// Counter-based random draws for a plant step (device-side noise mode).
//
// The reference draws its per-step noise from numpy generators inside the Python objects (one standard normal for the
// ConstantHeatSource noise, one for the pH sensor, up to three uniforms for pH equipment failures:
// heat_sources/constant_heat_source.py:178, systems/secondary/ph_control_system.py:288,409-420).  For parity runs the
// host supplies those very streams (StepInput); a production batch cannot afford 5 doubles per plant-step from the host
// (numpy makes ~1e8 per second per core, the engine consumes 8e8 per second per GPU).  In device-RNG mode each
// (plant, step) gets its draws from Philox4x32-10 keyed by the run seed: stateless, reproducible, independent of the
// batch shape, the launch grouping (K) and the number of GPUs - plant ids are global.  The streams are not numpy's
// MT19937 streams; statistically equivalent, not bit-identical (documented in DESIGN.md 4).
#pragma once
#include "hd.h"

namespace nps {

NPS_HD uint32_t philox_mulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }

// Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11)
NPS_HD void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = philox_mulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = philox_mulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// 53-bit uniform in [0, 1) from two 32-bit words (numpy's random_sample construction: (a >> 5, b >> 6))
NPS_HD double uniform53(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

struct StepDraws { double z_heat, z_ph, u_ph[3]; };

// The five draws of (plant, step) under `seed`.  Counter = (plant, step), one Philox block per pair of outputs;
// normals by Box-Muller on (1 - u1, u2) so the logarithm never sees zero.
NPS_HD StepDraws plant_step_draws(uint64_t seed, uint64_t plant, uint64_t step) {
    StepDraws d;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c[4] = {(uint32_t)plant, (uint32_t)(plant >> 32), (uint32_t)step, (uint32_t)(step >> 32) & 0x3fffffffu};
    uint32_t a[4] = {c[0], c[1], c[2], c[3]};
    philox4x32_10(a, k0, k1);                                  // stream 0: the two normals
    const double u1 = uniform53(a[0], a[1]), u2 = uniform53(a[2], a[3]);
    const double r = sqrt(-2.0 * log(1.0 - u1));
    const double phi = 2.0 * NPS_PI * u2;
    d.z_heat = r * cos(phi);
    d.z_ph = r * sin(phi);
    uint32_t b[4] = {c[0], c[1], c[2], c[3] | 0x40000000u};
    philox4x32_10(b, k0, k1);                                  // stream 1: failure draws 0 and 1
    d.u_ph[0] = uniform53(b[0], b[1]);
    d.u_ph[1] = uniform53(b[2], b[3]);
    uint32_t e[4] = {c[0], c[1], c[2], c[3] | 0x80000000u};
    philox4x32_10(e, k0, k1);                                  // stream 2: failure draw 2
    d.u_ph[2] = uniform53(e[0], e[1]);
    return d;
}

}  // namespace nps
