// TEST INFRASTRUCTURE — not part of the product.
//
// Host build (g++, -ffp-contract=off) of the scalar plant step.  It compiles the same
// single-source physics headers the CUDA kernel compiles (nuclear-sim_b200/csrc/plant/*.h,
// each function citing the reference file:line it restates), so its *independence* comes from
// being pinned against the live Python reference: tests/test_oracle_vs_reference_golden.py
// checks this library against fixtures produced by stepping /root/reference itself
// (oracle/make_golden.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
// leg may load this library; the product (libnps_b200.so) never links or calls it.
//
// Layout here is array-of-structs: state[p * n_state + f]  (one contiguous PlantState per plant).
#include <cstring>
#include <cmath>
#include <cstdint>
#include "plant_step.h"

using namespace nps;

extern "C" {

int nps_oracle_n_state(void) { return (int)(sizeof(PlantState) / sizeof(double)); }
int nps_oracle_n_params(void) { return (int)(sizeof(PlantParams) / sizeof(double)); }

// noise: [n_plants][k_steps][5] = z_heat, z_ph, u0, u1, u2 ; action/magnitude: [n_plants][k_steps]
int nps_oracle_step_sp(double* state, const double* params, const int8_t* action, const double* magnitude,
                       const double* noise, const double* setpoint, int64_t n_plants, int k_steps) {
    PlantParams p;
    std::memcpy(&p, params, sizeof(p));
    const int ns = nps_oracle_n_state();
    for (int64_t i = 0; i < n_plants; ++i) {
        PlantState st;
        std::memcpy(&st, state + i * ns, sizeof(st));
        for (int k = 0; k < k_steps; ++k) {
            StepInput in;
            in.action = action ? (int)action[i * k_steps + k] : (int)ACT_NO_ACTION;
            in.magnitude = magnitude ? magnitude[i * k_steps + k] : 1.0;
            const double* z = noise ? noise + (i * k_steps + k) * 5 : nullptr;
            in.z_heat = z ? z[0] : 0.0;
            in.z_ph = z ? z[1] : 0.0;
            in.u_ph[0] = z ? z[2] : 1.0; in.u_ph[1] = z ? z[3] : 1.0; in.u_ph[2] = z ? z[4] : 1.0;
            in.power_setpoint = setpoint ? setpoint[i * k_steps + k] : NAN;
            plant_step(st, p, in);
        }
        std::memcpy(state + i * ns, &st, sizeof(st));
    }
    return 0;
}

int nps_oracle_step(double* state, const double* params, const int8_t* action, const double* magnitude,
                    const double* noise, int64_t n_plants, int k_steps) {
    return nps_oracle_step_sp(state, params, action, magnitude, noise, nullptr, n_plants, k_steps);
}

// observation [22] + reward for each plant (array-of-structs state)
int nps_oracle_observe(const double* state, const double* params, int64_t n_plants, double* obs, double* reward) {
    PlantParams p; std::memcpy(&p, params, sizeof(p));
    const int ns = nps_oracle_n_state();
    for (int64_t i = 0; i < n_plants; ++i) {
        PlantState st; std::memcpy(&st, state + i * ns, sizeof(st));
        plant_observe(st, p, obs + i * 22);
        reward[i] = plant_reward(st, p);
    }
    return 0;
}

}  // extern "C"

// The device power function's algorithm (csrc/plant/fastpow.h) evaluated on the host, for testing it without a GPU.
// The oracle's own plant_step never uses it (host py_pow is libm pow); taken[i] = 0 where the guarded range refuses.
extern "C" int nps_oracle_fastpow(const double* x, const double* y, double* out, unsigned char* taken, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        double r = 0.0;
        taken[i] = nps::nps_pow_pos(x[i], y[i], r) ? 1 : 0;
        out[i] = r;
    }
    return 0;
}

extern "C" int nps_oracle_fastpow_tab(const double* x, const double* y, double* out, unsigned char* taken, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        double r = 0.0;
        taken[i] = nps::nps_pow_pos_tab(x[i], y[i], r) ? 1 : 0;
        out[i] = r;
    }
    return 0;
}

// maintenance effect on ONE plant (array-of-structs state): returns the MaintStatus code
#include "maintenance.h"
extern "C" int nps_oracle_apply_maintenance(double* state, const double* params, int target, int action, int arg) {
    PlantParams p; std::memcpy(&p, params, sizeof(p));
    PlantState st; std::memcpy(&st, state, sizeof(st));
    int rc = maintenance_apply(st, p, target, action, arg);
    std::memcpy(state, &st, sizeof(st));
    return rc;
}

// ---- per-subsystem entry points used by tests/ to localise a mismatch --------------------------
#include "feedwater.h"
extern "C" int nps_oracle_feedwater(double* state, const double* params, const double* sg_levels,
                                    const double* sg_steam_flows, const double* sg_qualities, double manual_flow,
                                    double fw_temp, double suction, double discharge, double dt, double* out5) {
    PlantParams p; std::memcpy(&p, params, sizeof(p));
    PlantState st; std::memcpy(&st, state, sizeof(st));
    FeedwaterResult r;
    feedwater_update(st.fw, st.wc_main, p, sg_levels, sg_steam_flows, sg_qualities, manual_flow, fw_temp, suction,
                     discharge, dt, r);
    std::memcpy(state, &st, sizeof(st));
    out5[0] = r.total_flow_rate; out5[1] = r.total_power_consumption; out5[2] = r.num_running_pumps;
    out5[3] = r.system_availability; out5[4] = r.sg_flow[0];
    return 0;
}

#include "sg.h"
extern "C" int nps_oracle_sg_system(double* state, const double* params, const double* tin, const double* tout,
                                    const double* flows, double ldf, double sys_ld, double fw_temp,
                                    const double* fw_flows, double dt) {
    PlantParams p; std::memcpy(&p, params, sizeof(p));
    PlantState st; std::memcpy(&st, state, sizeof(st));
    sg_system_update(st.sgs, p, tin, tout, flows, ldf, sys_ld, fw_temp, fw_flows, dt);
    std::memcpy(state, &st, sizeof(st));
    return 0;
}

#include "turbine.h"
extern "C" int nps_oracle_turbine(double* state, const double* params, double load_demand, double cond_p, double dt,
                                  double* out11) {
    PlantParams p; std::memcpy(&p, params, sizeof(p));
    PlantState st; std::memcpy(&st, state, sizeof(st));
    TurbineResult r;
    turbine_update(st.turb, p, turbine_inlet_from(st.sgs), load_demand, cond_p, dt, r);
    std::memcpy(state, &st, sizeof(st));
    std::memcpy(out11, &r, sizeof(r));
    return 0;
}

#include "condenser.h"
extern "C" int nps_oracle_condenser(double* state, const double* params, const double* in8, double dt, double* out7) {
    PlantParams p; std::memcpy(&p, params, sizeof(p));
    PlantState st; std::memcpy(&st, state, sizeof(st));
    CondenserResult r;
    condenser_update(st.cond, p, in8[0], in8[1], in8[2], in8[3], in8[4], in8[5], in8[6], in8[7], dt, r);
    std::memcpy(state, &st, sizeof(st));
    std::memcpy(out7, &r, sizeof(r));
    return 0;
}
