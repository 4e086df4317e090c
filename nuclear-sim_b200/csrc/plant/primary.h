// Primary-side plant physics: control actuation, heat source (constant or point-kinetics
// reactor), fuel/coolant thermal hydraulics, toy steam cycle, NaN guard and scram latch.
// Restates, in call order, PrimaryReactorPhysics.update_system
// (reference: nuclear_simulator/systems/primary/__init__.py:178-287).
#pragma once
#include "hd.h"
#include "state.h"

namespace nps {

// ControlAction enum values: systems/primary/__init__.py:28-45
enum Action : int {
    ACT_CONTROL_ROD_INSERT = 0, ACT_CONTROL_ROD_WITHDRAW = 1, ACT_INCREASE_COOLANT_FLOW = 2,
    ACT_DECREASE_COOLANT_FLOW = 3, ACT_OPEN_STEAM_VALVE = 4, ACT_CLOSE_STEAM_VALVE = 5,
    ACT_INCREASE_FEEDWATER = 6, ACT_DECREASE_FEEDWATER = 7, ACT_NO_ACTION = 8,
    ACT_DILUTE_BORON = 9, ACT_BORATE_COOLANT = 10
};

// Per-plant, per-step inputs (host-supplied random streams included).
struct StepInput {
    int action;        // ControlAction value (8 = NO_ACTION)
    double magnitude;  // action magnitude
    double z_heat;     // standard-normal draw for ConstantHeatSource noise (constant_heat_source.py:178)
    double z_ph;       // standard-normal draw for the pH sensor noise (ph_control_system.py:288)
    double u_ph[3];    // uniform draws for pH equipment failures (ph_control_system.py:409,414,420)
    double power_setpoint;  // heat_source.set_power_setpoint(%) applied before the step; NaN = unchanged
    // false on the non-final substeps of a fused launch: report-only quantities that nothing reads back (the pH
    // controller's 100-sample deviation RMS, the stage inlet entropy) are pure functions of the state at the moment
    // they are observed, and nobody observes the state between fused substeps, so they are evaluated once, at the end.
    bool emit_outputs = true;
    // NPS_STATUS_* bits raised by this step (the reference resets / latches silently; the batched engine surfaces them
    // per plant: SURVEY 8b).  Written by the physics, read by the caller after plant_step.
    mutable unsigned status = 0;
};
constexpr unsigned kStatusNanReset = 1u;   // thermal_hydraulics.py:257-269 reset five primary fields this step
constexpr unsigned kStatusScram = 2u;          // scram latched this step (scram_logic.py:34-61)

// _apply_control_actions: systems/primary/__init__.py:289-359 with the action routing of
// NuclearPlantSimulator._convert_action_to_control_inputs (simulator/core/sim.py:260-288).
NPS_HD void primary_apply_control(PrimaryState& s, int action, double mag, double dt) {
    const double max_control_rod_speed = 5.0, max_valve_speed = 10.0, max_flow_change_rate = 1000.0;
    if (action == ACT_CONTROL_ROD_INSERT)
        s.control_rod_position = py_max(0.0, s.control_rod_position - max_control_rod_speed * dt * mag);
    else if (action == ACT_CONTROL_ROD_WITHDRAW)
        s.control_rod_position = py_min(100.0, s.control_rod_position + max_control_rod_speed * dt * mag);
    if (action == ACT_INCREASE_COOLANT_FLOW)
        s.coolant_flow_rate = py_min(50000.0, s.coolant_flow_rate + max_flow_change_rate * dt * mag);
    else if (action == ACT_DECREASE_COOLANT_FLOW)
        s.coolant_flow_rate = py_max(5000.0, s.coolant_flow_rate - max_flow_change_rate * dt * mag);
    if (action == ACT_DILUTE_BORON)
        s.boron_concentration = py_max(0.0, s.boron_concentration - 50.0 * dt * mag);
    else if (action == ACT_BORATE_COOLANT)
        s.boron_concentration = py_min(3000.0, s.boron_concentration + 50.0 * dt * mag);
    if (action == ACT_OPEN_STEAM_VALVE)
        s.steam_valve_position = py_min(100.0, s.steam_valve_position + max_valve_speed * dt * mag);
    else if (action == ACT_CLOSE_STEAM_VALVE)
        s.steam_valve_position = py_max(0.0, s.steam_valve_position - max_valve_speed * dt * mag);
}

// ConstantHeatSource.update: heat_sources/constant_heat_source.py:104-141 (+ :68-83, :169-183)
NPS_HD void constant_heat_source_update(PrimaryState& s, const PlantParams& p, double dt, double z,
                                        double& thermal_power_mw, double& power_percent) {
    s.hs_time += dt;
    s.hs_current_power_mw = (s.hs_setpoint_percent / 100.0) * p.rated_power_mw;
    double final_power = s.hs_current_power_mw;
    if (is_true(p.noise_enabled)) {
        double noise_std_mw = (p.noise_std_percent / 100.0) * s.hs_current_power_mw;
        s.hs_raw_noise_mw = 0.0 + noise_std_mw * z;  // RandomState.normal(loc, scale) = loc + scale*gauss
        double alpha = dt / (p.noise_filter_time_constant + dt);
        s.hs_filtered_noise_mw = alpha * s.hs_raw_noise_mw + (1.0 - alpha) * s.hs_filtered_noise_mw;
        double noisy = s.hs_current_power_mw + s.hs_filtered_noise_mw;
        final_power = py_max(0.0, py_min(noisy, p.rated_power_mw));
    }
    s.hs_total_energy_mwh += final_power * (dt / 3600.0);
    thermal_power_mw = final_power;
    power_percent = (final_power / p.rated_power_mw) * 100.0;
}

// ReactivityModel.calculate_total_reactivity: reactor/reactivity_model.py:77-311 (pcm).
// Python sum() adds the ten terms left to right starting from integer 0.
NPS_HD double total_reactivity_pcm(const PrimaryState& s) {
    double pos_norm = np_clip(s.control_rod_position / 100.0, 0.0, 1.0);
    double rods = 3000.0 * (pos_norm - 0.5);
    double boron = -10.0 * s.boron_concentration;
    double doppler = -2.5e-5 * (s.fuel_temperature - 575.0) * 1e5;
    double mod_t = -3.0e-5 * (s.coolant_temperature - 280.0) * 1e5;
    double mod_void = -1000.0 * s.coolant_void_fraction;
    double pressure = 0.5 * (s.coolant_pressure - 15.5);
    double xenon = (s.xenon_concentration / 1.0e15) * -1800.0;
    double samarium = (s.samarium_concentration / 5.0e14) * -600.0;
    double depletion = 3340.0 + -0.15 * s.fuel_burnup;
    double bp = s.burnable_poison_worth * nps_exp(-0.0002 * s.fuel_burnup);
    double total = 0.0 + rods;
    total += boron; total += doppler; total += mod_t; total += mod_void; total += pressure;
    total += xenon; total += samarium; total += depletion; total += bp;
    return total;
}

// ReactorHeatSource.update: heat_sources/reactor_heat_source.py:40-107 with
// update_fission_products (reactivity_model.py:313-367) and the point-kinetics model
// (physics/point_kinetics.py:26-131).
NPS_HD void reactor_heat_source_update(PrimaryState& s, const PlantParams& p, double dt,
                                       double& thermal_power_mw, double& power_percent,
                                       double& reactivity_pcm) {
    const double flux = s.neutron_flux;
    {   // fission products, explicit Euler
        double iodine = s.iodine_concentration, xenon = s.xenon_concentration, samarium = s.samarium_concentration;
        double fission_rate = flux * 1e-12;
        double diodine = 0.064 * fission_rate - 2.87e-5 * iodine;
        double new_iodine = iodine + diodine * dt;
        double xe_prod = 0.061 * fission_rate;
        double xe_from_i = 2.87e-5 * iodine;
        double xe_decay = 2.09e-5 * xenon;
        double xe_abs = 2.65e6 * 1e-24 * flux * xenon;
        double dxenon = xe_prod + xe_from_i - xe_decay - xe_abs;
        double new_xenon = xenon + dxenon * dt;
        double sm_prod = 0.0137 * fission_rate;
        double sm_abs = 4.1e4 * 1e-24 * flux * samarium;
        double new_sm = samarium + (sm_prod - sm_abs) * dt;
        s.xenon_concentration = py_max(0.0, new_xenon);
        s.iodine_concentration = py_max(0.0, new_iodine);
        s.samarium_concentration = py_max(0.0, new_sm);
    }
    double total = total_reactivity_pcm(s);
    double reactivity = total / 100000.0;
    if (is_true(s.scram_status)) reactivity = -0.5;

    // solve_point_kinetics
    const double BETA = 0.0065, LAMBDA_PROMPT = 1e-5;
    const double LAMBDA[6] = {0.077, 0.311, 1.40, 3.87, 1.40, 0.195};
    double rho = np_clip(reactivity, -0.9, 0.1);
    double flux_dot = 0.0;
    double prec_dot[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (!(fabs(rho) < 0.01)) {
        flux_dot = (rho - BETA) / LAMBDA_PROMPT * s.neutron_flux;
        for (int i = 0; i < 6; ++i) flux_dot += LAMBDA[i] * s.precursors[i];
        double max_change = s.neutron_flux * 0.1;  // |rho| >= 0.01 here, other bands unreachable
        flux_dot = np_clip(flux_dot, -max_change, max_change);
        const double beta_i = BETA / 6;
        for (int i = 0; i < 6; ++i)
            prec_dot[i] = beta_i / LAMBDA_PROMPT * s.neutron_flux - LAMBDA[i] * s.precursors[i];
    }
    s.neutron_flux += flux_dot * dt;
    s.neutron_flux = np_clip(s.neutron_flux, 1e8, 1e14);
    for (int i = 0; i < 6; ++i) {
        s.precursors[i] += prec_dot[i] * dt;
        s.precursors[i] = np_clip(s.precursors[i], 0.0, 1.0);
    }
    double pfrac = s.neutron_flux / 1e13;
    thermal_power_mw = pfrac * p.rated_power_mw;
    power_percent = pfrac * 100.0;
    s.power_level = power_percent;
    s.reactivity = reactivity;
    reactivity_pcm = total;
}

// calculate_heat_transfer_coefficient: physics/thermal_hydraulics.py:111-166
NPS_HD double primary_overall_ua(double coolant_flow_rate) {
    const double fuel_rod_diameter = 0.0095, fuel_rod_length = 3.66, num_fuel_rods = 50000.0;
    const double area = 3.141592653589793 * fuel_rod_diameter * fuel_rod_length * num_fuel_rods;
    const double density = 700.0, viscosity = 9.0e-5, k = 0.55, cp = 5200.0, flow_area = 10.0;
    double velocity = coolant_flow_rate / (density * flow_area);
    double reynolds = density * velocity * fuel_rod_diameter / viscosity;
    reynolds = py_max(reynolds, 1000.0);
    double prandtl = viscosity * cp / k;
    double nusselt = 0.023 * py_pow(reynolds, 0.8) * py_pow(prandtl, 0.4);
    double h = nusselt * k / fuel_rod_diameter;
    double ua = h * area;
    ua = ua * 0.1;
    return np_clip(ua, 10e6, 50e6);
}

// calculate_thermal_hydraulics / calculate_steam_cycle / update_thermal_state /
// update_steam_state / check_for_nan_values: physics/thermal_hydraulics.py:26-270
NPS_HD bool primary_thermal_hydraulics(PrimaryState& s, double thermal_power_w, double dt) {
    const double FUEL_MASS = 200000.0, FUEL_CP = 1500.0;
    double heat_removal = primary_overall_ua(s.coolant_flow_rate) * (s.fuel_temperature - s.coolant_temperature);
    double fuel_temp_dot = (thermal_power_w - heat_removal) / (FUEL_MASS * FUEL_CP);
    bool near100 = fabs(s.power_level - 100.0) < 5.0;
    fuel_temp_dot = near100 ? np_clip(fuel_temp_dot, -1.0, 1.0) : np_clip(fuel_temp_dot, -10.0, 10.0);
    double power_fraction = s.power_level / 100.0;
    double target_hot = 293.0 + (34.0 * power_fraction);
    double target_avg = (target_hot + 293.0) / 2.0;
    double temp_error = target_avg - s.coolant_temperature;
    double coolant_temp_dot = 0.1 * temp_error;
    coolant_temp_dot = near100 ? np_clip(coolant_temp_dot, -0.5, 0.5) : np_clip(coolant_temp_dot, -5.0, 5.0);
    double temp_pressure_effect = 0.002 * (s.coolant_temperature - 293.0);
    double pressure_error = s.coolant_pressure - (15.5 + temp_pressure_effect);
    double pressure_dot = np_clip(-0.01 * pressure_error, -0.05, 0.05);

    // steam cycle derivatives are evaluated before any state update
    double steam_generation = py_min(s.coolant_flow_rate * 0.05, s.steam_valve_position / 100 * 2000);
    double steam_temp_dot = 0.1 * (s.coolant_temperature - s.steam_temperature);
    double steam_pressure_dot = 0.05 * (steam_generation - s.steam_flow_rate);
    double steam_flow_dot = s.steam_valve_position / 100 * 20 - 10;
    double feedwater_flow_dot = steam_generation - s.feedwater_flow_rate;

    s.fuel_temperature = np_clip(s.fuel_temperature + fuel_temp_dot * dt, 200.0, 2000.0);
    s.coolant_temperature = np_clip(s.coolant_temperature + coolant_temp_dot * dt, 200.0, 400.0);
    s.coolant_pressure = np_clip(s.coolant_pressure + pressure_dot * dt, 10.0, 20.0);

    s.steam_temperature = np_clip(s.steam_temperature + steam_temp_dot * dt, 200.0, 400.0);
    s.steam_pressure = np_clip(s.steam_pressure + steam_pressure_dot * dt, 1.0, 10.0);
    s.steam_flow_rate = np_clip(s.steam_flow_rate + steam_flow_dot * dt, 0.0, 3000.0);
    s.feedwater_flow_rate = np_clip(s.feedwater_flow_rate + feedwater_flow_dot * dt, 0.0, 3000.0);

    if (isnan(s.fuel_temperature) || isnan(s.neutron_flux) || isnan(s.coolant_temperature) ||
        isnan(s.coolant_pressure)) {
        s.neutron_flux = 1e12;
        s.fuel_temperature = 600.0;
        s.coolant_temperature = 280.0;
        s.coolant_pressure = 15.5;
        s.power_level = 100.0;
        return true;
    }
    return false;
}

// ScramSystem.check_safety_systems: reactor/safety/scram_logic.py:24-61
NPS_HD bool primary_check_scram(PrimaryState& s) {
    bool any = (s.fuel_temperature > 1200.0) || (s.coolant_pressure > 17.2) ||
               (s.coolant_flow_rate < 5000.0) || (s.power_level > 118.0);
    if (any && !is_true(s.scram_status)) {
        s.scram_status = 1.0;
        s.control_rod_position = 0.0;
        return true;
    }
    return false;
}

// PrimaryReactorPhysics.update_system: systems/primary/__init__.py:178-287
NPS_HD void primary_update(PrimaryState& s, const PlantParams& p, const StepInput& in, double dt) {
    if (!isnan(in.power_setpoint)) {   // ConstantHeatSource.set_power_setpoint: constant_heat_source.py:94-102
        s.hs_setpoint_percent = np_clip(in.power_setpoint, 0.0, 150.0);
        s.hs_current_power_mw = (s.hs_setpoint_percent / 100.0) * p.rated_power_mw;
    }
    primary_apply_control(s, in.action, in.magnitude, dt);
    double thermal_mw, power_pct, rho_pcm = 0.0;
    if (p.heat_source_type == 0.0) {
        constant_heat_source_update(s, p, dt, in.z_heat, thermal_mw, power_pct);
        s.thermal_power_mw = thermal_mw;
        s.power_level = power_pct;
        s.total_reactivity_pcm = 0.0;
    } else {
        reactor_heat_source_update(s, p, dt, thermal_mw, power_pct, rho_pcm);
        s.thermal_power_mw = thermal_mw;
        s.power_level = power_pct;
        s.total_reactivity_pcm = rho_pcm;
        s.reactivity = rho_pcm / 100000.0;
    }
    if (primary_thermal_hydraulics(s, s.thermal_power_mw * 1e6, dt)) in.status |= kStatusNanReset;
    const bool scram_now = primary_check_scram(s);
    if (scram_now) in.status |= kStatusScram;
    s.scram_activated = as_flag(scram_now);
}

}  // namespace nps
