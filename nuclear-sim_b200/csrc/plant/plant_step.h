// One plant, one timestep: NuclearPlantSimulator.step
// (reference: nuclear_simulator/simulator/core/sim.py:130-258), physics part.
// Maintenance/threshold monitoring (sim.py:209-223) runs in the flag kernel, not here.
#pragma once
#include "hd.h"
#include "state.h"
#include "prefetch.h"
#include "primary.h"
#include "secondary.h"
#include "report.h"

namespace nps {

// _calculate_primary_to_secondary_coupling: sim.py:335-427
NPS_HD void plant_primary_to_secondary(const PlantState& st, PrimaryConditions& pc) {
    const double power_level = st.pri.power_level;
    double reactor_power_mw = power_level / 100.0 * 3000.0;
    double power_fraction = power_level / 100.0;
    double flow_fraction = py_max(0.3, power_fraction);
    double total_primary_flow = 17100.0 * flow_fraction;
    double cold = 293.0 + 2.0 * (power_fraction - 1.0);
    cold = np_clip(cold, 285.0, 300.0);
    double delta_t_core = (total_primary_flow > 0) ? (reactor_power_mw * 1000.0) / (total_primary_flow * 5.2) : 0.0;
    double hot = cold + delta_t_core;
    hot = np_clip(hot, cold + 5.0, 350.0);
    if (power_fraction < 0.1) hot = cold + 5.0;
    if (is_true(st.sim.has_last_heat_removal_factor)) {
        double effect = (st.sim.last_heat_removal_factor - 1.0) * 3.0;
        cold += effect;
        cold = np_clip(cold, 285.0, 300.0);
        hot = cold + delta_t_core;
        hot = np_clip(hot, cold + 5.0, 350.0);
    }
    double flow_per_loop = total_primary_flow / 3;
    double power_per_loop = reactor_power_mw / 3;
    for (int i = 0; i < 3; ++i) {
        double var = sin(i * 2.0) * 1.0;
        double lh = hot + var;
        double lc = cold + var * 0.5;
        lh = np_clip(lh, lc + 5.0, 350.0);
        lc = np_clip(lc, 285.0, 300.0);
        pc.thermal_power[i] = power_per_loop;
        pc.flow[i] = flow_per_loop;
        pc.inlet_temp[i] = lh;
        pc.outlet_temp[i] = lc;
    }
}

// _apply_secondary_to_primary_feedback: sim.py:429-498
// (the load-factor line, :446/:492, needs the turbine's electrical output: secondary_update_sink writes it)
NPS_HD void plant_secondary_to_primary(PlantState& st) {
    double steam_demand = st.sec.total_steam_flow;
    double hrf = steam_demand / 1665.0;
    bool fw_avail = is_true(st.fw.system_availability);
    double fw_flow = st.fw.total_flow_rate;
    double n_pumps = st.fw.n_running_prev;
    if (!fw_avail) hrf *= 0.5;
    st.pri.steam_flow_rate = steam_demand;
    st.pri.feedwater_pump_status = as_flag(fw_avail);
    st.pri.feedwater_pump_speed = (fw_flow > 0) ? py_min(100.0, (fw_flow / 1665.0) * 100.0) : 0.0;
    st.pri.feedwater_system_available = as_flag(fw_avail);
    st.pri.feedwater_pump_power = st.fw.total_power_consumption;
    st.pri.feedwater_num_running_pumps = n_pumps;
    st.sim.has_last_heat_removal_factor = 1.0;
    st.sim.last_heat_removal_factor = hrf;
    st.sim.last_feedwater_flow_factor = fw_flow / 1665.0;
    st.sim.last_pump_reliability_factor = py_min(1.0, n_pumps / 3.0);
}

// The half of a step that owns primary side, feedwater, steam generators, chemistry and the clock; hands the turbine /
// condenser half what it needs in `h` (secondary disabled: h is left untouched and there is no other half).
NPS_HD void plant_step_source(PlantState& st, const PlantParams& p, const StepInput& in, SecHandoff& h) {
    const double dt = p.dt;
    primary_update(st.pri, p, in, dt);
    if (is_true(p.enable_secondary)) {
        PrimaryConditions pc;
        plant_primary_to_secondary(st, pc);
        st.sim.load_demand = st.pri.power_level;   // sim.py:161 (percent; overrides the caller)
        secondary_update_source<true>(st, p, pc, st.sim.load_demand, st.sim.cooling_water_temp, dt, in, h);
        plant_secondary_to_primary(st);
        if (in.emit_outputs) plant_report_state(st, p);
    }
    st.sim.time_minutes += dt;   // state_manager.advance_time(dt) / self.time += dt : sim.py:183-194 (minutes)
}
NPS_HD void plant_step_sink(PlantState& st, const PlantParams& p, const SecHandoff& h) {
    if (is_true(p.enable_secondary)) secondary_update_sink(st, p, h, p.dt);
}

NPS_HD void plant_step(PlantState& st, const PlantParams& p, const StepInput& in) {
    const double dt = p.dt;
    NPS_PREFETCH(st.pri);
    NPS_PREFETCH(st.sim);
    if (is_true(p.enable_secondary)) {   // first consumers after the primary update: feedwater control and pump 0
        NPS_PREFETCH(st.sec);
        NPS_PREFETCH(st.fw);
        NPS_PREFETCH(st.wc_main);
        NPS_PREFETCH_FAR(st.fw.pump[0]);
    }
    primary_update(st.pri, p, in, dt);
    if (is_true(p.enable_secondary)) {
        PrimaryConditions pc;
        plant_primary_to_secondary(st, pc);
        st.sim.load_demand = st.pri.power_level;   // sim.py:161 (percent; overrides the caller)
        SecHandoff h;
        secondary_update_source<false>(st, p, pc, st.sim.load_demand, st.sim.cooling_water_temp, dt, in, h);
        secondary_update_sink(st, p, h, dt);
        secondary_update_chemistry(st, dt, in);     // :634-665, after the condenser as in the reference
        plant_secondary_to_primary(st);
        if (in.emit_outputs) plant_report_state(st, p);
    }
    // state_manager.advance_time(dt) / self.time += dt : sim.py:183-194 (minutes)
    st.sim.time_minutes += dt;
}

// get_observation: sim.py:290-333
template <class Obs>
NPS_HD void plant_observe(const PlantState& st, const PlantParams& p, Obs&& obs) {
    const PrimaryState& s = st.pri;
    obs[0] = s.neutron_flux / 1e12;
    obs[1] = s.fuel_temperature / 1000;
    obs[2] = s.coolant_temperature / 300;
    obs[3] = s.coolant_pressure / 20;
    obs[4] = s.coolant_flow_rate / 50000;
    obs[5] = s.steam_temperature / 300;
    obs[6] = s.steam_pressure / 10;
    obs[7] = s.steam_flow_rate / 3000;
    obs[8] = s.control_rod_position / 100;
    obs[9] = s.steam_valve_position / 100;
    obs[10] = s.power_level / 100;
    obs[11] = is_true(s.scram_status) ? 1.0 : 0.0;
    if (is_true(p.enable_secondary)) {
        obs[12] = st.sec.electrical_power_output / 1100;
        obs[13] = st.sec.thermal_efficiency / 0.35;
        obs[14] = st.sec.total_steam_flow / 1665;
        obs[15] = st.sec.load_demand / 100;
        obs[16] = st.sec.feedwater_temperature / 250;
        obs[17] = st.sec.cooling_water_temperature / 35;
        obs[18] = st.fw.total_flow_rate / 1665;
        obs[19] = st.fw.total_power_consumption / 40;
        obs[20] = is_true(st.fw.system_availability) ? 1.0 : 0.0;
        obs[21] = st.fw.total_flow_rate / 1665;
    } else {
        for (int i = 12; i < 22; ++i) obs[i] = 0.0;
    }
}

// calculate_reward: sim.py:500-544 (secondary terms use the step's secondary_result)
NPS_HD double plant_reward(const PlantState& st, const PlantParams& p) {
    const PrimaryState& s = st.pri;
    double power_reward = -fabs(s.power_level - 100) / 100;
    double temp_penalty = 0;
    if (s.fuel_temperature > 800) temp_penalty = -(s.fuel_temperature - 800) / 100;
    double pressure_penalty = 0;
    if (s.coolant_pressure > 16) pressure_penalty = -(s.coolant_pressure - 16);
    double scram_penalty = is_true(s.scram_status) ? -100.0 : 0.0;
    double base = power_reward + temp_penalty + pressure_penalty + scram_penalty;
    if (!is_true(p.enable_secondary)) return base;
    double eff_reward = (st.sec.thermal_efficiency - 0.30) * 10;
    double target_el = st.sim.load_demand / 100.0 * 1100.0;
    double el_reward = -fabs(st.sec.electrical_power_output - target_el) / 100;
    double sp_pen = 0;
    if (st.sec.sg_avg_pressure < 5.0 || st.sec.sg_avg_pressure > 8.0) sp_pen = -fabs(st.sec.sg_avg_pressure - 6.895) * 5;
    double cond_pen = 0;
    if (st.sec.condenser_pressure > 0.01) cond_pen = -(st.sec.condenser_pressure - 0.007) * 100;
    double sec_reward = eff_reward + el_reward + sp_pen + cond_pen;
    return base + sec_reward * 0.5;
}

}  // namespace nps
