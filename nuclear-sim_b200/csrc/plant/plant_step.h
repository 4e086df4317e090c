// One plant, one timestep: NuclearPlantSimulator.step
// (reference: nuclear_simulator/simulator/core/sim.py:130-258), physics part.
// Maintenance/threshold monitoring (sim.py:209-223) runs in the flag kernel, not here.
#pragma once
#include "hd.h"
#include "state.h"
#include "primary.h"
#include "secondary.h"

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
NPS_HD void plant_secondary_to_primary(PlantState& st) {
    double steam_demand = st.sec.total_steam_flow;
    double hrf = steam_demand / 1665.0;
    double load_factor = st.sec.electrical_power_output / 1100.0;
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
    st.sim.last_load_factor = load_factor;
    st.sim.last_feedwater_flow_factor = fw_flow / 1665.0;
    st.sim.last_pump_reliability_factor = py_min(1.0, n_pumps / 3.0);
}

NPS_HD void plant_step(PlantState& st, const PlantParams& p, const StepInput& in) {
    const double dt = p.dt;
    primary_update(st.pri, p, in, dt);
    if (is_true(p.enable_secondary)) {
        PrimaryConditions pc;
        plant_primary_to_secondary(st, pc);
        st.sim.load_demand = st.pri.power_level;   // sim.py:161 (percent; overrides the caller)
        secondary_update(st, p, pc, st.sim.load_demand, st.sim.cooling_water_temp, dt, in);
        plant_secondary_to_primary(st);
    }
    // state_manager.advance_time(dt) / self.time += dt : sim.py:183-194 (minutes)
    st.sim.time_minutes += dt;
}

// get_observation: sim.py:290-333 (first 12 entries are primary-only)
NPS_HD void plant_observe_primary(const PlantState& st, double* obs) {
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
}

}  // namespace nps
