// Secondary-system orchestration: feedwater -> steam generators -> turbine -> condenser ->
// water chemistry / pH control -> electrical-power gating.
// Restates SecondaryReactorPhysics.update_system
// (reference: nuclear_simulator/systems/secondary/__init__.py:340-1021).
// The heat-flow and chemistry-flow trackers (:668-744) are read-only aggregators that feed
// nothing back into component state; their derived columns are not carried in PlantState.
#pragma once
#include "hd.h"
#include "state.h"
#include "prefetch.h"
#include "feedwater.h"
#include "sg.h"
#include "turbine.h"
#include "condenser.h"
#include "ph_control.h"

namespace nps {

struct PrimaryConditions { double inlet_temp[3], outlet_temp[3], flow[3], thermal_power[3]; };

// _saturation_temperature (Clausius-Clapeyron variant): systems/secondary/__init__.py:1455-1492
NPS_HD_SHARED double secondary_sat_temp(double p_mpa) {
    if (p_mpa <= 0.001) return 10.0;
    const double p_ref = 0.101325, t_ref = 100.0, h_fg = 2257.0, r_v = 0.4615;
    double t_ref_k = t_ref + 273.15;
    double ratio = p_mpa / p_ref;
    double t;
    if (ratio > 0) {
        double tk = 1.0 / (1.0 / t_ref_k - (r_v / h_fg) * nps_log(ratio));
        t = tk - 273.15;
    } else {
        t = t_ref;
    }
    return np_clip(t, 10.0, 374.0);
}

// What the turbine / condenser / energy-bookkeeping half of a step (secondary_update_sink) needs from the half that
// advances primary side, feedwater, steam generators and chemistry (secondary_update_source).  Nothing flows the other
// way: turbine and condenser are pure sinks of the step's dataflow — their outputs reach only the report columns,
// observation and reward (electrical power, efficiency, condenser pressure, heat rejection, `_last_load_factor`,
// none of which the dynamics read back: sim.py:429-498, systems/secondary/__init__.py:565-932).
struct SecHandoff {
    TurbineInlet inlet;
    double load_demand, cooling_water_temperature;
    double primary_thermal, total_heat_transfer, total_steam_flow, avg_p;
    double fw_total_flow, fw_total_power;
    bool emit_outputs;
};

// chemistry: systems/secondary/__init__.py:634-665 (reads neither turbine nor condenser)
NPS_HD void secondary_update_chemistry(PlantState& st, double dt, const StepInput& in) {
    const MakeupWater mk = {7.2, 100.0, 300.0, 30.0, 8.0};
    NPS_PREFETCH_SELF(st.ph);
    wc_update(st.wc_main, true, mk, 0.02, dt);
    ph_control_update(st.ph, st.wc_main.ph, dt, in.z_ph, in.u_ph, in.emit_outputs);
    wc_queue_effects(st.wc_main, st.ph.ph_setpoint, st.ph.ammonia_dose_rate, st.ph.morpholine_dose_rate);
}

// systems/secondary/__init__.py:340-563 and :629-665 — everything of update_system that does not involve the turbine or
// the condenser.  The chemistry block (:634-665) reads neither, so running it before them is the same arithmetic.
// CHEMISTRY = false leaves the chemistry block to the caller (the one-thread-per-plant step runs it after the sink half,
// in the reference's textual order: measured 1-2 % faster there, profiles/r02_ab_memo_chemistry.txt).
template <bool CHEMISTRY>
NPS_HD void secondary_update_source(PlantState& st, const PlantParams& p, const PrimaryConditions& pc, double load_demand,
                                    double cooling_water_temp, double dt, const StepInput& in, SecHandoff& h) {
    SecondaryState& S = st.sec;
    NPS_TOUCH(S.has_previous_feedwater_temp); NPS_TOUCH(S.previous_feedwater_temp); NPS_TOUCH(S.has_previous_sg_conditions); NPS_TOUCH(S.operating_hours);
    for (int i = 0; i < 3; ++i) { NPS_TOUCH(S.prev_sg_levels[i]); NPS_TOUCH(S.prev_sg_steam_flows[i]); NPS_TOUCH(S.prev_sg_steam_qualities[i]); }
    S.load_demand = load_demand;
    S.feedwater_temperature = 227.0;
    S.cooling_water_temperature = cooling_water_temp;

    // feedwater-temperature smoothing: :384-398
    if (!is_true(S.has_previous_feedwater_temp)) { S.has_previous_feedwater_temp = 1.0; S.previous_feedwater_temp = S.feedwater_temperature; }
    const double estimated_fw_temp = 40.0 + 187.0;
    const double alpha = 0.1;
    double actual_fw_temp = (alpha * estimated_fw_temp + (1 - alpha) * S.previous_feedwater_temp);
    S.previous_feedwater_temp = actual_fw_temp;

    // thermal power / load fraction: :420-440 (sg_i_thermal_power keys take precedence)
    double total_thermal_mw = 0.0 + pc.thermal_power[0]; total_thermal_mw += pc.thermal_power[1]; total_thermal_mw += pc.thermal_power[2];
    double load_fraction = py_min(1.0, total_thermal_mw / 3000.0);
    load_fraction = py_max(load_fraction, 0.2);
    double est_flow_per_sg = 555.0 * load_fraction;
    double est_total_flow = est_flow_per_sg * 3;
    if (!is_true(S.has_previous_sg_conditions)) {
        S.has_previous_sg_conditions = 1.0;
        for (int i = 0; i < 3; ++i) {
            S.prev_sg_levels[i] = 12.5; S.prev_sg_pressures[i] = 6.895;
            S.prev_sg_steam_flows[i] = est_flow_per_sg; S.prev_sg_steam_qualities[i] = 0.99;
        }
    }
    // STEP 1 feedwater: :445-491
    FeedwaterResult fwr;
    feedwater_update(st.fw, st.wc_main, p, S.prev_sg_levels, S.prev_sg_steam_flows, S.prev_sg_steam_qualities,
                     est_total_flow, 40.0, 0.5, 7.4, dt, fwr, &st.sgs.sg[0], in.emit_outputs ? &st.rep : nullptr);
    if (in.emit_outputs) {   // the SG conditions the feedwater system was given: feedwater/physics.py:685-688,1142-1145
        ReportState& R = st.rep;
        // np.mean of 3 values: add.reduce seeds the accumulator with element 0 and adds the pairwise sum of the rest
        R.fw_avg_sg_level = (S.prev_sg_levels[0] + (S.prev_sg_levels[1] + S.prev_sg_levels[2])) / 3.0;
        R.fw_avg_sg_pressure = (S.prev_sg_pressures[0] + (S.prev_sg_pressures[1] + S.prev_sg_pressures[2])) / 3.0;
        R.fw_total_steam_flow = ((0.0 + S.prev_sg_steam_flows[0]) + S.prev_sg_steam_flows[1]) + S.prev_sg_steam_flows[2];
        R.fw_avg_steam_quality = (S.prev_sg_steam_qualities[0] + (S.prev_sg_steam_qualities[1] + S.prev_sg_steam_qualities[2])) / 3.0;
    }
    // STEP 2 steam generators: :493-535 ('sg_i' keys are absent from sg_flow_distribution -> equal split)
    double fw_flows[3];
    for (int i = 0; i < 3; ++i) fw_flows[i] = fwr.total_flow_rate / 3;
    sg_system_update(st.sgs, p, pc.inlet_temp, pc.outlet_temp, pc.flow, load_fraction, S.load_demand / 100.0,
                     actual_fw_temp, fw_flows, dt * 60, nullptr);
    for (int i = 0; i < 3; ++i) {
        S.prev_sg_levels[i] = st.sgs.sg[i].water_level;
        S.prev_sg_pressures[i] = st.sgs.sg[i].secondary_pressure;
        S.prev_sg_steam_flows[i] = st.sgs.sg[i].steam_flow_rate;
        S.prev_sg_steam_qualities[i] = st.sgs.sg[i].steam_quality;
    }
    const double total_heat_transfer = st.sgs.total_thermal_power;
    const double avg_p = st.sgs.average_steam_pressure;
    double avg_t = st.sgs.average_steam_temperature;
    avg_t = py_max(avg_t, secondary_sat_temp(avg_p));
    double total_steam_flow = 0.0 + st.sgs.sg[0].steam_flow_rate; total_steam_flow += st.sgs.sg[1].steam_flow_rate;
    total_steam_flow += st.sgs.sg[2].steam_flow_rate;

    S.total_feedwater_flow = fwr.total_flow_rate;
    S.operating_hours += dt / 3600.0;

    if (CHEMISTRY) secondary_update_chemistry(st, dt, in);

    S.total_steam_flow = total_steam_flow;
    S.total_heat_transfer = total_heat_transfer;
    S.sg_avg_pressure = avg_p;
    S.sg_avg_temperature = avg_t;

    double primary_thermal = 0.0;
    for (int i = 0; i < 3; ++i) primary_thermal += pc.thermal_power[i];
    h.inlet = turbine_inlet_from(st.sgs);
    h.load_demand = S.load_demand;
    h.cooling_water_temperature = S.cooling_water_temperature;
    h.primary_thermal = primary_thermal;
    h.total_heat_transfer = total_heat_transfer;
    h.total_steam_flow = total_steam_flow;
    h.avg_p = avg_p;
    h.fw_total_flow = fwr.total_flow_rate;
    h.fw_total_power = fwr.total_power_consumption;
    h.emit_outputs = in.emit_outputs;
}

// systems/secondary/__init__.py:565-627 and :759-932 — turbine, LP-6 exhaust quality, condenser, energy bookkeeping and
// electrical-power gating, heat-flow report; plus the one line of _apply_secondary_to_primary_feedback (sim.py:492)
// that needs the electrical output.  Touches st.turb, st.cond, six fields of st.sec, st.sim.last_load_factor and the
// hf_* report fields — nothing the source half reads.
NPS_HD void secondary_update_sink(PlantState& st, const PlantParams& p, const SecHandoff& h, double dt) {
    SecondaryState& S = st.sec;
    const double cooling_water_flow = 45000.0;
    // STEP 5 turbine: :565-570
    TurbineResult tr;
    turbine_update(st.turb, p, h.inlet, h.load_demand, 0.007, dt / 60.0, tr, &st.cond, h.emit_outputs);

    // LP-6 exhaust quality: :591-607
    double lp_quality = 0.90;
    {
        double h_f = cond_h_f(tr.condenser_pressure), h_g = cond_h_g(tr.condenser_pressure);
        double h_fg = h_g - h_f;
        if (h_fg > 0) {
            lp_quality = (tr.lp6_outlet_enthalpy - h_f) / h_fg;
            lp_quality = py_max(0.0, py_min(1.0, lp_quality));
        }
    }
    NPS_PREFETCH_SELF(st.cond);
    NPS_PREFETCH_SELF(st.cond.ejector[0]);
    NPS_PREFETCH_SELF(st.cond.ejector[1]);
    CondenserResult cr;
    condenser_update(st.cond, p, tr.condenser_pressure, tr.condenser_temperature, tr.effective_steam_flow, lp_quality,
                     cooling_water_flow, h.cooling_water_temperature, 1.2, 185.0, dt / 60.0, cr);

    // energy bookkeeping and electrical-power gating: :759-932
    const double primary_thermal = h.primary_thermal;
    const double thermal_power_mw = primary_thermal;
    const double total_heat_transfer = h.total_heat_transfer;
    const double turbine_electrical = tr.electrical_power_net;
    S.total_system_heat_rejection = (primary_thermal - turbine_electrical) * 1e6;
    const double actual_fw_flow = h.fw_total_flow;
    double factor = 1.0;
    if (actual_fw_flow < 300.0) factor = 0.0;
    if (factor > 0.0) {
        if (h.total_steam_flow < (300.0 * 0.5)) factor *= 0.1;
        if (h.avg_p < (1.0 * 0.5)) factor *= 0.1;
        if (thermal_power_mw > (primary_thermal * 1.1)) factor = 0.0;
    }
    S.power_reduction_factor = factor;
    S.electrical_power_output = turbine_electrical * factor;
    S.thermal_efficiency = (primary_thermal > 0) ? S.electrical_power_output / primary_thermal : 0.0;
    S.heat_rate_kj_kwh = (S.electrical_power_output > 0)
                             ? (total_heat_transfer / 1000.0) / (S.electrical_power_output * 1000.0) * 3600.0 : 0.0;
    S.condenser_pressure = cr.condenser_pressure;
    st.sim.last_load_factor = S.electrical_power_output / 1100.0;   // sim.py:446,492

    if (h.emit_outputs) {   // HeatFlowTracker: :680-744 and heat_flow_tracker.py:248-327 (all MW)
        ReportState& R = st.rep;
        const double sg_in = total_heat_transfer / 1e6;
        const double steam_out = total_heat_transfer * 0.98 / 1e6;
        const double sg_losses = total_heat_transfer * 0.02 / 1e6;
        const double mech = tr.mechanical_power;
        const double turb_losses = mech * 0.05;
        const double pump_work = h.fw_total_power;
        const double fw_losses = h.fw_total_power * 0.1;
        const double total_losses = (sg_losses + turb_losses + fw_losses);
        const double rejection = sg_in - mech - total_losses;
        const double cond_losses = rejection * 0.01;
        const double gen_out = mech * 0.985;
        const double gen_losses = mech - gen_out;
        const double aux = gen_out * 0.02;
        const double net = gen_out - aux;
        const double e_in = (sg_in + pump_work);
        const double e_out = (net + rejection + sg_losses + turb_losses + cond_losses + fw_losses + gen_losses + 0.0);
        R.hf_steam_enthalpy_flow = steam_out;
        R.hf_turbine_work_output = mech;
        R.hf_condenser_heat_rejection = rejection;
        R.hf_net_electrical_output = net;
        R.hf_energy_balance_error = e_in - e_out;
        if (e_in > 0) {
            R.hf_energy_balance_percent = ((e_in - e_out) / e_in) * 100.0;
            R.hf_overall_efficiency = net / e_in;
        } else {
            R.hf_energy_balance_percent = 0.0;
            R.hf_overall_efficiency = 0.0;   // a fresh HeatFlowState() every step: heat_flow_tracker.py:255
        }
    }
}

}  // namespace nps
