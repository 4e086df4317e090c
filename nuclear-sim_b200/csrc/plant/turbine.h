// Turbine: bearing-lubrication wrapper (runs first, with default operating conditions), 14-stage
// serial steam expansion with extraction and per-stage degradation, rotor dynamics with four
// bearings and vibration response, metal-temperature tracker and protection trips.
// Restates the wrapped EnhancedTurbinePhysics.update_state
// (reference: nuclear_simulator/systems/secondary/turbine/turbine_bearing_lubrication.py:715-782
//  around turbine/enhanced_physics.py:694-890) and callees.
#pragma once
#include "hd.h"
#include "state.h"
#include "prefetch.h"
#include "lubrication.h"
#include "sg.h"

namespace nps {

// Antoine saturation temperature shared by stage_system.py:458-466 and enhanced_physics.py:1295-1303
NPS_HD_SHARED double turb_sat_temp(double p_mpa) {
    if (p_mpa <= 0.001) return 10.0;
    double p_bar = np_clip(p_mpa * 10.0, 0.01, 100.0);
    double t = 1730.63 / (8.07131 - nps_log10(p_bar)) - 233.426;
    return np_clip(t, 10.0, 374.0);
}
NPS_HD_SHARED double turb_h_g(double p_mpa) {   // stage_system.py:468-473
    double temp = turb_sat_temp(p_mpa);
    double h_f = 4.18 * temp;
    double h_fg = 2257.0 * py_pow(1.0 - temp / 374.0, 0.38);
    return h_f + h_fg;
}
// Saturation properties are pure functions of pressure, and the 14-stage chain asks for them at the same pressures
// again and again (a stage's outlet pressure is the next stage's inlet pressure; enthalpy, entropy and
// enthalpy->temperature all start from T_sat(p)): 117 log10 and 55 pow(., 0.38) per plant-step in the reference.  A
// two-entry memo keyed on the EXACT pressure bits returns the value the function would recompute, so results are
// unchanged bit for bit; it only removes the repeats.
struct TurbSatMemo {
    double p[2], sat[2], hg[2];
    int has_hg[2];
    int next;
};
NPS_HD void turb_memo_init(TurbSatMemo& m) { m.p[0] = m.p[1] = -1.0; m.has_hg[0] = m.has_hg[1] = 0; m.next = 0; }
NPS_HD int turb_memo_slot(TurbSatMemo& m, double p_mpa) {
    if (m.p[0] == p_mpa) return 0;
    if (m.p[1] == p_mpa) return 1;
    const int s = m.next;
    m.next = 1 - s;
    m.p[s] = p_mpa; m.sat[s] = turb_sat_temp(p_mpa); m.has_hg[s] = 0;
    return s;
}
NPS_HD double turb_sat_temp_m(TurbSatMemo& m, double p_mpa) { return m.sat[turb_memo_slot(m, p_mpa)]; }
NPS_HD double turb_h_g_m(TurbSatMemo& m, double p_mpa) {   // == turb_h_g(p_mpa)
    const int s = turb_memo_slot(m, p_mpa);
    if (!m.has_hg[s]) {
        double temp = m.sat[s];
        double h_f = 4.18 * temp;
        double h_fg = 2257.0 * py_pow(1.0 - temp / 374.0, 0.38);
        m.hg[s] = h_f + h_fg;
        m.has_hg[s] = 1;
    }
    return m.hg[s];
}
// TurbineStage._steam_enthalpy: stage_system.py:418-443
NPS_HD double stage_steam_enthalpy(double t, double p_mpa, TurbSatMemo& m) {
    p_mpa = py_max(0.001, py_min(p_mpa, 22.0));
    t = py_max(0.0, py_min(t, 800.0));
    double sat = turb_sat_temp_m(m, p_mpa);
    if (t <= sat) return turb_h_g_m(m, p_mpa);
    double h_g = turb_h_g_m(m, p_mpa);
    double superheat = t - sat;
    double cp = (p_mpa > 10.0) ? 2.5 : ((p_mpa > 1.0) ? 2.2 : 2.0);
    return h_g + cp * superheat;
}
NPS_HD double stage_steam_enthalpy(double t, double p_mpa) { TurbSatMemo m; turb_memo_init(m); return stage_steam_enthalpy(t, p_mpa, m); }
// TurbineStage._steam_entropy: stage_system.py:445-456
NPS_HD double stage_steam_entropy(double t, double p_mpa, TurbSatMemo& m) {
    double sat = turb_sat_temp_m(m, p_mpa);
    double s_f = 4.18 * nps_log((sat + 273.15) / 273.15);
    double s_fg = 2257.0 / (sat + 273.15);
    double s_g = s_f + s_fg;
    if (t > sat) return s_g + 2.1 * nps_log((t + 273.15) / (sat + 273.15));
    return s_g;
}
// EnhancedTurbinePhysics._steam_enthalpy: enhanced_physics.py:1285-1293
NPS_HD double turbine_steam_enthalpy(double t, double p_mpa) {
    double sat = turb_sat_temp(p_mpa);
    if (t <= sat) return turb_h_g(p_mpa);
    double h_g = turb_h_g(p_mpa);
    return h_g + 2.1 * (t - sat);
}

// Turbine lubrication component order: hp_journal_bearing, lp_journal_bearing, thrust_bearing,
// seal_oil_system, oil_coolers (turbine_bearing_lubrication.py:105-186)
enum { TBL_HP = 0, TBL_LP = 1, TBL_THRUST = 2, TBL_SEAL = 3, TBL_COOLERS = 4, TBL_NCOMP = 5 };
NPS_HD LubComponent turb_lub_component(int i) {
    switch (i) {
        case TBL_HP:     return {0.0003, 1.8, 1.5, 3.0, 0.02, 0.5, 8.0, 20.0};
        case TBL_LP:     return {0.0004, 1.6, 1.5, 2.8, 0.018, 0.45, 10.0, 25.0};
        case TBL_THRUST: return {0.0006, 2.5, 1.3, 4.0, 0.03, 0.7, 5.0, 15.0};
        case TBL_SEAL:   return {0.0008, 1.4, 1.1, 3.5, 0.025, 0.6, 6.0, 18.0};
        default:         return {0.0001, 1.0, 0.5, 1.5, 0.01, 0.2, 20.0, 40.0};
    }
}
NPS_HD double turb_lub_oil_flow_requirement(int i) {
    const double f[5] = {25.0, 30.0, 40.0, 15.0, 100.0};
    return f[i];
}

// update_with_lubrication pre-step: turbine_bearing_lubrication.py:715-782 (+ :800-1128).
// All keyword arguments of the wrapped call are unknown to the wrapper, so it uses its defaults
// (steam_quality 0.99) and takes load_factor from the PREVIOUS turbine.load_demand (percent).
NPS_HD void turbine_lubrication_prestep(TurbineState& T, const PlantParams& p, double dt) {
    const double load_factor = T.load_demand;
    const double rotor_speed = T.rotor_speed;
    double friction_heat[4], b_load_factor[4], b_temp[4];
    double total_heat = 0.0;
    const double omega = rotor_speed * 2 * NPS_PI / 60.0;
    for (int b = 0; b < 4; ++b) {   // collect_bearing_states / calculate_bearing_friction_heat
        double load_n = T.bearing[b].current_load * 1000.0;
        double clearance_m = p.rd_bearing_clearance / 1000.0;
        double fp = load_n * p.rd_friction_coefficient * omega * clearance_m;
        friction_heat[b] = py_max(0.0, fp);
        b_load_factor[b] = T.bearing[b].current_load / py_max(p.rd_design_load_capacity, 1.0);
        b_temp[b] = T.bearing[b].metal_temperature;
        total_heat += friction_heat[b];
    }
    const double speed_factor = rotor_speed / 3600.0;
    // update_lubrication_with_feedback: :866-929
    double base_oil_temp = 40.0 + load_factor * 15.0;
    double system_oil_temp;
    if (total_heat > 0) {
        double oil_mass_flow = 100.0 / 60.0 * 0.85;
        system_oil_temp = base_oil_temp + total_heat / (oil_mass_flow * 2000.0);
    } else {
        system_oil_temp = base_oil_temp;
    }
    const double steam_quality = 0.99;
    double contamination_input = load_factor * 0.02 + (1.0 - steam_quality) * 0.5;
    double moisture_input = (1.0 - steam_quality) * 0.01;
    const LubLimits lim = {p.tl_contamination_limit, p.tl_acidity_limit, p.tl_moisture_limit, p.tl_viscosity_change_limit};
    lub_update_oil_quality(T.lub, TBL_NCOMP, lim, system_oil_temp, contamination_input, moisture_input, dt);
    // update_component_wear with TurbineBearingLubricationSystem.calculate_component_wear (:261-329)
    for (int c = 0; c < TBL_NCOMP; ++c) {
        const LubComponent k = turb_lub_component(c);
        double rate;
        if (c == TBL_HP) {
            double stf = py_max(1.0, (b_temp[0] - 70.0) / 20.0);
            double lfa = b_load_factor[0] * 1.2;
            rate = (k.base_wear_rate * py_pow(lfa, k.load_wear_exponent) * py_pow(speed_factor, k.speed_wear_exponent) * stf);
        } else if (c == TBL_LP) {
            double mf = py_max(1.0, (1.0 - 0.99) * 10.0);
            double tf = py_max(1.0, (b_temp[1] - 60.0) / 25.0);
            rate = (k.base_wear_rate * py_pow(b_load_factor[1], k.load_wear_exponent) *
                    py_pow(speed_factor, k.speed_wear_exponent) * mf * tf);
        } else if (c == TBL_THRUST) {
            double axial = b_load_factor[2] * 1.0;
            double tf = py_max(1.0, (b_temp[2] - 50.0) / 30.0);
            rate = (k.base_wear_rate * py_pow(axial, k.load_wear_exponent) * py_pow(speed_factor, k.speed_wear_exponent) * tf);
        } else if (c == TBL_SEAL) {
            double cf = 1.0 + T.lub.oil_contamination_level / 10.0;
            rate = (k.base_wear_rate * py_pow(1.0, k.load_wear_exponent) * cf);
        } else {
            rate = (k.base_wear_rate * 1.0 * 1.0);
        }
        lub_apply_component_wear(T.lub, c, k, rate, dt);
    }
    lub_update_health(T.lub, TBL_NCOMP);
    // calculate_component_oil_temperatures (:967-1071) + inject_lubrication_into_bearings (:1073-1128)
    for (int b = 0; b < 4; ++b) {
        double oil_flow_lpm = turb_lub_oil_flow_requirement(b);
        double heat = friction_heat[b];
        double ct;
        if (oil_flow_lpm > 0 && heat > 0) {
            double mflow = oil_flow_lpm / 60.0 * 0.85;
            double rise_basic = heat / (mflow * 2000.0);
            double dissipated = 50.0 * 0.5 * py_max(0.0, rise_basic);
            double net = py_max(0.0, heat - dissipated);
            double rise = (net > 0) ? net / (mflow * 2000.0) : 0.0;
            double max_rise = (b == 0) ? 25.0 : ((b == 2) ? 20.0 : ((b == 3) ? 15.0 : 20.0));
            rise = py_min(max_rise, py_max(0.0, rise));
            ct = system_oil_temp + rise;
            if (b == 0) ct += 2.0; else if (b == 2) ct += 1.0; else if (b == 3) ct -= 5.0;
            ct = py_max(35.0, py_min(70.0, ct));
        } else {
            if (b == 0) ct = system_oil_temp + 2.0;
            else if (b == 2) ct = system_oil_temp + 1.0;
            else if (b == 3) ct = system_oil_temp - 5.0;
            else ct = system_oil_temp;
            ct = py_max(35.0, py_min(70.0, ct));
        }
        TurbineBearingState& B = T.bearing[b];
        // BearingModel.set_lubrication_state: rotor_dynamics.py:132-155
        if (isfinite(ct) && ct > 0) { B.oil_temperature = ct; B.external_oil_temp = 1.0; }
        B.oil_flow_rate = oil_flow_lpm;
        if (isfinite(T.lub.oil_contamination_level)) B.oil_contamination_level = T.lub.oil_contamination_level;
        B.efficiency_factor = py_min(B.efficiency_factor, T.lub.lubrication_effectiveness);
    }
}

// The seven members of a stage that are carried from one step to the next (everything else of the stage is written
// before it is read).  The stage loop fetches the NEXT stage's copy while it works on the current one (NPS_STAGE_AHEAD):
// a stage's first touches then wait behind ~1 300 instructions of the previous stage instead of in front of its own.
struct StageCarried {
    double actual_efficiency, blade_condition_factor, fouling_factor, blade_wear_factor, deposit_thickness, operating_hours,
           efficiency_degradation;
};
NPS_HD StageCarried stage_carried(const TurbineStageState& s) {
#if defined(__CUDA_ARCH__) && !defined(NPS_NO_STAGE_AHEAD)
    const volatile TurbineStageState& v = s;      // volatile: issued here, not sunk to the first use
    return StageCarried{v.actual_efficiency, v.blade_condition_factor, v.fouling_factor, v.blade_wear_factor, v.deposit_thickness,
                        v.operating_hours, v.efficiency_degradation};
#else
    return StageCarried{s.actual_efficiency, s.blade_condition_factor, s.fouling_factor, s.blade_wear_factor, s.deposit_thickness,
                        s.operating_hours, s.efficiency_degradation};
#endif
}

// TurbineStage.calculate_stage_expansion: stage_system.py:98-292
NPS_HD void stage_expand(TurbineStageState& s, const StageCarried& c, const PlantParams& p, int k, double inlet_pressure,
                         double inlet_temperature, double inlet_flow, double outlet_pressure, double extraction_demand,
                         TurbSatMemo& memo, bool emit_outputs = true) {
    s.inlet_pressure = inlet_pressure;
    s.inlet_temperature = inlet_temperature;
    s.inlet_flow = inlet_flow;
    const double d_in = p.ts_design_inlet_pressure[k], d_out = p.ts_design_outlet_pressure[k], d_flow = p.ts_design_steam_flow[k];
    double design_ratio = d_out / d_in;
    double load_factor = (d_flow > 0) ? s.inlet_flow / d_flow : 1.0;
    load_factor = np_clip(load_factor, 0.3, 1.5);
    double temp_factor = (inlet_temperature + 273.15) / (285.8 + 273.15);
    temp_factor = np_clip(temp_factor, 0.8, 1.2);
    double load_adj = 0.9 + 0.2 * load_factor;
    double temp_adj = 0.95 + 0.1 * (temp_factor - 1.0);
    double adj_ratio = design_ratio * load_adj * temp_adj;
    adj_ratio = np_clip(adj_ratio, design_ratio * 0.85, design_ratio * 1.15);
    double physics_outlet = inlet_pressure * adj_ratio;
    if (outlet_pressure >= inlet_pressure) {
        s.outlet_pressure = physics_outlet;
    } else {
        double min_allowed, max_allowed;
        if (k == 13) { min_allowed = 0.002; max_allowed = 0.009; }   // stage_id == "LP-6"
        else { min_allowed = inlet_pressure * (design_ratio * 0.7); max_allowed = inlet_pressure * (design_ratio * 1.3); }
        if (outlet_pressure < min_allowed) s.outlet_pressure = min_allowed;
        else if (outlet_pressure > max_allowed) s.outlet_pressure = max_allowed;
        else s.outlet_pressure = outlet_pressure;
    }
    s.inlet_enthalpy = stage_steam_enthalpy(inlet_temperature, inlet_pressure, memo);
    if (emit_outputs) s.inlet_entropy = stage_steam_entropy(inlet_temperature, inlet_pressure, memo);   // logged only
    if (is_true(p.ts_has_extraction[k]) && extraction_demand > 0) {
        s.extraction_flow = np_clip(extraction_demand, p.ts_min_extraction_flow[k],
                                    py_min(p.ts_max_extraction_flow[k], inlet_flow * 0.3));
        s.extraction_pressure = inlet_pressure * 0.7 + outlet_pressure * (1 - 0.7);
        TurbSatMemo em; turb_memo_init(em);   // extraction pressure: its own lookups, leaves the chain's entries alone
        double et = turb_sat_temp_m(em, s.extraction_pressure);
        s.extraction_enthalpy = stage_steam_enthalpy(et, s.extraction_pressure, em);
    } else {
        s.extraction_flow = 0.0;
    }
    s.outlet_flow = s.inlet_flow - s.extraction_flow;
    double pr = s.outlet_pressure / inlet_pressure;
    double t_isen = (inlet_temperature + 273.15) * py_pow(pr, 0.25) - 273.15;
    double h_isen = stage_steam_enthalpy(t_isen, s.outlet_pressure, memo);
    const double quality_factor = 1.0;   // steam_quality hard-coded 0.99 at stage_system.py:217
    double total_eff = (c.actual_efficiency * c.blade_condition_factor * c.fouling_factor * c.blade_wear_factor * quality_factor);
    double isen_drop = s.inlet_enthalpy - h_isen;
    if (isen_drop <= 0) {
        double ratio = s.outlet_pressure / inlet_pressure;
        double min_drop = 50.0 * (1.0 - ratio);
        isen_drop = py_max(min_drop, 10.0);
    }
    double actual_drop = total_eff * isen_drop;
    if (actual_drop <= 0) actual_drop = py_max(1.0, isen_drop * 0.5);
    s.enthalpy_drop = actual_drop;
    s.outlet_enthalpy = s.inlet_enthalpy - actual_drop;
    {   // _enthalpy_to_temperature with the REQUESTED outlet pressure (stage_system.py:256)
        double sat = turb_sat_temp_m(memo, outlet_pressure);
        double h_g = turb_h_g_m(memo, outlet_pressure);
        s.outlet_temperature = (s.outlet_enthalpy <= h_g) ? sat : sat + (s.outlet_enthalpy - h_g) / 2.1;
    }
    double main_power = s.outlet_flow * actual_drop / 1000.0;
    if (main_power < 0) main_power = 0.0;
    double ext_power = 0.0;
    if (s.extraction_flow > 0) ext_power = s.extraction_flow * (s.inlet_enthalpy - s.extraction_enthalpy) / 1000.0;
    s.power_output = main_power + ext_power;
    double design_drop = p.ts_design_efficiency[k] * isen_drop;
    s.loading_factor = actual_drop / py_max(1.0, design_drop);
}

// get_dynamic_pressure_ratio closure: stage_system.py:794-873
NPS_HD double stage_dynamic_pressure_ratio(const PlantParams& p, int k, int total, double current_pressure, double inlet_flow) {
    double design_ratio = p.ts_design_outlet_pressure[k] / p.ts_design_inlet_pressure[k];
    double d_flow = p.ts_design_steam_flow[k];
    double lf = (d_flow > 0) ? inlet_flow / d_flow : 1.0;
    lf = np_clip(lf, 0.3, 1.5);
    double d_in = p.ts_design_inlet_pressure[k];
    double pf = (d_in > 0) ? current_pressure / d_in : 1.0;
    pf = np_clip(pf, 0.5, 1.5);
    double la = 0.90 + 0.2 * (lf - 1.0);
    double pa = 0.95 + 0.1 * (pf - 1.0);
    double dyn = design_ratio * la * pa;
    const bool is_lp = (k >= 8);
    double min_ratio = is_lp ? 0.50 : 0.70, max_ratio = is_lp ? 0.85 : 0.95;
    if (is_lp) {
        int remaining = total - k - 1;
        if (remaining > 0) {
            double min_outlet = 0.007 / py_pow(0.85, (double)remaining);
            double max_allowable = min_outlet / current_pressure;
            min_ratio = py_max(min_ratio, max_allowable);
        }
    }
    if (k == 13) {
        double calc = 0.007 / current_pressure;
        dyn = py_max(calc, 0.05);
    } else {
        dyn = np_clip(dyn, min_ratio, max_ratio);
    }
    return dyn;
}

struct TurbineResult {
    double mechanical_power, electrical_power_gross, electrical_power_net, overall_efficiency, steam_rate;
    double hp_power, lp_power, condenser_pressure, condenser_temperature, effective_steam_flow, lp6_outlet_enthalpy;
};

// What the turbine reads from the steam-generator system (EnhancedTurbinePhysics.update_state arguments assembled in
// systems/secondary/__init__.py:537-570): the aggregates, the three SG pressures and the SG-system availability.  A
// plain value record, so the turbine / condenser half of a step can run on a different thread than the half that
// produced it (nps_capi.cu: split launch shape for small batches).
struct TurbineInlet {
    double average_steam_pressure, average_steam_temperature, total_steam_flow, system_availability;
    double sg_pressure[3];
};
NPS_HD TurbineInlet turbine_inlet_from(const SGSystemState& S) {
    TurbineInlet t;
    t.average_steam_pressure = S.average_steam_pressure; t.average_steam_temperature = S.average_steam_temperature;
    t.total_steam_flow = S.total_steam_flow; t.system_availability = S.system_availability;
    for (int i = 0; i < 3; ++i) t.sg_pressure[i] = S.sg[i].secondary_pressure;
    return t;
}

// A bearing record read with loads that are issued where they stand (device: volatile), so that the NEXT bearing's
// record travels while the current one is processed - the stage loop's scheme (StageCarried).
NPS_HD TurbineBearingState bearing_fetch(const TurbineBearingState& b) {
#if defined(__CUDA_ARCH__) && !defined(NPS_NO_STAGE_AHEAD)
    const volatile TurbineBearingState& v = b;
    TurbineBearingState r;
    r.current_load = v.current_load; r.metal_temperature = v.metal_temperature; r.vibration_displacement = v.vibration_displacement;
    r.operating_hours = v.operating_hours; r.wear_factor = v.wear_factor; r.efficiency_factor = v.efficiency_factor;
    r.clearance_increase = v.clearance_increase; r.oil_temperature = v.oil_temperature; r.oil_flow_rate = v.oil_flow_rate;
    r.oil_contamination_level = v.oil_contamination_level; r.external_oil_temp = v.external_oil_temp;
    return r;
#else
    return b;
#endif
}

// Wrapped EnhancedTurbinePhysics.update_state (dt in hours; load_demand as passed = percent)
NPS_HD void turbine_update(TurbineState& T, const PlantParams& p, const TurbineInlet& S, double load_demand,
                           double condenser_pressure, double dt, TurbineResult& out,
                           const CondenserState* prefetch_next = nullptr, bool emit_outputs = true) {
    NPS_PREFETCH(T.stage[0]);
    NPS_PREFETCH_SELF(T);
    for (int b = 0; b < 4; ++b) NPS_PREFETCH_SELF(T.bearing[b]);
    turbine_lubrication_prestep(T, p, dt);

    const double steam_pressure = S.average_steam_pressure, steam_temperature_in = S.average_steam_temperature;
    const double steam_flow = S.total_steam_flow;
    T.load_demand = load_demand;
    // _calculate_pressure_variation_effects: enhanced_physics.py:1312-1350
    double psf;
    {
        double avg = (0.0 + S.sg_pressure[0] + S.sg_pressure[1] + S.sg_pressure[2]) / 3;
        double md = py_max3(fabs(S.sg_pressure[0] - avg), fabs(S.sg_pressure[1] - avg),
                            fabs(S.sg_pressure[2] - avg));
        double vf = md / 0.1;
        if (md < 0.02) psf = 1.0;
        else if (md < 0.05) psf = 1.0 - (md - 0.02) / 0.03 * 0.05;
        else psf = 0.95 - py_min(vf - 0.5, 0.25);
        psf = np_clip(psf, 0.7, 1.0);
    }
    // TurbineStageSystem.update_state: stage_system.py:928-1016 (control logic has no effect on dynamics)
    double total_power = 0.0, total_extraction = 0.0, hp_power = 0.0, lp_power = 0.0;
    {
        double cur_p = steam_pressure, cur_t = steam_temperature_in, cur_f = steam_flow;
        const double final_pressure = 0.007;
        TurbSatMemo memo; turb_memo_init(memo);
        StageCarried nxt = stage_carried(T.stage[0]);
        NPS_UNIT_LOOP
        for (int k = 0; k < 14; ++k) {
            const StageCarried c = nxt;
            if (k < 13) nxt = stage_carried(T.stage[k + 1]);
            double ratio = stage_dynamic_pressure_ratio(p, k, 14, cur_p, steam_flow);
            double outp = cur_p * ratio;
            outp = py_max(outp, final_pressure);
            int remaining = 14 - k - 1;
            if (remaining == 0) outp = final_pressure;
            else if (remaining == 1) outp = py_max(outp, final_pressure / 0.5);
            if (outp >= cur_p) { outp = cur_p * 0.95; outp = py_max(outp, final_pressure); }
            double ed = 0.0;
            if (k == 2) ed = 25.0 * load_demand; else if (k == 3) ed = 30.0 * load_demand;
            else if (k == 4) ed = 20.0 * load_demand; else if (k == 8) ed = 15.0 * load_demand;
            else if (k == 9) ed = 10.0 * load_demand;
            stage_expand(T.stage[k], c, p, k, cur_p, cur_t, cur_f, outp, ed, memo, emit_outputs);
            total_power += T.stage[k].power_output;
            total_extraction += T.stage[k].extraction_flow;
            if (k < 8) hp_power += T.stage[k].power_output; else lp_power += T.stage[k].power_output;
            cur_p = T.stage[k].outlet_pressure; cur_t = T.stage[k].outlet_temperature; cur_f = T.stage[k].outlet_flow;
            // TurbineStage.update_degradation: stage_system.py:294-339.  The reference runs it in a second pass over
            // the 14 stages (stage_system.py:988-990); a stage's degradation touches only that stage's own
            // factors, which no later stage reads, so doing it here is the same arithmetic on the same values while
            // the stage record is still on chip.
            TurbineStageState& s = T.stage[k];
            s.efficiency_degradation = c.efficiency_degradation + p.ts_fouling_rate * dt;
            s.deposit_thickness = c.deposit_thickness + p.ts_deposit_buildup_rate * dt;
            s.fouling_factor = 1.0 / (1.0 + s.deposit_thickness / 0.5);
            double wear_inc = p.ts_erosion_rate * dt;
            double blade_wear = wear_inc * py_pow(s.loading_factor, 2.0);
            s.blade_wear_factor = py_max(0.7, c.blade_wear_factor - blade_wear);
            s.blade_condition_factor = py_min(s.fouling_factor, s.blade_wear_factor);
            s.actual_efficiency = py_max(0.7, p.ts_design_efficiency[k] - s.efficiency_degradation);
            s.operating_hours = c.operating_hours + dt;
        }
        if (prefetch_next) {
            NPS_PREFETCH_FAR(*prefetch_next);
            NPS_PREFETCH_FAR(prefetch_next->ejector[0]);
            NPS_PREFETCH_FAR(prefetch_next->ejector[1]);
        }
        T.ss_total_power_output = total_power * psf;
        T.ss_total_steam_flow = steam_flow;
        T.ss_total_extraction_flow = total_extraction;
        T.ss_system_efficiency = py_min(1.0, T.ss_system_efficiency * psf);
        if (steam_flow > 0) {
            double h_out = stage_steam_enthalpy(cur_t, cur_p, memo);     // last outlet pressure: still in the memo
            double h_in = stage_steam_enthalpy(steam_temperature_in, steam_pressure, memo);
            T.ss_overall_efficiency = (h_in - h_out) / h_in;
        } else {
            T.ss_overall_efficiency = 0.0;
        }
        T.ss_operating_hours += dt;
    }
    const double stage_power_mw = T.ss_total_power_output;
    const double applied_torque = stage_power_mw * 1e6 / (2 * NPS_PI * 3600 / 60);

    // RotorDynamicsModel.update_state: rotor_dynamics.py:956-1070
    {
        NPS_TOUCH(T.rotor_speed); NPS_TOUCH(T.overspeed_events); NPS_TOUCH(T.rotor_temperature); NPS_TOUCH(T.thermal_bow); NPS_TOUCH(T.rotor_operating_hours); NPS_TOUCH(T.prot_trip_reasons); NPS_TOUCH(T.prot_timer_overspeed); NPS_TOUCH(T.prot_timer_vibration); NPS_TOUCH(T.prot_timer_bearing_temp); NPS_TOUCH(T.operating_hours);
        const double dt_seconds = dt * 3600.0;
        double total_friction = 0.0;
        for (int b = 0; b < 4; ++b)
            total_friction += (T.bearing[b].current_load * 1000.0 * p.rd_friction_coefficient * p.rd_bearing_clearance / 1000.0);
        T.friction_torque = total_friction;
        T.net_torque = applied_torque - T.friction_torque;
        double ang_acc = T.net_torque / p.rd_rotor_inertia;
        T.rotor_acceleration = ang_acc * 60.0 / (2 * NPS_PI);
        T.rotor_speed += T.rotor_acceleration * dt_seconds;
        T.rotor_speed = py_max(0.0, py_min(T.rotor_speed, p.rd_max_speed));
        if (T.rotor_speed > p.rd_max_speed * 0.99) T.overspeed_events += 1.0;
        // calculate_thermal_effects(steam_temperature, 25.0, dt)
        double temp_change = (steam_temperature_in - T.rotor_temperature) / 2.0 * dt;
        T.rotor_temperature += temp_change;
        T.thermal_expansion = ((T.rotor_temperature - 25.0) * p.rd_thermal_expansion_coefficient * p.rd_rotor_length * 1000.0);
        if (T.rotor_speed < 100.0) {
            double grad = (dt > 0) ? fabs(temp_change) / dt : 0.0;
            T.thermal_bow = py_min(p.rd_thermal_bow_limit, T.thermal_bow + grad * 0.001 * dt);
        } else {
            T.thermal_bow *= 0.95;
        }
        const double weight_per_bearing = p.rd_rotor_mass * 9.81 / 1000.0 / 4;
        const double steam_thrust = (100.0 * load_demand) / 4;
        TurbineBearingState nxtB = bearing_fetch(T.bearing[0]);
        NPS_UNIT_LOOP
        for (int b = 0; b < 4; ++b) {
            TurbineBearingState B = nxtB;           // 11 fields, fetched while the previous bearing was processed; written back after the block
            if (b < 3) nxtB = bearing_fetch(T.bearing[b + 1]);
            // calculate_bearing_loads: rotor_dynamics.py:83-130 (TB-003 is the thrust bearing)
            double thrust_load = (b == 2) ? steam_thrust : 0.0;
            double thermal_load = fabs(T.thermal_expansion) * p.rd_bearing_stiffness / 1000.0;
            double unbalance = py_pow(T.rotor_speed / 3600.0, 2.0) * 0.1;
            double total_load = weight_per_bearing + thrust_load + thermal_load + unbalance;
            total_load *= (2.0 - B.wear_factor);
            B.current_load = total_load;
            // calculate_bearing_temperature(40.0, load, speed, dt): rotor_dynamics.py:157-270
            double oil_in = 40.0;
            double load = isfinite(B.current_load) ? py_max(0.0, B.current_load) : 0.0;
            double speed = isfinite(T.rotor_speed) ? py_max(0.0, T.rotor_speed) : 0.0;
            double ang = speed * 2 * NPS_PI / 60.0;
            double f_torque = p.rd_friction_coefficient * (load * 1000.0) * 0.15;
            double f_power = py_min(f_torque * ang, 50000.0);
            if (!isfinite(f_power) || f_power < 0) f_power = 0.0;
            if (!isfinite(B.oil_flow_rate) || B.oil_flow_rate <= 0) B.oil_flow_rate = 10.0;
            double mflow = B.oil_flow_rate / 60.0 * 850.0 / 1000.0;
            double rise = 0.0;
            if (mflow > 0 && isfinite(mflow)) rise = py_min(50.0, py_max(0.0, f_power / (mflow * 2000.0)));
            if (!isfinite(rise)) rise = 0.0;
            if (!is_true(B.external_oil_temp)) {
                double target = py_max(20.0, py_min(150.0, oil_in + rise));
                if (!isfinite(B.oil_temperature)) B.oil_temperature = oil_in;
                double tc = (target - B.oil_temperature) / 30.0 * dt * 3600.0;
                tc = py_max(-10.0, py_min(10.0, tc));
                B.oil_temperature += tc;
                B.oil_temperature = py_max(20.0, py_min(150.0, B.oil_temperature));
            }
            B.metal_temperature = py_max(30.0, py_min(200.0, oil_in + rise * 1.5));
            // update_bearing_wear(load, 5.0, dt): rotor_dynamics.py:272-315
            double lfac = B.current_load / p.rd_design_load_capacity;
            double load_wear = 0.00001 * py_pow(lfac, 2.0) * dt;
            double cont_wear = 0.000005 * 5.0 * dt;
            double tw = load_wear + cont_wear;
            B.wear_factor = py_max(0.5, B.wear_factor - tw);
            B.clearance_increase += tw * 0.01;
            B.efficiency_factor = B.wear_factor * 0.9 + 0.1;
            B.operating_hours += dt;
            T.bearing[b] = B;
        }
        // VibrationMonitor.calculate_vibration_response: rotor_dynamics.py:624-705
        double avg_k = (0.0 + p.rd_bearing_stiffness + p.rd_bearing_stiffness + p.rd_bearing_stiffness + p.rd_bearing_stiffness) / 4;
        double avg_c = (0.0 + p.rd_bearing_damping + p.rd_bearing_damping + p.rd_bearing_damping + p.rd_bearing_damping) / 4;
        double unb = py_pow(T.rotor_speed / 60.0, 2.0) * 0.1;
        double rot_f = T.rotor_speed / 60.0;
        double om = 2 * NPS_PI * rot_f;
        const double rotor_mass = 15000.0;
        double nat_f = sqrt(avg_k / rotor_mass) / (2 * NPS_PI);
        double fr = rot_f / nat_f;
        double crit_d = 2 * sqrt(avg_k * rotor_mass);
        double zeta = avg_c / crit_d;
        double denom = sqrt(py_pow(1 - py_pow(fr, 2.0), 2.0) + py_pow(2 * zeta * fr, 2.0));
        double unb_resp = unb / avg_k / denom;
        double th_resp = T.thermal_bow * py_pow(fr, 2.0) / denom;
        double d1 = (unb_resp + th_resp) * 39.37;
        double v1 = d1 * om / 1000.0;
        double a1 = v1 * om / 9.81;
        double d2 = d1 * 0.1, d3 = d1 * 0.05;
        double td = sqrt(py_pow(d1, 2.0) + py_pow(d2, 2.0) + py_pow(d3, 2.0));
        double tv = sqrt(py_pow(v1, 2.0) + py_pow(v1 * 0.1, 2.0) + py_pow(v1 * 0.05, 2.0));
        double ta = sqrt(py_pow(a1, 2.0) + py_pow(a1 * 0.1, 2.0) + py_pow(a1 * 0.05, 2.0));
        T.vib_displacement_x = td; T.vib_displacement_y = td * 0.8;
        T.vib_velocity_x = tv; T.vib_velocity_y = tv * 0.8;
        T.vib_acceleration_x = ta; T.vib_acceleration_y = ta * 0.8;
        T.vib_harmonic[0] = d1; T.vib_harmonic[1] = d2; T.vib_harmonic[2] = d3;
        // check_critical_speeds / update_alarms: rotor_dynamics.py:707-758
        double m1 = fabs(T.rotor_speed - p.rd_first_critical_speed) / p.rd_first_critical_speed;
        double m2 = fabs(T.rotor_speed - p.rd_second_critical_speed) / p.rd_second_critical_speed;
        T.vib_critical_speed_alarm = as_flag(m1 < p.rd_critical_speed_margin || m2 < p.rd_critical_speed_margin);
        T.vib_displacement_alarm = as_flag(py_max(fabs(T.vib_displacement_x), fabs(T.vib_displacement_y)) > p.rd_displacement_alarm);
        T.vib_velocity_alarm = as_flag(py_max(fabs(T.vib_velocity_x), fabs(T.vib_velocity_y)) > p.rd_velocity_alarm);
        T.vib_acceleration_alarm = as_flag(py_max(fabs(T.vib_acceleration_x), fabs(T.vib_acceleration_y)) > p.rd_acceleration_alarm);
        T.rotor_operating_hours += dt;
    }

    // MetalTemperatureTracker.update_temperatures: enhanced_physics.py:73-165 (14 stage outlet temps)
    {
#if defined(__CUDA_ARCH__)
#pragma unroll
        for (int i = 0; i < 14; ++i) { NPS_TOUCH(T.th_blade_temperatures[i]); if (i < 8) NPS_TOUCH(T.th_rotor_temperatures[i]); if (i < 6) NPS_TOUCH(T.th_casing_temperatures[i]); }
#endif
        const double tc = p.tt_thermal_time_constant / 3600.0;
        for (int i = 0; i < 8; ++i) {
            double target = T.stage[i].outlet_temperature - 50.0;
            double ch = (target - T.th_rotor_temperatures[i]) / tc * dt;
            double mr = 5.0 * dt;
            ch = np_clip(ch, -mr, mr);
            T.th_rotor_temperatures[i] += ch;
            T.th_temperature_rates[i] = ch / dt * 60.0;
        }
        for (int i = 0; i < 6; ++i) {
            double target = T.stage[i].outlet_temperature - 80.0;
            double ch = (target - T.th_casing_temperatures[i]) / tc * dt;
            ch = np_clip(ch, -3.0 * dt, 3.0 * dt);
            T.th_casing_temperatures[i] += ch;
        }
        for (int i = 0; i < 14; ++i) {
            double target = T.stage[i].outlet_temperature - 20.0;
            double ch = (target - T.th_blade_temperatures[i]) / (tc * 0.5) * dt;
            ch = np_clip(ch, -10.0 * dt, 10.0 * dt);
            T.th_blade_temperatures[i] += ch;
        }
        double max_grad_r = 0.0, max_grad_c = 0.0;
        for (int i = 0; i < 7; ++i) {
            T.th_rotor_gradients[i] = fabs(T.th_rotor_temperatures[i + 1] - T.th_rotor_temperatures[i]) / (1.0 * 100);
            max_grad_r = (i == 0) ? T.th_rotor_gradients[0] : py_max(max_grad_r, T.th_rotor_gradients[i]);
        }
        for (int i = 0; i < 5; ++i) {
            T.th_casing_gradients[i] = fabs(T.th_casing_temperatures[i + 1] - T.th_casing_temperatures[i]) / (1.5 * 100);
            max_grad_c = (i == 0) ? T.th_casing_gradients[0] : py_max(max_grad_c, T.th_casing_gradients[i]);
        }
        double max_stress = 0.0, max_rate = 0.0;
        for (int i = 0; i < 8; ++i) {
            double strain = p.tt_thermal_expansion_coeff * (T.th_rotor_temperatures[i] - 25.0);
            T.th_stress_levels[i] = strain * p.tt_elastic_modulus * 0.1;
            max_stress = (i == 0) ? T.th_stress_levels[0] : py_max(max_stress, T.th_stress_levels[i]);
            max_rate = (i == 0) ? fabs(T.th_temperature_rates[0]) : py_max(max_rate, fabs(T.th_temperature_rates[i]));
        }
        T.th_max_thermal_stress = max_stress;
        double max_gradient = py_max(max_grad_r, max_grad_c);
        double rate_risk = py_min(1.0, max_rate / 10.0);
        double grad_risk = py_min(1.0, max_gradient / p.tt_max_thermal_gradient);
        double stress_risk = py_min(1.0, T.th_max_thermal_stress / p.tt_max_thermal_stress);
        T.th_thermal_shock_risk = py_max3(rate_risk, grad_risk, stress_risk);
    }

    // TurbineProtectionSystem.check_trip_conditions: enhanced_physics.py:348-437
    bool any_trip = false;
    {
        const double dt_seconds = dt * 3600.0;
        int reasons = (int)T.prot_trip_reasons;
        if (T.rotor_speed > p.tp_overspeed_trip) {
            T.prot_timer_overspeed += dt_seconds;
            if (T.prot_timer_overspeed >= p.tp_overspeed_delay) { any_trip = true; reasons |= 1; }
        } else T.prot_timer_overspeed = 0.0;
        if (T.vib_displacement_x > p.tp_vibration_trip) {
            T.prot_timer_vibration += dt_seconds;
            if (T.prot_timer_vibration >= p.tp_vibration_delay) { any_trip = true; reasons |= 2; }
        } else T.prot_timer_vibration = 0.0;
        double mbt = T.bearing[0].metal_temperature;
        for (int b = 1; b < 4; ++b) mbt = py_max(mbt, T.bearing[b].metal_temperature);
        if (mbt > p.tp_bearing_temp_trip) {
            T.prot_timer_bearing_temp += dt_seconds;
            if (T.prot_timer_bearing_temp >= p.tp_bearing_temp_delay) { any_trip = true; reasons |= 4; }
        } else T.prot_timer_bearing_temp = 0.0;
        if (T.thermal_expansion > p.tp_thrust_bearing_trip) { any_trip = true; reasons |= 8; }
        if (condenser_pressure > p.tp_low_vacuum_trip) { any_trip = true; reasons |= 16; }
        if (T.th_max_thermal_stress > p.tp_max_thermal_stress) { any_trip = true; reasons |= 32; }
        T.prot_trip_reasons = (double)reasons;
        T.prot_trip_active = as_flag(any_trip);
    }
    double power_reduction = any_trip ? (is_true(T.prot_trip_active) ? 0.0 : 1.0) : 1.0;
    double sg_avail = is_true(S.system_availability) ? 1.0 : 0.5;
    double total_reduction = power_reduction * sg_avail;
    T.total_power_output = stage_power_mw * total_reduction;
    T.overall_efficiency = T.ss_overall_efficiency;
    if (T.total_power_output > 0) {
        T.steam_rate = steam_flow / (T.total_power_output * 1000) * 3600;
        double h = turbine_steam_enthalpy(steam_temperature_in, steam_pressure);
        T.heat_rate = h * T.steam_rate / 1000;
    } else {
        T.steam_rate = 0.0; T.heat_rate = 0.0;
    }
    double rotor_eff = 1.0 - T.friction_torque / py_max(1.0, applied_torque) * 0.1;
    double thermal_eff = 1.0 - T.th_thermal_shock_risk * 0.1;
    T.performance_factor = T.ss_system_efficiency * rotor_eff * thermal_eff;
    T.availability_factor = is_true(T.prot_trip_active) ? 0.0 : 1.0;
    T.operating_hours += dt;

    out.mechanical_power = T.total_power_output / 0.985;
    out.electrical_power_gross = T.total_power_output;
    out.electrical_power_net = T.total_power_output * 0.98;
    out.overall_efficiency = T.overall_efficiency;
    out.steam_rate = T.steam_rate;
    out.hp_power = hp_power;
    out.lp_power = lp_power;
    out.condenser_pressure = condenser_pressure;
    out.condenser_temperature = turb_sat_temp(condenser_pressure);
    out.effective_steam_flow = steam_flow - T.ss_total_extraction_flow;
    out.lp6_outlet_enthalpy = T.stage[13].outlet_enthalpy;
}

}  // namespace nps
