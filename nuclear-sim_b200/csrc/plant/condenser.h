// Condenser: own water chemistry, tube degradation, three-species fouling, vacuum system with two
// steam-jet ejectors under lead/lag control, and LMTD heat transfer with air blanketing.
// Restates EnhancedCondenserPhysics.update_state
// (reference: nuclear_simulator/systems/secondary/condenser/physics.py:730-910) and callees.
#pragma once
#include "hd.h"
#include "state.h"
#include "water_chemistry.h"
#include "sg.h"

namespace nps {

// condenser/physics.py:1824-1853
NPS_HD_SHARED double cond_sat_temp(double p_mpa) {
    if (p_mpa <= 0.001) return 10.0;
    double p_bar = np_clip(p_mpa * 10.0, 0.01, 100.0);
    double t = 1730.63 / (8.07131 - nps_log10(p_bar)) - 233.426;
    if (p_mpa >= 0.005 && p_mpa <= 0.01) t = np_clip(t, 35.0, 45.0);
    return np_clip(t, 10.0, 374.0);
}
NPS_HD double cond_h_f(double p_mpa) { return 4.18 * cond_sat_temp(p_mpa); }
NPS_HD double cond_h_g(double p_mpa) {
    double temp = cond_sat_temp(p_mpa);
    double h_f = cond_h_f(p_mpa);
    double h_fg = 2257.0 * py_pow(1.0 - temp / 374.0, 0.38);
    return h_f + h_fg;
}

// SteamJetEjector.update_state: condenser/vacuum_pump.py:470-536 (+ :96-275)
NPS_HD void ejector_update(EjectorState& e, const PlantParams& p, int i, double suction_pressure, double required_capacity,
                           double motive_p, double motive_t, double dt, double& capacity_out, double& steam_out) {
    NPS_TOUCH(e.is_operating); NPS_TOUCH(e.overall_performance_factor); NPS_TOUCH(e.nozzle_fouling_factor); NPS_TOUCH(e.diffuser_fouling_factor); NPS_TOUCH(e.nozzle_erosion_factor); NPS_TOUCH(e.operating_hours);
    e.suction_pressure = suction_pressure;
    e.motive_steam_pressure_actual = motive_p;
    e.motive_steam_temp_actual = motive_t;
    if (is_true(e.is_operating)) {   // ejector.motive_steam_available is never cleared
        double capacity, msf = 0.0, scr = 0.0, cr = 1.0, er = 0.0;
        if (suction_pressure < p.ej_min_suction_pressure[i] || suction_pressure > p.ej_max_suction_pressure[i] ||
            motive_p < p.ej_min_motive_pressure[i]) {
            capacity = 0.0;
        } else {
            double pcf = py_pow(motive_p / p.ej_motive_steam_pressure[i], 0.5);
            double tr = (motive_t + 273.15) / (p.ej_motive_steam_temperature[i] + 273.15);
            double tcf = py_pow(tr, 0.25);
            double spr = suction_pressure / p.ej_design_suction_pressure[i];
            double scf = 1.0 / (1.0 + 0.5 * (spr - 1.0));
            double avail = (p.ej_design_capacity[i] * pcf * tcf * scf * e.overall_performance_factor);
            capacity = py_min(avail, required_capacity);
            capacity = py_max(0.0, capacity);
            if (capacity > 0) {
                double cf = capacity / p.ej_design_capacity[i];
                double base = (p.ej_base_steam_consumption[i] * py_pow(cf, p.ej_steam_consumption_exponent[i]));
                double pe = py_pow(spr, p.ej_pressure_effect_coefficient[i]);
                double df = 1.0 / py_max(0.5, e.overall_performance_factor);
                scr = base * pe * df;
                msf = capacity * scr;
            }
            cr = 0.101 / suction_pressure;
            er = capacity / py_max(0.001, msf);
        }
        e.current_capacity = capacity;
        e.motive_steam_flow = msf;
        e.steam_consumption_rate = scr;
        e.compression_ratio_actual = cr;
        e.entrainment_ratio = er;
        if (is_true(p.ej_two_stage[i])) {   // calculate_multistage_performance
            e.first_stage_capacity = capacity;
            double slf = 1.0 - 0.8 * 0.90;
            e.second_stage_capacity = capacity * slf;
            double stc = capacity * 2.0;
            e.intercondenser_load = (stc * 0.90 * 2200.0);
        }
        // update_degradation (is_operating)
        e.nozzle_fouling_factor = py_max(0.5, e.nozzle_fouling_factor - p.ej_nozzle_fouling_rate[i] * dt);
        e.diffuser_fouling_factor = py_max(0.6, e.diffuser_fouling_factor - p.ej_diffuser_fouling_rate[i] * dt);
        e.nozzle_erosion_factor = py_max(0.7, e.nozzle_erosion_factor - p.ej_erosion_rate[i] * dt);
        e.operating_hours += dt;
        e.overall_performance_factor = (e.nozzle_fouling_factor * e.diffuser_fouling_factor * e.nozzle_erosion_factor);
    } else {
        e.current_capacity = 0.0;
        e.motive_steam_flow = 0.0;
        e.steam_consumption_rate = 0.0;
    }
    capacity_out = e.current_capacity;
    steam_out = e.motive_steam_flow;
}

NPS_HD void ejector_command(EjectorState& e, const PlantParams& p, int i, int cmd, double motive_p) {
    if (cmd == 1) {          // start_ejector: vacuum_pump.py:277-295
        if (motive_p < p.ej_min_motive_pressure[i]) return;
        e.is_operating = 1.0;
        e.motive_steam_pressure_actual = motive_p;
    } else if (cmd == 0) {   // stop_ejector: :297-307
        e.is_operating = 0.0;
        e.current_capacity = 0.0;
        e.motive_steam_flow = 0.0;
    }
}

// VacuumSystem.update_state: condenser/vacuum_system.py:425-546 (lead_lag control strategy)
NPS_HD void vacuum_update(CondenserState& C, const PlantParams& p, double target_pressure, double motive_p_in,
                          double motive_t, double dt) {
    C.vs_motive_steam_pressure = motive_p_in - p.cd_vs_steam_pressure_drop;
    C.vs_motive_steam_temperature = motive_t;
    C.vs_motive_steam_available = as_flag(C.vs_motive_steam_pressure > p.cd_vs_low_motive_pressure_alarm);
    // update_air_leakage
    C.vs_current_air_leakage += p.cd_vs_leakage_degradation_rate * dt;
    C.vs_current_air_leakage = py_min(C.vs_current_air_leakage, p.cd_vs_base_air_leakage * 3.0);
    // calculate_required_capacity
    double perr = C.vs_condenser_pressure - target_pressure;
    double required = C.vs_current_air_leakage + 50.0 * perr;
    double max_cap = 0.0 + p.ej_design_capacity[0]; max_cap += p.ej_design_capacity[1];
    required = np_clip(required, 0.0, max_cap * 1.2);
    // VacuumControlLogic.update_control_logic: vacuum_system.py:61-118 ; cmd: -1 none, 0 stop, 1 start
    int cmd[2] = {-1, -1};
    {
        C.vc_rotation_timer += dt;
        if ((int)C.vc_lead_ejector < 0) { C.vc_lead_ejector = 0.0; C.vc_lag_ejector = 1.0; }
        int lead = (int)C.vc_lead_ejector, lag = (int)C.vc_lag_ejector;
        if (!is_true(C.ejector[lead].is_operating)) cmd[lead] = 1;
        if (lag >= 0) {
            bool lag_on = is_true(C.ejector[lag].is_operating);
            if (C.vs_condenser_pressure > p.cd_vs_auto_start_pressure && !lag_on) cmd[lag] = 1;
            else if (C.vs_condenser_pressure < p.cd_vs_auto_stop_pressure && lag_on) cmd[lag] = 0;
        }
        if (C.vc_rotation_timer >= p.cd_vs_rotation_interval) {   // _rotate_ejectors with two available ejectors
            int lag_index = lag;
            int new_lead = (lead + 1) % 2;
            int new_lag = (lag_index >= 0) ? (lag_index + 1) % 2 : -1;
            if (new_lag == new_lead) new_lag = (new_lag + 1) % 2;
            int old_lead = lead;
            C.vc_lead_ejector = (double)new_lead;
            C.vc_lag_ejector = (double)new_lag;
            if (old_lead != new_lead) { cmd[old_lead] = 0; cmd[new_lead] = 1; }
            C.vc_rotation_timer = 0.0;
        }
    }
    for (int i = 0; i < 2; ++i) if (cmd[i] >= 0) ejector_command(C.ejector[i], p, i, cmd[i], C.vs_motive_steam_pressure);
    double total_capacity = 0.0, total_steam = 0.0;
    NPS_UNIT_LOOP
    for (int i = 0; i < 2; ++i) {
        int n_running = (is_true(C.ejector[0].is_operating) ? 1 : 0) + (is_true(C.ejector[1].is_operating) ? 1 : 0);
        double req = is_true(C.ejector[i].is_operating) ? required / (double)(n_running > 1 ? n_running : 1) : 0.0;
        double cap, st;
        ejector_update(C.ejector[i], p, i, C.vs_condenser_pressure, req, C.vs_motive_steam_pressure,
                       C.vs_motive_steam_temperature, dt, cap, st);
        total_capacity += cap;
        total_steam += st;
    }
    {   // calculate_air_mass_balance
        double dt_s = dt * 3600.0;
        double rate = C.vs_current_air_leakage - total_capacity;
        double m = py_max(0.001, C.vs_air_mass_in_condenser + rate * dt_s);
        double ct = 39.0 + 273.15;
        double dens = m / p.cd_vs_condenser_volume;
        double pp = np_clip((dens * 287.0 * ct) / 1e6, 0.0001, 0.005);
        C.vs_air_mass_in_condenser = m;
        C.vs_air_partial_pressure = pp;
    }
    C.vs_steam_partial_pressure = py_max(0.005, target_pressure - C.vs_air_partial_pressure);
    C.vs_condenser_pressure = C.vs_steam_partial_pressure + C.vs_air_partial_pressure;
    C.vs_total_air_removal_rate = total_capacity;
    C.vs_total_steam_consumption = total_steam;
    C.vs_operating_hours += dt;
    {
        int n = 0; double s = 0.0;
        for (int i = 0; i < 2; ++i) if (is_true(C.ejector[i].is_operating)) { s += C.ejector[i].overall_performance_factor; n++; }
        C.vs_system_efficiency = (n > 0) ? s / n : 0.95;
        C.vs_alarm_high_pressure = as_flag(C.vs_condenser_pressure > p.cd_vs_high_pressure_alarm);
        C.vs_trip_high_pressure = as_flag(C.vs_condenser_pressure > p.cd_vs_high_pressure_trip);
        C.vs_alarm_low_motive_pressure = as_flag(C.vs_motive_steam_pressure < p.cd_vs_low_motive_pressure_alarm);
        C.vs_alarm_ejector_failure = as_flag(n == 0 && is_true(C.vs_motive_steam_available));
        C.vs_alarm_excessive_air_leakage = as_flag(C.vs_current_air_leakage / p.cd_vs_base_air_leakage > 2.0);
    }
}

struct CondenserResult {
    double heat_rejection_rate, condenser_pressure, cooling_water_temp_rise, cooling_water_outlet_temp;
    double thermal_performance_factor, vacuum_system_efficiency, condensate_temperature;
};

// EnhancedCondenserPhysics.update_state: condenser/physics.py:730-910 (dt in hours)
NPS_HD void condenser_update(CondenserState& C, const PlantParams& p, double steam_pressure, double steam_temperature,
                             double steam_flow, double steam_quality, double cw_flow, double cw_temp_in,
                             double motive_p, double motive_t, double dt, CondenserResult& out) {
    NPS_TOUCH(C.td_active_tube_count); NPS_TOUCH(C.td_vibration_damage_accumulation); NPS_TOUCH(C.td_average_wall_thickness); NPS_TOUCH(C.td_corrosion_damage_accumulation); NPS_TOUCH(C.td_plugged_tube_count); NPS_TOUCH(C.td_operating_hours); NPS_TOUCH(C.cooling_water_outlet_temp); NPS_TOUCH(C.fl_biofouling_thickness); NPS_TOUCH(C.fl_scale_thickness); NPS_TOUCH(C.fl_corrosion_product_thickness); NPS_TOUCH(C.fl_time_since_cleaning); NPS_TOUCH(C.fl_distribution_factor); NPS_TOUCH(C.vs_condenser_pressure); NPS_TOUCH(C.operating_hours); NPS_TOUCH(C.vs_current_air_leakage); NPS_TOUCH(C.vs_air_mass_in_condenser); NPS_TOUCH(C.vc_rotation_timer); NPS_TOUCH(C.vs_operating_hours);
    C.steam_inlet_pressure = steam_pressure;
    C.steam_inlet_temperature = steam_temperature;
    C.steam_inlet_flow = steam_flow;
    C.steam_inlet_quality = steam_quality;
    C.cooling_water_flow = cw_flow;
    C.cooling_water_inlet_temp = cw_temp_in;
    const MakeupWater mk = {7.2, 100.0, 300.0, 30.0, 8.0};
    wc_update(C.wc, true, mk, 0.02, dt);
    const double aggr = C.wc.water_aggressiveness;
    const double nutrient = py_min(2.0, C.wc.total_dissolved_solids / 500.0);

    double tube_area = NPS_PI * py_pow(p.cd_tube_inner_diameter / 2.0, 2.0);
    double total_flow_area = tube_area * C.td_active_tube_count;
    double velocity = (cw_flow / 1000.0) / total_flow_area;
    {   // TubeDegradationModel.update_tube_failures: physics.py:73-146
        if (velocity > p.cd_td_vibration_damage_threshold) {
            double r = py_pow(velocity - p.cd_td_vibration_damage_threshold, 2.0) * 0.001;
            C.td_vibration_damage_accumulation += r * dt;
        }
        double corr = (p.cd_td_corrosion_rate * aggr * (1.0 + C.td_vibration_damage_accumulation));
        double loss = corr * dt;
        C.td_average_wall_thickness = py_max(p.cd_td_wall_thickness_minimum, C.td_average_wall_thickness - loss);
        C.td_corrosion_damage_accumulation += loss;
        double vf = 1.0 + 10.0 * C.td_vibration_damage_accumulation;
        double cf = 1.0 + 5.0 * (C.td_corrosion_damage_accumulation / p.cd_td_wall_thickness_initial);
        double chf = 1.0 + aggr;
        double eff_rate = (p.cd_td_tube_failure_rate * vf * cf * chf);
        double failed = eff_rate * C.td_active_tube_count * dt;
        failed = py_min(failed, C.td_active_tube_count * 0.01);
        C.td_plugged_tube_count += failed;
        C.td_active_tube_count = py_max(1000.0, p.cd_td_initial_tube_count - C.td_plugged_tube_count);
        C.td_area_factor = C.td_active_tube_count / p.cd_td_initial_tube_count;
        double vif = p.cd_td_initial_tube_count / C.td_active_tube_count;
        C.td_pressure_drop_factor = py_pow(vif, 1.8);
        double leak_prone = py_min(failed * 0.1, C.td_active_tube_count * 0.001);
        C.td_tube_leak_rate = leak_prone * 0.001;
        C.td_operating_hours += dt;
    }
    {   // AdvancedFoulingModel.update_fouling: physics.py:324-388 (+ :167-322)
        double wt = (cw_temp_in + C.cooling_water_outlet_temp) / 2.0;
        double chlorine = C.wc.chlorine_residual, hardness = C.wc.hardness, ph = C.wc.ph;
        double antiscalant = C.wc.antiscalant_concentration, inhibitor = C.wc.corrosion_inhibitor_level;
        const double dissolved_oxygen = 8.0;
        // biofouling
        double tf = nps_exp(p.cd_fl_biofouling_temp_coefficient * (wt - 25.0));
        double clf = 1.0 / (1.0 + chlorine * 2.0);
        double nf = nutrient * p.cd_fl_biofouling_nutrient_factor;
        double gr = (p.cd_fl_biofouling_base_rate * tf * clf * nf);
        double thf = 1.0 / (1.0 + C.fl_biofouling_thickness / 2.0);
        double bio_inc = py_max(0.0, gr * thf * (dt / 1000.0));
        // scale
        double stf = nps_exp(p.cd_fl_scale_temp_coefficient * (wt - 25.0) / 10.0);
        double hf = (hardness / 150.0) * p.cd_fl_scale_hardness_coefficient;
        double phf = py_max(0.1, (ph - 6.0) / 2.0);
        double af = 1.0 / (1.0 + antiscalant / 5.0);
        double fr = (p.cd_fl_scale_base_rate * stf * hf * phf * af);
        double sthf = 1.0 / (1.0 + C.fl_scale_thickness / 1.0);
        double scale_inc = py_max(0.0, fr * sthf * (dt / 1000.0));
        // corrosion products
        double ctf = nps_exp((wt - 25.0) / 20.0);
        double of = dissolved_oxygen * p.cd_fl_corrosion_oxygen_coefficient;
        double cphf = 1.0 + fabs(ph - p.cd_fl_corrosion_ph_optimum) / 2.0;
        double inf = 1.0 / (1.0 + inhibitor / 10.0);
        double vlf = 1.0 / (1.0 + velocity / 2.0);
        double cfr = (p.cd_fl_corrosion_base_rate * ctf * of * cphf * inf * vlf);
        double corr_inc = py_max(0.0, cfr * (dt / 1000.0));
        C.fl_biofouling_thickness += bio_inc;
        C.fl_scale_thickness += scale_inc;
        C.fl_corrosion_product_thickness += corr_inc;
        C.fl_time_since_cleaning += dt;
        double tr = (C.fl_biofouling_thickness / 1000.0) / 0.5 + (C.fl_scale_thickness / 1000.0) / 2.0 +
                    (C.fl_corrosion_product_thickness / 1000.0) / 1.0;
        tr *= C.fl_distribution_factor;
        C.fl_total_fouling_resistance = tr;
        C.fl_distribution_factor = py_min(1.5, 1.0 + C.fl_time_since_cleaning / 8760.0);
    }
    vacuum_update(C, p, 0.007, motive_p, motive_t, dt);

    // calculate_enhanced_heat_transfer: physics.py:564-728
    double heat_transfer, cw_out, cw_rise;
    {
        double sat = cond_sat_temp(steam_pressure);
        double h_g = cond_h_g(steam_pressure), h_f = cond_h_f(steam_pressure);
        double h_fg = h_g - h_f;
        double h_cond = 4.18 * sat;
        double q = np_clip(steam_quality, 0.0, 1.0);
        double h_in = h_f + q * h_fg;
        double per_kg = h_in - h_cond;
        if (per_kg <= 0) per_kg = h_fg * q;
        double avail_w = (steam_flow * per_kg) * 1000;
        const double cp = 4180.0;
        double rise_est = avail_w / (cw_flow * cp);
        double t_out = cw_temp_in + rise_est;
        double d1 = py_max(sat - cw_temp_in, 0.1), d2 = py_max(sat - t_out, 0.1);
        double lmtd;
        if (fabs(d1 - d2) < 0.1) lmtd = (d1 + d2) / 2.0;
        else if (d1 > 0 && d2 > 0) lmtd = (d1 - d2) / nps_log(d1 / d2);
        else lmtd = (d1 + d2) / 2.0;
        double air_conc = (C.vs_air_partial_pressure / py_max(0.001, C.vs_condenser_pressure));
        double h_steam = p.cd_steam_side_htc * (1.0 - 0.5 * air_conc);
        double ff = py_pow(cw_flow / p.cd_design_cooling_water_flow, 0.8);
        double h_water = (p.cd_water_side_htc * ff) * py_pow(C.td_pressure_drop_factor, 0.2);
        double r_total = 1.0 / h_steam + C.fl_total_fouling_resistance +
                         p.cd_tube_wall_thickness / p.cd_tube_wall_conductivity + 1.0 / h_water;
        double htc = 1.0 / r_total;
        double eff_area = (p.cd_heat_transfer_area * C.td_area_factor);
        double theo = htc * eff_area * lmtd;
        heat_transfer = py_min(avail_w, theo);
        if (heat_transfer > avail_w) heat_transfer = avail_w;
        cw_rise = heat_transfer / (cw_flow * cp);
        cw_out = cw_temp_in + cw_rise;
        C.overall_htc = htc;
    }
    C.cooling_water_outlet_temp = cw_out;
    C.heat_rejection_rate = heat_transfer;
    C.condensate_temperature = cond_sat_temp(C.vs_condenser_pressure);
    C.condensate_flow = steam_flow;
    C.operating_hours += dt;
    double fouling_factor = py_max(0.3, 1.0 - C.fl_total_fouling_resistance * 5);
    C.thermal_performance_factor = C.td_area_factor * fouling_factor * C.vs_system_efficiency;

    out.heat_rejection_rate = C.heat_rejection_rate;
    out.condenser_pressure = C.vs_condenser_pressure;
    out.cooling_water_temp_rise = cw_rise;
    out.cooling_water_outlet_temp = C.cooling_water_outlet_temp;
    out.thermal_performance_factor = C.thermal_performance_factor;
    out.vacuum_system_efficiency = C.vs_system_efficiency;
    out.condensate_temperature = C.condensate_temperature;
}

}  // namespace nps
