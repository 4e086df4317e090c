// Steam generators: LMTD heat transfer with TSP and tube-scale fouling, first-order secondary
// pressure / level / quality dynamics, flow restrictions, and the 3-SG coordinator.
// Restates EnhancedSteamGeneratorPhysics.update_system
// (reference: nuclear_simulator/systems/secondary/steam_generator/enhanced_physics.py:433-547)
// and SteamGenerator.update_state (steam_generator/steam_generator.py:664-848) with callees.
#pragma once
#include "hd.h"
#include "state.h"
#include "prefetch.h"

namespace nps {

#define NPS_PI 3.141592653589793

// Thermodynamic correlations: steam_generator.py:854-941
NPS_HD_SHARED double sg_sat_temp(double p_mpa) {
    if (p_mpa <= 0.001) return 10.0;
    double p_bar = p_mpa * 10.0;
    double t;
    if (p_bar > 0) {
        double ln_p = nps_log(p_bar);
        t = 42.6776 + 34.5194 * ln_p + 2.8896 * py_pow(ln_p, 2.0) + 0.1153 * py_pow(ln_p, 3.0);
    } else {
        t = 10.0;
    }
    return np_clip(t, 10.0, 374.0);
}
NPS_HD double sg_h_f(double p_mpa) { return 4.18 * sg_sat_temp(p_mpa); }
NPS_HD double sg_h_g(double p_mpa) {
    double temp = sg_sat_temp(p_mpa);
    double h_f = sg_h_f(p_mpa);
    double h_fg = 2257.0 * py_pow(1.0 - temp / 374.0, 0.38);
    return h_f + h_fg;
}
NPS_HD double sg_water_enthalpy(double t, double p_mpa) { return 4.18 * t + 0.001 * (p_mpa - 0.1) * t; }
NPS_HD double sg_water_density(double t, double p_mpa) {
    double rho_temp = 1000.0 * (1.0 - 0.0003 * t);
    double pe = 1.0 + 4.5e-10 * p_mpa * 1e6;
    return rho_temp * pe;
}
NPS_HD double sg_steam_density(double t, double p_mpa) { return (p_mpa * 1e6) / (461.5 * (t + 273.15)); }

NPS_HD double tsp_total_thickness(const SGState& g, int level) {
    return (g.tsp_thickness[level][0] + g.tsp_thickness[level][1] + g.tsp_thickness[level][2] + g.tsp_thickness[level][3]);
}
NPS_HD double tsp_average_thickness(const SGState& g) {
    double t = 0.0;
    for (int i = 0; i < 7; ++i) t += tsp_total_thickness(g, i);
    return t / 7;
}

// calculate_flow_restriction / calculate_heat_transfer_degradation / calculate_flow_maldistribution /
// determine_fouling_stage: tsp_fouling_model.py:302-411.  Shared by the per-step update and by perform_cleaning.
NPS_HD void tsp_recompute_restriction(SGState& g) {
    double restr[7];
    double total_restriction = 0.0;
    const double hole_mm = 0.023 * 1000.0;
    for (int level = 0; level < 7; ++level) {
        double tt = tsp_total_thickness(g, level);
        double eff_d = hole_mm - 2.0 * tt;
        eff_d = py_max(eff_d, hole_mm * 0.1);
        double orig_area = NPS_PI * py_pow(hole_mm / 2.0, 2.0);
        double eff_area = NPS_PI * py_pow(eff_d / 2.0, 2.0);
        double r = 1.0 - eff_area / orig_area;
        restr[level] = r;
        total_restriction += r;
    }
    g.tsp_fouling_fraction = total_restriction / 7;
    double avg_area_ratio = py_max(1.0 - g.tsp_fouling_fraction, 0.1);
    g.tsp_pressure_drop_ratio = py_pow(1.0 / avg_area_ratio, 2.0);
    {   // calculate_heat_transfer_degradation
        double mixing = py_pow(g.tsp_fouling_fraction, 1.5);
        double mald = g.tsp_fouling_fraction * 0.3;
        g.tsp_heat_transfer_degradation = py_min((mixing + mald) * 0.6, 0.9);
    }
    {   // calculate_flow_maldistribution: np.mean / np.std over 7 values (sequential sums)
        double s = 0.0;
        for (int i = 0; i < 7; ++i) s += restr[i];
        double mean = s / 7.0;
        double v = 0.0;
        for (int i = 0; i < 7; ++i) { double x = restr[i] - mean; v += x * x; }
        double sd = sqrt(v / 7.0);
        g.tsp_flow_maldistribution = py_min(sd / (mean + 0.01), 1.0);
    }
    double ff = g.tsp_fouling_fraction;
    g.tsp_fouling_stage = (ff < 0.4) ? 0.0 : ((ff < 0.7) ? 1.0 : ((ff < 0.85) ? 2.0 : 3.0));
}

// TSPFoulingModel.update_fouling_state: tsp_fouling_model.py:654-724 (+ :195-445); the chemistry
// comes from the SG system's own WaterChemistry (WAT-001), which the step path never updates.
NPS_HD void tsp_update(SGState& g, const PlantParams& p, double temperature, double flow_velocity, double dt_hours) {
    double dt_seconds = dt_hours * 3600.0;
    double dty = dt_seconds / (365.25 * 24.0 * 3600.0);
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int lv = 0; lv < 7; ++lv) { NPS_TOUCH(g.tsp_thickness[lv][0]); NPS_TOUCH(g.tsp_thickness[lv][1]); NPS_TOUCH(g.tsp_thickness[lv][2]); NPS_TOUCH(g.tsp_thickness[lv][3]); }
#endif
    NPS_TOUCH(g.tsp_last_cleaning_time); NPS_TOUCH(g.tsp_cumulative_power_loss);
    g.tsp_operating_years += dty;
    g.tsp_last_cleaning_time += dty;
    double dt_years = dt_hours / (365.25 * 24.0);

    double temp_kelvin = temperature + 273.15;
    double temp_factor = nps_exp(-45000.0 / (8.314 * temp_kelvin));
    temp_factor = temp_factor / exp(-45000.0 / (8.314 * 573.15));   // constant argument: inlined so that it folds
    double ph_factor = 1.0 + 0.5 * fabs(p.sgwc_ph - 9.2);
    double velocity_factor = py_pow(flow_velocity / 3.0, 0.5);
    velocity_factor = np_clip(velocity_factor, 0.5, 2.0);
    double magnetite_rate = (2.5 * (1.0 + p.sgwc_iron * 1.5) * temp_factor * ph_factor * velocity_factor);
    double copper_rate = (0.8 * (1.0 + p.sgwc_copper * 2.0) * temp_factor * velocity_factor);
    double silica_rate = (1.2 * (1.0 + p.sgwc_silica / 100.0 * 1.8) * temp_factor * ph_factor);
    double bio_temp_factor = (temperature < 60) ? 1.0 : nps_exp(-(temperature - 60) / 20);
    double bio_rate = (0.5 * (1.0 + p.sgwc_dissolved_oxygen * 10.0) * bio_temp_factor * velocity_factor);

    const double max_thickness = 0.023 / 2.0 * 1000.0 * 0.9;
    for (int level = 0; level < 7; ++level) {
        double lf = 1.0 + 0.3 * (7 - level - 1) / (7 - 1);
        double mi = ((magnetite_rate * lf) / 1000.0) / 5.2 * 10.0;
        double ci = ((copper_rate * lf) / 1000.0) / 8.9 * 10.0;
        double si = ((silica_rate * lf) / 1000.0) / 2.2 * 10.0;
        double bi = ((bio_rate * lf) / 1000.0) / 1.2 * 10.0;
        g.tsp_thickness[level][0] += mi * dt_years;
        g.tsp_thickness[level][1] += ci * dt_years;
        g.tsp_thickness[level][2] += si * dt_years;
        g.tsp_thickness[level][3] += bi * dt_years;
        g.tsp_thickness[level][0] = py_min(g.tsp_thickness[level][0], max_thickness * 0.4);
        g.tsp_thickness[level][1] = py_min(g.tsp_thickness[level][1], max_thickness * 0.2);
        g.tsp_thickness[level][2] = py_min(g.tsp_thickness[level][2], max_thickness * 0.3);
        g.tsp_thickness[level][3] = py_min(g.tsp_thickness[level][3], max_thickness * 0.1);
    }
    tsp_recompute_restriction(g);
    double ff = g.tsp_fouling_fraction;
    int reasons = 0;
    if (ff >= 0.85) reasons |= 1;
    if (g.tsp_heat_transfer_degradation >= (1.0 - 0.60)) reasons |= 2;
    if (g.tsp_pressure_drop_ratio >= 5.0) reasons |= 4;
    if (g.tsp_flow_maldistribution >= 0.30) reasons |= 8;
    if (g.tsp_operating_years > 40.0 && ff > 0.5) reasons |= 16;
    g.tsp_shutdown_reasons = (double)reasons;
    g.tsp_shutdown_required = as_flag(reasons != 0);
    g.tsp_replacement_recommended = as_flag(ff >= 0.80 || g.tsp_operating_years > 40.0);
    double power_loss_mw = 100.0 * g.tsp_heat_transfer_degradation;
    g.tsp_cumulative_power_loss += power_loss_mw * dt_hours / 1000.0;
}

// TubeInteriorFouling.calculate_thermal_resistance: tube_interior_fouling.py:190-243
NPS_HD double tif_thermal_resistance(const SGState& g) {
    if (g.tif_scale_thickness <= 0) return 0.0;
    double thickness_m = g.tif_scale_thickness / 1000.0;
    double total = py_max(g.tif_scale_thickness, 0.001);
    double k = ((g.tif_comp[0] / total) * 0.5 + (g.tif_comp[1] / total) * 0.15 + (g.tif_comp[2] / total) * 0.3);
    k = py_max(k, 0.05);
    return thickness_m / k + 1e-5 + thickness_m * 0.001;
}

// TubeInteriorFouling.update_fouling_state: tube_interior_fouling.py:273-325 (+ :117-188, 245-271)
NPS_HD void tif_update(SGState& g, double temperature, double flow_velocity, double dt_seconds) {
    double dty = dt_seconds / (365.25 * 24.0 * 3600.0);
    // inputs as one group of independent loads (see lub_update_oil_quality)
    const double years0 = g.tif_operating_years, since0 = g.tif_last_cleaning_time, thick0 = g.tif_scale_thickness;
    const double c0 = g.tif_comp[0], c1 = g.tif_comp[1], c2 = g.tif_comp[2], loss0 = g.tif_cumulative_performance_loss;
    g.tif_operating_years = years0 + dty;
    g.tif_last_cleaning_time = since0 + dty;
    g.tif_scale_thickness = thick0; g.tif_comp[0] = c0; g.tif_comp[1] = c1; g.tif_comp[2] = c2;
    g.tif_cumulative_performance_loss = loss0;
    // chemistry dict passed by SteamGenerator.update_state: B 1000, Li 2.0, pH 7.2, O2 0.005
    double tk = temperature + 273.15, rk = 320.0 + 273.15;
    double temp_factor = nps_exp(-65000.0 / (8.314 * tk)) / exp(-65000.0 / (8.314 * rk));   // constant argument: folds
    double boric = 1.0 / (1.0 + 1000.0 / 1000.0 * 0.5);
    double lithium = py_max(0.5, 1.0 + (2.0 - 2.0) * 0.1);
    double ph_factor = 1.0 + 0.5 * fabs(7.2 - 7.2);
    double velocity_factor = np_clip(py_pow(flow_velocity / 5.0, -0.6), 0.5, 2.0);
    double oxygen = 1.0 + 0.005 * 10.0;
    double saturation = nps_exp(-g.tif_scale_thickness / 2.0);
    double rate = (0.001 * temp_factor * boric * lithium * ph_factor * velocity_factor * oxygen * saturation);
    g.tif_scale_formation_rate = np_clip(rate, 0.0, 0.1);
    double inc = g.tif_scale_formation_rate * dty;
    g.tif_scale_thickness += inc;
    g.tif_comp[0] += inc * 0.6;
    g.tif_comp[1] += inc * 0.3;
    g.tif_comp[2] += inc * 0.1;
    g.tif_scale_thermal_resistance = tif_thermal_resistance(g);
    g.tif_fouling_fraction = py_min(g.tif_scale_thermal_resistance / 0.001, 1.0);
    double loss = g.tif_scale_thermal_resistance * 1000.0;
    double dt_hours = dt_seconds / 3600.0;
    g.tif_cumulative_performance_loss += loss * dt_hours / 8760.0;
    g.tif_replacement_recommended = as_flag(g.tif_scale_thickness >= 2.0 || g.tif_operating_years > 40.0);
}

// SteamGenerator.update_state: steam_generator.py:664-848 (dt in seconds)
NPS_HD void sg_update(SGState& g, const PlantParams& p, double t_in, double t_out, double primary_flow,
                      double steam_flow_out, double feedwater_flow_in, double feedwater_temp, double dt) {
    NPS_TOUCH(g.water_level); NPS_TOUCH(g.steam_quality); NPS_TOUCH(g.steam_flow_rate); NPS_TOUCH(g.feedwater_flow_rate); NPS_TOUCH(g.tsp_heat_transfer_degradation); NPS_TOUCH(g.tif_scale_thermal_resistance); NPS_TOUCH(g.tsp_fouling_fraction); NPS_TOUCH(g.tsp_operating_years);
    // --- calculate_heat_transfer: steam_generator.py:150-314 ---
    double sat_temp = sg_sat_temp(g.secondary_pressure);
    double d1 = t_in - sat_temp, d2 = t_out - sat_temp;
    double lmtd = (fabs(d1 - d2) < 1.0) ? (d1 + d2) / 2.0 : (d1 - d2) / nps_log(d1 / d2);
    double flow_factor = py_pow(primary_flow / p.sg_primary_design_flow, 0.8);
    double h_primary = p.sg_primary_htc * flow_factor;
    double pressure_factor = py_pow(g.secondary_pressure / p.sg_design_pressure_secondary, 0.15);
    double h_secondary = p.sg_secondary_htc * pressure_factor;
    double r_primary = 1.0 / h_primary;
    double r_wall = p.sg_tube_wall_thickness / p.sg_tube_conductivity;
    double r_secondary = 1.0 / h_secondary;
    double overall_htc = 1.0 / (r_primary + r_wall + r_secondary);
    double htc_tsp = overall_htc * (1.0 - g.tsp_heat_transfer_degradation);
    double htc_all;
    if (g.tif_scale_thermal_resistance > 0) htc_all = 1.0 / (1.0 / htc_tsp + g.tif_scale_thermal_resistance);
    else htc_all = htc_tsp;
    double level_factor;
    if (g.water_level >= 12.5) level_factor = 1.0;
    else if (g.water_level <= 8.0) level_factor = 0.1;
    else level_factor = 0.1 + 0.9 * (g.water_level - 8.0) / (12.5 - 8.0);
    double effective_area = p.sg_heat_transfer_area * level_factor;
    double q = htc_all * effective_area * lmtd;
    double q_max = primary_flow * 5200.0 * (t_in - t_out);
    double td = t_in - t_out;
    if (td < 1.0) q = 0.0;
    else if (td < 5.0) q = py_min(q, q_max * 0.1);
    else q = py_min(q, q_max);
    if (primary_flow < 100.0) q = 0.0;
    if (q < 0) q = 0.0;
    double op_flux = py_max(q / p.sg_heat_transfer_area, 5000.0);
    double r_scale_primary = g.tif_scale_thermal_resistance;
    double r_scale_secondary;
    {   // _calculate_tsp_scale_thermal_resistance: steam_generator.py:603-634
        double avg_t = tsp_average_thickness(g);
        r_scale_secondary = (avg_t > 0) ? ((avg_t / 1000.0) / 3.0) * g.tsp_fouling_fraction : 0.0;
    }
    double r_to_wall = (1.0 / h_secondary + r_scale_secondary + (r_wall / 2.0) + r_scale_primary);
    g.tube_wall_temp = sat_temp + (op_flux * r_to_wall);
    g.overall_htc = overall_htc;
    g.heat_flux = op_flux;
    const double heat_transfer = q;

    g.secondary_temperature = sg_sat_temp(g.secondary_pressure);
    double tube_cs = NPS_PI * py_pow(p.sg_tube_inner_diameter / 2.0, 2.0);
    double total_flow_area = p.sg_tube_count * tube_cs;
    double avg_velocity = primary_flow / (1000.0 * total_flow_area);
    tsp_update(g, p, g.secondary_temperature, avg_velocity, dt / 3600.0);
    tif_update(g, (t_in + t_out) / 2.0, avg_velocity, dt);

    // _apply_tsp_flow_restrictions: steam_generator.py:516-547
    double cap = 1.0 / sqrt(g.tsp_pressure_drop_ratio);
    double actual_steam = py_min(steam_flow_out, p.sg_design_steam_flow_per_sg * cap);
    double actual_fw = py_min(feedwater_flow_in, p.sg_design_feedwater_flow_per_sg * cap);

    // --- calculate_secondary_side_dynamics: steam_generator.py:316-514 ---
    double P = g.secondary_pressure;
    double sat = sg_sat_temp(P);
    double h_f = sg_h_f(P), h_g = sg_h_g(P);
    double h_fg = h_g - h_f;
    double h_fw = sg_water_enthalpy(feedwater_temp, P);
    double rho_f = sg_water_density(sat, P), rho_g = sg_steam_density(sat, P);
    double mass_change_rate = actual_fw - actual_steam;
    double heat_kj = heat_transfer / 1000.0;
    double e_steam = heat_kj - actual_fw * (h_f - h_fw);
    double gen_rate = py_max(0.0, e_steam / h_fg);
    if (actual_fw < 0.1) gen_rate = 0.0;
    double design_heat = p.sg_design_thermal_power_per_sg / 1000.0;
    double heat_factor = (design_heat > 0) ? heat_kj / design_heat : 0.0;
    double eq_p = p.sg_design_pressure_secondary * (0.7 + 0.3 * heat_factor);
    eq_p = np_clip(eq_p, 3.0, 8.5);
    double demand_factor = (p.sg_secondary_design_flow > 0) ? actual_steam / p.sg_secondary_design_flow : 0.0;
    eq_p += -demand_factor * 0.5;
    eq_p = np_clip(eq_p, 3.0, 8.5);
    double decay = nps_exp(-dt / 60.0);
    double base_new_p = eq_p + (P - eq_p) * decay;
    double corr = 0.0;
    if (actual_fw < 0.1 && actual_steam > 100.0) {
        double depletion = -actual_steam / p.sg_secondary_water_mass;
        corr += depletion * P * 2.0 * dt;
    }
    double supply_factor = (p.sg_secondary_design_flow > 0) ? gen_rate / p.sg_secondary_design_flow : 0.0;
    double imbalance = supply_factor - demand_factor;
    corr += imbalance * 0.005 * dt;
    corr = np_clip(corr, -0.2, 0.2);
    double new_p = np_clip(base_new_p + corr, 1.0, 8.0);
    double cs = NPS_PI * py_pow(4.0 / 2.0, 2.0);
    double dl_mass = mass_change_rate * dt / (rho_f * cs);
    double vol_exp = gen_rate * dt * (1.0 / rho_g - 1.0 / rho_f);
    double dl_swell = vol_exp / cs;
    double new_level = np_clip(g.water_level + (dl_mass + dl_swell), 8.0, 16.0);
    double qdeg = 0.0;
    if (new_level < 11.0) qdeg += ((11.0 - new_level) / 3.0) * 0.02;
    double ffac = actual_steam / p.sg_secondary_design_flow;
    if (ffac > 1.1) qdeg += py_min((ffac - 1.1) * 0.01, 0.03);
    double design_flux = p.sg_design_thermal_power_per_sg / p.sg_heat_transfer_area;
    double cur_flux = heat_transfer / p.sg_heat_transfer_area;
    double flux_ratio = cur_flux / design_flux;
    if (flux_ratio > 1.2) qdeg += py_min((flux_ratio - 1.2) * 0.005, 0.02);
    double target_q = np_clip(0.995 - qdeg, 0.90, 1.0);
    double q_rate = (target_q - g.steam_quality) / 30.0;
    double new_q = np_clip(g.steam_quality + q_rate * dt, 0.90, 1.0);
    double new_void;
    if (new_q > 0) new_void = (new_q * rho_f) / (new_q * rho_f + (1 - new_q) * rho_g);
    else new_void = 0.0;
    new_void = np_clip(new_void, 0.0, 0.8);

    g.primary_inlet_temp = t_in;
    g.primary_outlet_temp = t_out;
    g.secondary_pressure = new_p;
    g.water_level = new_level;
    g.steam_quality = new_q;
    g.steam_void_fraction = new_void;
    g.steam_flow_rate = actual_steam;
    g.feedwater_flow_rate = actual_fw;
    g.feedwater_temperature = feedwater_temp;
    g.heat_transfer_rate = heat_transfer;
    g.thermal_efficiency = heat_transfer / p.sg_design_thermal_power_per_sg;
}

// EnhancedSteamGeneratorPhysics.update_system: steam_generator/enhanced_physics.py:433-547
NPS_HD void sg_system_update(SGSystemState& S, const PlantParams& p, const double* inlet_temps, const double* outlet_temps,
                             const double* flow_rates, double load_demand_fraction, double system_load_demand,
                             double feedwater_temperature, const double* actual_feedwater_flows, double dt,
                             const TurbineState* prefetch_next = nullptr) {
    S.load_demand = system_load_demand;
    double total_steam = p.sg_design_total_steam_flow * load_demand_fraction;
    double demands[3];
    if (!is_true(p.sg_auto_load_balancing)) {
        for (int i = 0; i < 3; ++i) demands[i] = total_steam / 3;
    } else {
        double tpf = 0.0 + flow_rates[0]; tpf += flow_rates[1]; tpf += flow_rates[2];
        if (tpf > 0) { for (int i = 0; i < 3; ++i) demands[i] = total_steam * (flow_rates[i] / tpf); }
        else { for (int i = 0; i < 3; ++i) demands[i] = total_steam / 3; }
    }
    NPS_UNIT_LOOP
    for (int i = 0; i < 3; ++i) {
        NPS_PREFETCH_SELF(S.sg[i]);
        if (i < 2) NPS_PREFETCH_FAR(S.sg[i + 1]);
        else if (prefetch_next) {   // the turbine's lubrication pre-step and bearing set come next
            NPS_PREFETCH_FAR(*prefetch_next);
            for (int b = 0; b < 4; ++b) NPS_PREFETCH_FAR(prefetch_next->bearing[b]);
        }
        sg_update(S.sg[i], p, inlet_temps[i], outlet_temps[i], flow_rates[i], demands[i], actual_feedwater_flows[i],
                  feedwater_temperature, dt);
    }
    double q = 0.0, f = 0.0, pr = 0.0, te = 0.0, ql = 0.0;
    int effective = 0;
    for (int i = 0; i < 3; ++i) {
        q += S.sg[i].heat_transfer_rate; f += S.sg[i].steam_flow_rate;
        pr += S.sg[i].secondary_pressure; te += S.sg[i].secondary_temperature; ql += S.sg[i].steam_quality;
        if (S.sg[i].thermal_efficiency > 0.1) effective++;
    }
    S.total_thermal_power = q;
    S.total_steam_flow = f;
    S.average_steam_pressure = pr / 3;
    S.average_steam_temperature = te / 3;
    S.average_steam_quality = ql / 3;
    S.system_availability = as_flag(effective >= 2);
    S.operating_hours += dt / 3600.0;
}

}  // namespace nps
