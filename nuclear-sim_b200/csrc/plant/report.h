// Report-only quantities that are pure functions of the plant state at the moment a row is logged (ReportState in
// state.h).  Evaluated at the end of the last fused substep and again by the maintenance kernel, because the reference
// logs AFTER maintenance (simulator/core/sim.py:209-223) and computes these in get_state_dict() at that moment.
#pragma once
#include "hd.h"
#include "state.h"
#include "feedwater.h"

namespace nps {

NPS_HD void plant_report_state(PlantState& st, const PlantParams& p) {
    ReportState& R = st.rep;
    double tsp_fouling_sum = 0.0, tsp_deg_sum = 0.0, scale_sum = 0.0, scale_res_sum = 0.0;
    for (int i = 0; i < 3; ++i) {
        const SGState& g = st.sgs.sg[i];
        // _calculate_primary_flow_restriction(design primary flow): steam_generator/steam_generator.py:549-601
        const double clean_d = p.sg_tube_inner_diameter;
        double eff_d = clean_d - 2.0 * (g.tif_scale_thickness / 1000.0);
        eff_d = py_max(eff_d, clean_d * 0.5);
        const double clean_area = NPS_PI * py_pow(clean_d / 2.0, 2.0);
        const double eff_area = NPS_PI * py_pow(eff_d / 2.0, 2.0);
        const double area_ratio = eff_area / clean_area;
        const double d_ratio = eff_d / clean_d;
        const double dp_ratio = 1.0 / py_pow(d_ratio, 4.0);
        double cap_factor;
        if (dp_ratio <= 3.0) cap_factor = area_ratio;
        else cap_factor = area_ratio * py_pow(3.0 / dp_ratio, 0.5);
        const double requested = p.sg_primary_design_flow;
        const double actual_primary = py_min(requested, p.sg_primary_design_flow * cap_factor);
        R.sg_max_primary_flow_capacity[i] = actual_primary;
        R.sg_primary_flow_restriction_factor[i] = (requested > 0) ? actual_primary / requested : 1.0;
        // _apply_tsp_flow_restrictions(design steam flow, design feedwater flow): steam_generator.py:516-547
        const double cap = 1.0 / sqrt(g.tsp_pressure_drop_ratio);
        const double actual_steam = py_min(p.sg_design_steam_flow_per_sg, p.sg_design_steam_flow_per_sg * cap);
        const double actual_fw = py_min(p.sg_design_feedwater_flow_per_sg, p.sg_design_feedwater_flow_per_sg * cap);
        R.sg_max_steam_flow_capacity[i] = actual_steam;
        R.sg_max_feedwater_flow_capacity[i] = actual_fw;
        R.sg_secondary_flow_restriction_factor[i] =
            (p.sg_design_steam_flow_per_sg > 0) ? actual_steam / p.sg_design_steam_flow_per_sg : 1.0;
        // _calculate_pump_energy_consumption: steam_generator.py:635-662
        const double penalty = 5.0 * (g.tsp_pressure_drop_ratio - 1.0) * 0.5;
        R.sg_fouling_energy_penalty_mw[i] = penalty;
        R.sg_total_pump_power_mw[i] = 5.0 + penalty;
        tsp_fouling_sum += g.tsp_fouling_fraction;
        tsp_deg_sum += g.tsp_heat_transfer_degradation;
        scale_sum += g.tif_scale_thickness;
        scale_res_sum += g.tif_scale_thermal_resistance;
    }
    // enhanced_physics.py:689-723
    const double avg_tsp_fouling = tsp_fouling_sum / 3, avg_tsp_deg = tsp_deg_sum / 3;
    const double avg_scale = scale_sum / 3, avg_scale_res = scale_res_sum / 3;
    R.sgs_total_fouling_impact = avg_tsp_deg + avg_scale_res * 1000.0;
    R.sgs_fouling_maintenance_needed = as_flag(avg_tsp_fouling > 0.15 || avg_scale > 0.5);
    for (int k = 0; k < 4; ++k) {
        R.fwp_efficiency_factor[k] = fwp_efficiency_factor(st.fw.pump[k]);
        R.fwp_flow_factor[k] = fwp_flow_factor(st.fw.pump[k]);
    }
    R.fw_diag_maintenance_urgency = 1.0 - st.fw.diag_health_score;   // performance_monitoring.py:565
}

}  // namespace nps
