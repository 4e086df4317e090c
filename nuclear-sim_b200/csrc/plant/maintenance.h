// Maintenance action effects: the state mutations of component.perform_maintenance(maintenance_type=...)
// as executed by AutoMaintenanceSystem._execute_work_order
// (reference: nuclear_simulator/systems/maintenance/auto_maintenance.py:504-673), applied to ONE plant.
//
// Which Python method runs is decided by the class of the registered component:
//   FWP-n  FeedwaterPump.perform_maintenance -> FeedwaterPumpLubricationSystem.perform_maintenance
//          (feedwater/pump_system.py:750-766 -> feedwater/pump_lubrication.py:625-1410)
//   SG-n   SteamGenerator.perform_maintenance                    (steam_generator/steam_generator.py:1092-1326)
//   HP-n / LP-n  TurbineStage.perform_maintenance                (turbine/stage_system.py:341-377)
//   SECONDARY-COMP-001-TURB  EnhancedTurbinePhysics.perform_maintenance  (turbine/enhanced_physics.py:1055-1267)
//   SECONDARY-COMP-001-COND  EnhancedCondenserPhysics.perform_maintenance (condenser/physics.py:1188-1372)
// The return value mirrors MaintenanceResult.success (decides whether StateManager.record_maintenance_result
// clears violations and resets cooldowns); report strings / cost / duration fields are host-side bookkeeping.
#pragma once
#include "hd.h"
#include "state.h"
#include "feedwater.h"
#include "sg.h"
#include "turbine.h"
#include "condenser.h"
#include "report.h"

namespace nps {

// Component targets, in the order the reference's threshold table lists them (state_manager.maintenance_thresholds).
enum MaintTarget : int {
    MT_PUMP0 = 0,          // FWP-1 .. FWP-4
    MT_FW_SYSTEM = 4,      // FEE-001 (FeedwaterPumpSystem)
    MT_SG0 = 5,            // SG-0 .. SG-2
    MT_SG_SYSTEM = 8,      // SECONDARY-COMP-001-SG
    MT_STAGE0 = 9,         // HP-1 .. HP-8, LP-1 .. LP-6
    MT_TURBINE = 23,       // SECONDARY-COMP-001-TURB
    MT_CONDENSER = 24,     // SECONDARY-COMP-001-COND
    MT_TURBINE_LUB = 25,   // TB-LUB-001 (TurbineBearingLubricationSystem)
    MT_EJECTOR0 = 26,      // SJE-001, SJE-002 (SteamJetEjector)
    MT_N_TARGETS = 28
};

// Action codes (names in nuclear_sim_b200/maintenance.py ACTION_CODES, same order).
enum MaintAction : int {
    MA_OIL_CHANGE = 0, MA_OIL_TOP_OFF, MA_BEARING_REPLACEMENT, MA_SEAL_REPLACEMENT, MA_COMPONENT_OVERHAUL,
    MA_SYSTEM_CLEANING, MA_BEARING_INSPECTION, MA_IMPELLER_INSPECTION, MA_IMPELLER_REPLACEMENT,
    MA_LUBRICATION_SYSTEM_CHECK, MA_MOTOR_INSPECTION, MA_OIL_ANALYSIS, MA_VIBRATION_ANALYSIS,
    MA_TSP_CHEMICAL_CLEANING, MA_TSP_MECHANICAL_CLEANING, MA_TUBE_BUNDLE_INSPECTION, MA_MOISTURE_SEPARATOR_MAINTENANCE,
    MA_SCALE_REMOVAL, MA_EDDY_CURRENT_TESTING, MA_SECONDARY_SIDE_CLEANING, MA_ROUTINE_MAINTENANCE,
    MA_TUBE_INTERIOR_SCALE_CLEANING, MA_PRIMARY_SCALE_CLEANING,
    MA_STAGE_CLEANING, MA_BLADE_REPLACEMENT, MA_STAGE_OVERHAUL,
    MA_CONDENSER_TUBE_CLEANING, MA_CONDENSER_TUBE_PLUGGING, MA_CONDENSER_CHEMICAL_CLEANING, MA_VACUUM_SYSTEM_TEST,
    MA_VACUUM_LEAK_DETECTION,
    MA_TURBINE_PERFORMANCE_TEST, MA_TURBINE_SYSTEM_OPTIMIZATION, MA_TURBINE_PROTECTION_TEST, MA_THERMAL_STRESS_ANALYSIS,
    MA_SYSTEM_COORDINATION_MAINTENANCE, MA_SYSTEM_STEAM_QUALITY_MAINTENANCE, MA_LOAD_BALANCING_MAINTENANCE,
    MA_WATER_CHEMISTRY_ADJUSTMENT, MA_TSP_INSPECTION, MA_TSP_FLOW_TEST, MA_TUBE_INTERIOR_INSPECTION,
    MA_TUBE_INTERIOR_EDDY_CURRENT_TESTING, MA_PRIMARY_CHEMISTRY_OPTIMIZATION, MA_CONDENSER_WATER_TREATMENT,
    MA_TURBINE_OIL_CHANGE, MA_TURBINE_OIL_TOP_OFF, MA_OIL_FILTER_REPLACEMENT, MA_OIL_COOLER_CLEANING,
    MA_LUBRICATION_SYSTEM_TEST,
    MA_VACUUM_EJECTOR_CLEANING, MA_VACUUM_EJECTOR_NOZZLE_REPLACEMENT, MA_VACUUM_EJECTOR_INSPECTION,
    MA_VACUUM_EJECTOR_MECHANICAL_CLEANING,
    MA_OTHER,              // any action name without a handler on the target: no state change
    MA_N_ACTIONS
};

// result codes written back per request
enum MaintStatus : int { MS_FAILED = 0, MS_SUCCESS = 1, MS_UNSUPPORTED_TARGET = 2 };

// FeedwaterPumpLubricationSystem._calculate_lubrication_effectiveness: feedwater/pump_lubrication.py:240-269
// (limits: contamination 15.0, acidity 1.6, moisture 0.08 - the same literals fwp_update_lubrication uses)
NPS_HD void fwp_weighted_lubrication_effectiveness(LubCore& L) {
    double contamination_factor = py_max(0.1, 1.0 - L.oil_contamination_level / 15.0);
    double acidity_factor = py_max(0.1, 1.0 - L.oil_acidity_number / 1.6);
    double moisture_factor = py_max(0.1, 1.0 - L.oil_moisture_content / 0.08);
    double antioxidant_factor = py_max(0.1, L.antioxidant_level / 100.0);
    double aw_factor = py_max(0.1, L.anti_wear_additive_level / 100.0);
    double ci_factor = py_max(0.1, L.corrosion_inhibitor_level / 100.0);
    L.lubrication_effectiveness = (contamination_factor * 0.25 + antioxidant_factor * 0.20 + aw_factor * 0.20 +
                                   ci_factor * 0.15 + acidity_factor * 0.10 + moisture_factor * 0.10);
    L.lubrication_effectiveness = py_max(0.1, py_min(1.0, L.lubrication_effectiveness));
}

// feedwater/pump_lubrication.py:677-1410.  bearing: 0 all, 1 motor_bearings, 2 pump_bearings, 3 thrust_bearing
NPS_HD int maintain_pump(FWPumpState& u, int action, int bearing) {
    LubCore& L = u.lub;
    double* w = L.component_wear;
    switch (action) {
        case MA_OIL_CHANGE:   // :677-709
            L.oil_level = 100.0; L.oil_temperature = 40.0; L.oil_contamination_level = 5.0;
            L.oil_acidity_number = 0.5; L.oil_moisture_content = 0.02;
            fwp_weighted_lubrication_effectiveness(L);
            fwp_performance_factors(u, 0.0);
            u.seal_leakage_rate = py_max(0.0, u.seal_leakage_rate * 0.5);
            return MS_SUCCESS;
        case MA_OIL_TOP_OFF: {   // :711-753 (target_level = 95.0)
            double oil_added = py_max(0.0, 95.0 - L.oil_level);
            if (oil_added > 0) {
                L.oil_level = py_min(100.0, 95.0);
                double dilution = oil_added / 100.0;
                L.oil_contamination_level *= (1.0 - dilution * 0.5);
                L.oil_acidity_number *= (1.0 - dilution * 0.3);
                L.oil_moisture_content *= (1.0 - dilution * 0.4);
                fwp_weighted_lubrication_effectiveness(L);
                fwp_performance_factors(u, 0.0);
            }
            return MS_SUCCESS;
        }
        case MA_BEARING_REPLACEMENT: {   // :755-810
            double removed;
            if (bearing == 0) {
                removed = w[FWL_MOTOR_BRG] + w[FWL_PUMP_BRG] + w[FWL_THRUST_BRG];
                w[FWL_MOTOR_BRG] = 0.0; w[FWL_PUMP_BRG] = 0.0; w[FWL_THRUST_BRG] = 0.0;
            } else if (bearing >= 1 && bearing <= 3) {
                int c = (bearing == 1) ? FWL_MOTOR_BRG : ((bearing == 2) ? FWL_PUMP_BRG : FWL_THRUST_BRG);
                removed = w[c];
                w[c] = 0.0;
            } else {
                return MS_FAILED;
            }
            fwp_weighted_lubrication_effectiveness(L);
            fwp_performance_factors(u, 0.0);
            u.vibration_increase = py_max(0.0, u.vibration_increase - removed * 0.1);
            return MS_SUCCESS;
        }
        case MA_SEAL_REPLACEMENT:   // :812-840
            w[FWL_SEALS] = 0.0;
            u.seal_leakage_rate = 0.0;
            fwp_weighted_lubrication_effectiveness(L);
            fwp_performance_factors(u, 0.0);
            return MS_SUCCESS;
        case MA_COMPONENT_OVERHAUL:   // :842-888
            for (int c = 0; c < FWL_NCOMP; ++c) w[c] = 0.0;
            L.oil_level = 100.0; L.oil_temperature = 40.0; L.oil_contamination_level = 5.0;
            L.oil_acidity_number = 0.5; L.oil_moisture_content = 0.02;
            u.seal_leakage_rate = 0.0;
            u.vibration_increase = 0.0;
            fwp_weighted_lubrication_effectiveness(L);
            fwp_performance_factors(u, 0.0);
            return MS_SUCCESS;
        case MA_SYSTEM_CLEANING: {   // :890-925
            double old_c = L.oil_contamination_level;
            double red = py_min(old_c * 0.7, 50.0);
            L.oil_contamination_level = py_max(5.0, old_c - red);
            L.oil_acidity_number *= 0.8;
            L.oil_moisture_content *= 0.9;
            for (int c = 0; c < FWL_NCOMP; ++c) w[c] = py_max(0.0, w[c] - 0.5);
            fwp_weighted_lubrication_effectiveness(L);
            fwp_performance_factors(u, 0.0);
            return MS_SUCCESS;
        }
        case MA_BEARING_INSPECTION: {   // :927-975
            double mx = py_max3(w[FWL_MOTOR_BRG], w[FWL_PUMP_BRG], w[FWL_THRUST_BRG]);
            if (mx > 5.0) {
                w[FWL_MOTOR_BRG] *= 0.9; w[FWL_PUMP_BRG] *= 0.9; w[FWL_THRUST_BRG] *= 0.9;
                fwp_performance_factors(u, 0.0);
            }
            return MS_SUCCESS;
        }
        case MA_IMPELLER_INSPECTION: {   // :977-1063
            double iw = w[FWL_IMPELLER];
            double mb = w[FWL_MOTOR_BRG], pb = w[FWL_PUMP_BRG], tb = w[FWL_THRUST_BRG];
            double mx = py_max3(mb, pb, tb);
            if (iw > 3.0 || mx > 5.0) {
                w[FWL_IMPELLER] = py_max(0.0, iw * 0.9);
                w[FWL_MOTOR_BRG] = py_max(0.0, mb - 0.5);
                w[FWL_PUMP_BRG] = py_max(0.0, pb - 0.5);
                w[FWL_THRUST_BRG] = py_max(0.0, tb - 0.5);
                fwp_performance_factors(u, 0.0);
            }
            return MS_SUCCESS;
        }
        case MA_IMPELLER_REPLACEMENT: {   // :1065-1118
            double old = w[FWL_IMPELLER];
            w[FWL_IMPELLER] = 0.0;
            fwp_performance_factors(u, 0.0);
            u.vibration_increase = py_max(0.0, u.vibration_increase - old * 0.08);
            return MS_SUCCESS;
        }
        case MA_LUBRICATION_SYSTEM_CHECK: {   // :1120-1270
            if (L.oil_level < 95.0) {
                double target = py_min(95.0, L.oil_level + 5.0);
                double added = target - L.oil_level;
                L.oil_level = target;
                if (added > 0) {
                    double dilution = added / 100.0;
                    double cd = dilution * 0.5;
                    L.oil_contamination_level = py_max(1.0, L.oil_contamination_level * (1.0 - cd));
                    double boost = dilution * 15.0;
                    L.antioxidant_level = py_min(100.0, L.antioxidant_level + boost);
                    L.anti_wear_additive_level = py_min(100.0, L.anti_wear_additive_level + boost * 0.8);
                }
            }
            double red = py_min(L.oil_contamination_level * 0.3, 5.0);
            L.oil_contamination_level = py_max(1.0, L.oil_contamination_level - red);
            const double restore = 15.0;
            L.antioxidant_level = py_min(100.0, L.antioxidant_level + restore);
            L.anti_wear_additive_level = py_min(100.0, L.anti_wear_additive_level + restore * 0.8);
            L.corrosion_inhibitor_level = py_min(100.0, L.corrosion_inhibitor_level + restore * 0.6);
            w[FWL_MOTOR_BRG] = py_max(0.0, w[FWL_MOTOR_BRG] - 0.5);
            w[FWL_PUMP_BRG] = py_max(0.0, w[FWL_PUMP_BRG] - 0.5);
            w[FWL_THRUST_BRG] = py_max(0.0, w[FWL_THRUST_BRG] - 0.5);
            u.seal_leakage_rate = py_max(0.0, u.seal_leakage_rate * 0.9);
            fwp_weighted_lubrication_effectiveness(L);
            return MS_SUCCESS;
        }
        case MA_MOTOR_INSPECTION:   // :1272-1310
            if (w[FWL_MOTOR_BRG] > 3.0) {
                w[FWL_MOTOR_BRG] *= 0.95;
                fwp_performance_factors(u, 0.0);
            }
            return MS_SUCCESS;
        case MA_OIL_ANALYSIS:         // :1312-1358 (assessment only)
        case MA_VIBRATION_ANALYSIS:   // :1360-1410 (assessment only)
            return MS_SUCCESS;
        default:                      // "Unknown maintenance type": :667-675
            return MS_FAILED;
    }
}

// TSPFoulingModel.perform_cleaning: steam_generator/tsp_fouling_model.py:447-487
// (chemical_cleaning_effectiveness 0.75, mechanical 0.85: tsp_fouling_model.py:109-110)
NPS_HD void tsp_perform_cleaning(SGState& g, double effectiveness) {
    for (int level = 0; level < 7; ++level) {
        g.tsp_thickness[level][0] *= (1.0 - effectiveness);
        g.tsp_thickness[level][1] *= (1.0 - effectiveness * 0.8);
        g.tsp_thickness[level][2] *= (1.0 - effectiveness * 0.9);
        g.tsp_thickness[level][3] *= (1.0 - effectiveness);
    }
    g.tsp_total_cleaning_cycles += 1.0;
    g.tsp_last_cleaning_time = 0.0;
    tsp_recompute_restriction(g);
}

// steam_generator/steam_generator.py:1092-1326
NPS_HD int maintain_sg(SGState& g, int action) {
    switch (action) {
        case MA_TSP_CHEMICAL_CLEANING: tsp_perform_cleaning(g, 0.75); return MS_SUCCESS;     // :1105-1122
        case MA_TSP_MECHANICAL_CLEANING: tsp_perform_cleaning(g, 0.85); return MS_SUCCESS;   // :1124-1140
        case MA_TUBE_BUNDLE_INSPECTION: return MS_SUCCESS;                                   // :1142-1166
        case MA_MOISTURE_SEPARATOR_MAINTENANCE: {                                            // :1168-1185
            double q = g.steam_quality;
            double improvement = 0.999 - q;
            g.steam_quality = py_min(0.999, q + improvement * 0.8);
            return MS_SUCCESS;
        }
        case MA_SCALE_REMOVAL:                   // :1187-1201 -> TubeInteriorFouling._primary_scale_cleaning
        case MA_TUBE_INTERIOR_SCALE_CLEANING:    // :1278-1281
        case MA_PRIMARY_SCALE_CLEANING: {        // :1293-1296; tube_interior_fouling.py:361-420 (chemical: 0.90)
            const double eff = 0.90;
            double removed = g.tif_scale_thickness * eff;
            g.tif_scale_thickness -= removed;
            g.tif_scale_thickness = py_max(0.0, g.tif_scale_thickness);
            for (int c = 0; c < 3; ++c) g.tif_comp[c] *= (1.0 - eff);
            g.tif_scale_thermal_resistance = tif_thermal_resistance(g);
            g.tif_fouling_fraction = py_min(g.tif_scale_thermal_resistance / 0.001, 1.0);
            g.tif_last_cleaning_time = 0.0;   // FoulingModelBase._update_maintenance_history: fouling_model_base.py:210-214
            return MS_SUCCESS;
        }
        // :1218-1241 reads tsp_state['operating_years'], a key TSPFoulingModel.get_state_dict does not provide: the
        // KeyError is caught by _perform_maintenance_action (auto_maintenance.py:656-665) -> success False, no change
        case MA_EDDY_CURRENT_TESTING: return MS_FAILED;
        case MA_SECONDARY_SIDE_CLEANING: g.tsp_fouling_fraction *= (1.0 - 0.3); return MS_SUCCESS;   // :1243-1261
        case MA_ROUTINE_MAINTENANCE: g.steam_quality = py_min(0.999, g.steam_quality + 0.001); return MS_SUCCESS;   // :1303-1315
        // :1203-1216 WaterChemistry.reset() on WAT-001, the SG system's own instance: it is never updated, so its members
        // stay at the design values reset() would write (they are PlantParams sgwc_* here) -> success, no change
        case MA_WATER_CHEMISTRY_ADJUSTMENT: return MS_SUCCESS;
        // :1263-1271 -> TSPFoulingModel.perform_maintenance (tsp_fouling_model.py:489-521): report only, then
        // FoulingModelBase._update_maintenance_history (fouling_model_base.py:210-214) counts it as a cleaning cycle
        case MA_TSP_INSPECTION:
        case MA_TSP_FLOW_TEST: g.tsp_total_cleaning_cycles += 1.0; g.tsp_last_cleaning_time = 0.0; return MS_SUCCESS;
        // :1273-1291 -> TubeInteriorFouling.perform_maintenance (tube_interior_fouling.py:327-359): inspection and eddy
        // current testing report only; primary_chemistry_optimization writes boric acid 1000 / lithium 2.0 / pH 7.2, the
        // values the model is built with (:82-84) and nothing else ever changes; all three end in _update_maintenance_history
        case MA_TUBE_INTERIOR_INSPECTION:
        case MA_TUBE_INTERIOR_EDDY_CURRENT_TESTING:
        case MA_PRIMARY_CHEMISTRY_OPTIMIZATION: g.tif_last_cleaning_time = 0.0; return MS_SUCCESS;
        default: return MS_FAILED;               // :1317-1324
    }
}

// TurbineStage.perform_maintenance: turbine/stage_system.py:341-377 (always returns a dict -> success)
NPS_HD int maintain_stage(TurbineStageState& s, const PlantParams& p, int k, int action) {
    if (action == MA_STAGE_CLEANING) {
        s.deposit_thickness = 0.0; s.fouling_factor = 1.0; s.efficiency_degradation *= 0.3;
    } else if (action == MA_BLADE_REPLACEMENT) {
        s.blade_wear_factor = 1.0; s.blade_condition_factor = s.fouling_factor;
    } else if (action == MA_STAGE_OVERHAUL) {
        s.deposit_thickness = 0.0; s.fouling_factor = 1.0; s.blade_wear_factor = 1.0; s.blade_condition_factor = 1.0;
        s.efficiency_degradation = 0.0; s.actual_efficiency = p.ts_design_efficiency[k];
    }
    return MS_SUCCESS;
}

// AdvancedFoulingModel.calculate_total_fouling_resistance: condenser/physics.py:297-322
NPS_HD double cond_total_fouling_resistance(const CondenserState& C) {
    double tr = (C.fl_biofouling_thickness / 1000.0) / 0.5 + (C.fl_scale_thickness / 1000.0) / 2.0 +
                (C.fl_corrosion_product_thickness / 1000.0) / 1.0;
    tr *= C.fl_distribution_factor;
    return tr;
}

// AdvancedFoulingModel.perform_cleaning("chemical"): condenser/physics.py:386-450
NPS_HD void cond_perform_chemical_cleaning(CondenserState& C) {
    double bio_removed = C.fl_biofouling_thickness * 0.8;
    double scale_removed = C.fl_scale_thickness * 0.6;
    double corrosion_removed = C.fl_corrosion_product_thickness * 0.3;
    C.fl_biofouling_thickness -= bio_removed;
    C.fl_scale_thickness -= scale_removed;
    C.fl_corrosion_product_thickness -= corrosion_removed;
    C.fl_time_since_cleaning = 0.0;
    C.fl_distribution_factor = 1.0;
    C.fl_total_fouling_resistance = cond_total_fouling_resistance(C);
}

// EnhancedCondenserPhysics.perform_maintenance: condenser/physics.py:1188-1372
NPS_HD int maintain_condenser(CondenserState& C, const PlantParams& p, int action) {
    switch (action) {
        case MA_CONDENSER_TUBE_CLEANING: {   // :1204-1232
            double old_f = C.fl_total_fouling_resistance;
            cond_perform_chemical_cleaning(C);
            double reduction = old_f - C.fl_total_fouling_resistance;
            double improvement = (reduction / py_max(0.001, old_f)) * 100;
            C.thermal_performance_factor = py_min(1.0, C.thermal_performance_factor + improvement * 0.01);
            return MS_SUCCESS;
        }
        case MA_CONDENSER_TUBE_PLUGGING: {   // :1234-1263 (tubes_to_plug = 10)
            const double tubes = 10.0;
            double old_active = C.td_active_tube_count;
            C.td_plugged_tube_count += tubes;
            C.td_active_tube_count = py_max(1000.0, old_active - tubes);
            // :1244-1246 divides by self.config.tube_count, an attribute CondenserConfig does not have (it is
            // heat_transfer.tube_count, condenser/config.py:43): AttributeError after the two counts were updated,
            // caught by _perform_maintenance_action -> success False, area factor and leak rate untouched.
            (void)p;
            return MS_FAILED;
        }
        case MA_CONDENSER_CHEMICAL_CLEANING:   // :1265-1289
            cond_perform_chemical_cleaning(C);
            C.thermal_performance_factor = py_min(1.0, C.thermal_performance_factor + 0.1);
            return MS_SUCCESS;
        case MA_VACUUM_SYSTEM_TEST: return MS_SUCCESS;   // :1319-1343 (assessment only)
        case MA_VACUUM_LEAK_DETECTION:                   // :1345-1362
            C.vs_current_air_leakage *= 0.5;
            return MS_SUCCESS;
        case MA_CONDENSER_WATER_TREATMENT: {             // :1291-1317
            // WaterChemistry.perform_chemical_treatment("standard") on WAT-002: water_chemistry.py:508-521
            WaterChemState& w = C.wc;
            w.ph += (9.2 - w.ph) * 0.3;
            w.chlorine_residual = 1.0; w.antiscalant_concentration = 5.0; w.corrosion_inhibitor_level = 10.0;
            w.treatment_efficiency = 0.95;
            w.last_treatment_time = 0.0;
            wc_composite(w);
            C.fl_biofouling_thickness *= 0.9; C.fl_scale_thickness *= 0.8; C.fl_corrosion_product_thickness *= 0.7;
            C.fl_total_fouling_resistance = cond_total_fouling_resistance(C);
            return MS_SUCCESS;
        }
        default: return MS_FAILED;                       // :1364-1371
    }
}

// EnhancedTurbinePhysics.perform_maintenance: turbine/enhanced_physics.py:1055-1267 (Python min/max argument order kept)
NPS_HD int maintain_turbine(TurbineState& T, int action) {
    switch (action) {
        case MA_TURBINE_PERFORMANCE_TEST:        // :1066-1101
            T.overall_efficiency = py_min(0.34, T.overall_efficiency + 0.02);
            T.performance_factor = py_min(1.0, T.performance_factor + 0.05);
            return MS_SUCCESS;
        case MA_TURBINE_SYSTEM_OPTIMIZATION:     // :1103-1135
            T.performance_factor = py_min(1.0, T.performance_factor + 0.08);
            T.ss_system_efficiency = py_min(1.0, T.ss_system_efficiency + 0.03);
            for (int b = 0; b < 4; ++b) T.bearing[b].efficiency_factor = py_min(1.0, T.bearing[b].efficiency_factor + 0.02);
            T.lub.lubrication_effectiveness = py_min(1.0, T.lub.lubrication_effectiveness + 0.05);
            return MS_SUCCESS;
        case MA_TURBINE_PROTECTION_TEST:         // :1137-1167; reset_protection_system :473-479
            if (is_true(T.prot_trip_active)) {
                T.prot_trip_active = 0.0; T.prot_trip_reasons = 0.0;
                T.prot_timer_overspeed = 0.0; T.prot_timer_vibration = 0.0; T.prot_timer_bearing_temp = 0.0;
            }
            T.availability_factor = py_min(1.0, T.availability_factor + 0.03);
            return MS_SUCCESS;
        case MA_THERMAL_STRESS_ANALYSIS: {       // :1169-1199
            const double original = T.th_max_thermal_stress;
            const double reduction = py_min(100e6, original * 0.1);
            T.th_max_thermal_stress -= reduction;
            T.th_thermal_shock_risk *= 0.8;
            return MS_SUCCESS;
        }
        case MA_VIBRATION_ANALYSIS: {            // :1201-1236; last_update_results['vibration_displacement'] = total displacement
            const double current = T.vib_displacement_x;
            const double reduction = py_min(5.0, current * 0.3);
            for (int b = 0; b < 4; ++b) T.bearing[b].vibration_displacement = py_max(0.0, T.bearing[b].vibration_displacement - reduction);
            T.thermal_bow *= 0.7;                // bearing.vibration_velocity (*= 0.8) is not read by the step path
            return MS_SUCCESS;
        }
        case MA_ROUTINE_MAINTENANCE:             // :1238-1257
            T.performance_factor = py_min(1.0, T.performance_factor + 0.01);
            T.overall_efficiency = py_min(0.34, T.overall_efficiency + 0.002);
            for (int b = 0; b < 4; ++b) {
                T.bearing[b].efficiency_factor = py_min(1.0, T.bearing[b].efficiency_factor + 0.005);
                T.bearing[b].metal_temperature = py_max(80.0, T.bearing[b].metal_temperature - 0.5);
            }
            return MS_SUCCESS;
        default: return MS_FAILED;               // :1259-1266 (unknown maintenance type)
    }
}

// EnhancedSteamGeneratorPhysics.perform_maintenance: steam_generator/enhanced_physics.py:1062-1190.  Its own
// performance_factor / load_balance_factor start at 1.0 and are only ever raised towards 1.0 on the step path
// (the chemistry update that lowers them, :905-917, is never called: SURVEY a16), so they are not carried.
NPS_HD int maintain_sg_system(SGSystemState& S, int action) {
    switch (action) {
        case MA_SYSTEM_COORDINATION_MAINTENANCE: S.system_availability = 1.0; return MS_SUCCESS;   // :1073-1088
        case MA_SYSTEM_STEAM_QUALITY_MAINTENANCE:                                                    // :1090-1116
            for (int i = 0; i < 3; ++i)
                if (S.sg[i].steam_quality < 0.99) maintain_sg(S.sg[i], MA_MOISTURE_SEPARATOR_MAINTENANCE);
            return MS_SUCCESS;
        case MA_LOAD_BALANCING_MAINTENANCE: {    // :1118-1154: TSP chemical cleaning on the first two degraded SGs
            int done = 0;
            for (int i = 0; i < 3 && done < 2; ++i)
                if (S.sg[i].tsp_heat_transfer_degradation > 0.05) { maintain_sg(S.sg[i], MA_TSP_CHEMICAL_CLEANING); ++done; }
            return MS_SUCCESS;
        }
        case MA_ROUTINE_MAINTENANCE:             // :1156-1174
            for (int i = 0; i < 3; ++i) maintain_sg(S.sg[i], MA_ROUTINE_MAINTENANCE);
            return MS_SUCCESS;
        default: return MS_FAILED;               // :1176-1190: delegation needs an sg_index keyword nobody passes
    }
}

// TurbineBearingLubricationSystem.perform_maintenance: turbine/turbine_bearing_lubrication.py:481-665
// (component_wear order: hp_journal_bearing, lp_journal_bearing, thrust_bearing, seal_oil_system, oil_coolers)
NPS_HD int maintain_turbine_lub(TurbineState& T, int action) {
    LubCore& L = T.lub;
    switch (action) {
        case MA_TURBINE_OIL_CHANGE:              // :492-523
            L.oil_contamination_level = 1.0; L.oil_acidity_number = 0.05; L.oil_moisture_content = 0.01; L.oil_level = 100.0;
            L.lubrication_effectiveness = py_min(1.0, L.lubrication_effectiveness + 0.15);
            L.oil_temperature = py_max(45.0, L.oil_temperature - 5.0);
            return MS_SUCCESS;
        case MA_TURBINE_OIL_TOP_OFF: {           // :525-547
            double oil_added = py_min(100.0 - L.oil_level, 50.0);
            L.oil_level += oil_added;
            double dilution = oil_added / 100.0;
            L.oil_contamination_level = py_max(1.0, L.oil_contamination_level - dilution * 2.0);
            L.oil_acidity_number = py_max(0.05, L.oil_acidity_number - dilution * 0.1);
            return MS_SUCCESS;
        }
        case MA_OIL_FILTER_REPLACEMENT: {        // :549-571
            double reduction = py_min(5.0, L.oil_contamination_level * 0.6);
            L.oil_contamination_level -= reduction;
            L.oil_contamination_level = py_max(1.0, L.oil_contamination_level);
            L.lubrication_effectiveness = py_min(1.0, L.lubrication_effectiveness + 0.05);
            return MS_SUCCESS;
        }
        case MA_OIL_COOLER_CLEANING: {           // :573-598
            double original = T.lub_oil_cooling_effectiveness;
            T.lub_oil_cooling_effectiveness = py_min(1.0, T.lub_oil_cooling_effectiveness + 0.15);
            double temp_reduction = (1.0 - original) * 15.0;
            L.oil_temperature = py_max(45.0, L.oil_temperature - temp_reduction);
            L.component_wear[4] = py_max(0.0, L.component_wear[4] - 5.0);
            return MS_SUCCESS;
        }
        case MA_LUBRICATION_SYSTEM_TEST:         // :600-637
            L.lubrication_effectiveness = py_min(1.0, L.lubrication_effectiveness + 0.1);
            return MS_SUCCESS;
        case MA_ROUTINE_MAINTENANCE:             // :639-657
            L.lubrication_effectiveness = py_min(1.0, L.lubrication_effectiveness + 0.02);
            L.oil_contamination_level = py_max(1.0, L.oil_contamination_level - 0.5);
            L.oil_temperature = py_max(45.0, L.oil_temperature - 1.0);
            for (int c = 0; c < 5; ++c) L.component_wear[c] = py_max(0.0, L.component_wear[c] - 0.5);
            return MS_SUCCESS;
        default: return MS_FAILED;               // :659-667
    }
}

// SteamJetEjector.perform_maintenance: condenser/vacuum_pump.py:338-466 (perform_cleaning :309-336); every name succeeds,
// names without a branch of their own take the "general" reset
NPS_HD int maintain_ejector(EjectorState& e, int action) {
    switch (action) {
        case MA_VACUUM_EJECTOR_CLEANING:         // chemical
            e.nozzle_fouling_factor = py_min(1.0, e.nozzle_fouling_factor + 0.3);
            e.diffuser_fouling_factor = py_min(1.0, e.diffuser_fouling_factor + 0.4);
            break;
        case MA_VACUUM_EJECTOR_MECHANICAL_CLEANING:
            e.nozzle_fouling_factor = py_min(1.0, e.nozzle_fouling_factor + 0.4);
            e.diffuser_fouling_factor = py_min(1.0, e.diffuser_fouling_factor + 0.5);
            e.nozzle_erosion_factor = py_min(1.0, e.nozzle_erosion_factor + 0.1);
            break;
        case MA_VACUUM_EJECTOR_NOZZLE_REPLACEMENT:
            e.nozzle_fouling_factor = 1.0; e.diffuser_fouling_factor = 1.0; e.nozzle_erosion_factor = 1.0;
            break;
        case MA_VACUUM_EJECTOR_INSPECTION: return MS_SUCCESS;
        case MA_ROUTINE_MAINTENANCE:
            e.nozzle_fouling_factor = py_min(1.0, e.nozzle_fouling_factor + 0.05);
            e.diffuser_fouling_factor = py_min(1.0, e.diffuser_fouling_factor + 0.05);
            break;
        default:
            e.nozzle_fouling_factor = 1.0; e.diffuser_fouling_factor = 1.0; e.nozzle_erosion_factor = 1.0;
            e.overall_performance_factor = 1.0;
            return MS_SUCCESS;
    }
    e.overall_performance_factor = (e.nozzle_fouling_factor * e.diffuser_fouling_factor * e.nozzle_erosion_factor);
    return MS_SUCCESS;
}

// One request: (target, action, arg).  Targets without a perform_maintenance restatement report
// MS_UNSUPPORTED_TARGET so the host can refuse instead of silently diverging.
NPS_HD int maintenance_apply_target(PlantState& st, const PlantParams& p, int target, int action, int arg) {
    if (target >= MT_PUMP0 && target < MT_PUMP0 + 4) return maintain_pump(st.fw.pump[target - MT_PUMP0], action, arg);
    if (target >= MT_SG0 && target < MT_SG0 + 3) return maintain_sg(st.sgs.sg[target - MT_SG0], action);
    if (target >= MT_STAGE0 && target < MT_STAGE0 + 14) return maintain_stage(st.turb.stage[target - MT_STAGE0], p, target - MT_STAGE0, action);
    if (target == MT_CONDENSER) return maintain_condenser(st.cond, p, action);
    if (target == MT_TURBINE) return maintain_turbine(st.turb, action);
    if (target == MT_SG_SYSTEM) return maintain_sg_system(st.sgs, action);
    if (target == MT_TURBINE_LUB) return maintain_turbine_lub(st.turb, action);
    if (target >= MT_EJECTOR0 && target < MT_EJECTOR0 + 2) return maintain_ejector(st.cond.ejector[target - MT_EJECTOR0], action);
    // FEE-001: FeedwaterPumpSystem has no perform_maintenance method (the one in pump_system.py:750 belongs to
    // FeedwaterPump), so _perform_maintenance_action reports "does not support maintenance": success False, no change
    if (target == MT_FW_SYSTEM) return MS_FAILED;
    // Not a target: SECONDARY-COMP-001-FW (EnhancedFeedwaterPhysics.perform_maintenance, feedwater/physics.py:984-1128).
    // No threshold of the reference template is bound to that id, so no automatic work order reaches it; its
    // system_cleaning branch scales the three members of PerformanceDiagnostics.wear_tracking separately, of which the
    // state log (and therefore PlantState) carries only the sum (rep.fw_diag_total_wear).  The host refuses the id.
    return MS_UNSUPPORTED_TARGET;
}

// The reference logs its state row AFTER maintenance (sim.py:209-223), and the report-only columns are computed from
// the component state at that moment, so they are refreshed here.
NPS_HD int maintenance_apply(PlantState& st, const PlantParams& p, int target, int action, int arg) {
    const int rc = maintenance_apply_target(st, p, target, action, arg);
    if (is_true(p.enable_secondary)) plant_report_state(st, p);
    return rc;
}

}  // namespace nps
