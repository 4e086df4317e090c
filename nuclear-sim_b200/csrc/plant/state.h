// Plant state and parameter records (all members FP64; flags/enums/ints are stored as exact
// small doubles so the whole record maps 1:1 onto a structure-of-arrays slab in HBM:
// field f of plant p lives at slab[f * n_plants + p]).
//
// RESTRICTED SYNTAX: nuclear_sim_b200/_layout.py parses this file to derive the flat field
// table (names, offsets) used by the host API, the oracle extractor and the tests.  Allowed
// member forms:  `double name;`  `double name[A];`  `double name[A][B];`
//                `StructName name;`  `StructName name[A];`
// One member per line, `//` comments only.
#pragma once

namespace nps {

// ---- primary side ------------------------------------------------------------------------
// ReactorState: systems/primary/__init__.py:48-106
struct PrimaryState {
    double neutron_flux;
    double reactivity;
    double precursors[6];
    double fuel_temperature;
    double coolant_temperature;
    double coolant_pressure;
    double coolant_flow_rate;
    double coolant_void_fraction;
    double steam_temperature;
    double steam_pressure;
    double steam_flow_rate;
    double feedwater_flow_rate;
    double control_rod_position;
    double steam_valve_position;
    double boron_concentration;
    double feedwater_pump_status;   // @discrete
    double feedwater_pump_speed;
    double feedwater_system_available;   // @discrete
    double feedwater_pump_power;
    double feedwater_num_running_pumps;   // @discrete
    double xenon_concentration;
    double iodine_concentration;
    double samarium_concentration;
    double burnable_poison_worth;
    double fuel_burnup;
    double power_level;
    double scram_status;   // @discrete
    // PrimaryReactorPhysics members: systems/primary/__init__.py:168-171
    double thermal_power_mw;
    double total_reactivity_pcm;
    double scram_activated;   // @discrete
    // ConstantHeatSource members: heat_sources/constant_heat_source.py:44-65
    double hs_setpoint_percent;
    double hs_current_power_mw;
    double hs_time;
    double hs_total_energy_mwh;
    double hs_filtered_noise_mw;
    double hs_raw_noise_mw;
};

// NuclearPlantSimulator members: simulator/core/sim.py:76-84,486-498
struct SimState {
    double time_minutes;
    double load_demand;
    double cooling_water_temp;
    double has_last_heat_removal_factor;   // @discrete
    double last_heat_removal_factor;
    double last_load_factor;
    double last_feedwater_flow_factor;
    double last_pump_reliability_factor;
};

// ---- water chemistry -----------------------------------------------------------------------
// WaterChemistry members: systems/secondary/water_chemistry.py:219-272 (+ pending effects :659-682)
struct WaterChemState {
    double ph;
    double iron_concentration;
    double copper_concentration;
    double silica_concentration;
    double dissolved_oxygen;
    double hardness;
    double total_dissolved_solids;
    double chloride;
    double alkalinity;
    double chlorine_residual;
    double antiscalant_concentration;
    double corrosion_inhibitor_level;
    double biocide_concentration;
    double water_aggressiveness;
    double particle_content;
    double scaling_tendency;
    double corrosion_tendency;
    double concentration_factor;
    double treatment_efficiency;
    double blowdown_rate;
    double operating_hours;
    double last_treatment_time;
    double chemistry_stability_factor;
    double pending_effects;   // @discrete
    double pend_ph_setpoint;
    double pend_ammonia_dose_rate;
    double pend_morpholine_dose_rate;
};

// ---- lubrication (shared base) ----------------------------------------------------------------
// BaseLubricationSystem members: systems/secondary/lubrication_base.py:131-175.
// component_wear / component_perf are indexed in the owning system's component order
// (6 used by the feedwater pump system, 5 by the turbine bearing system).
struct LubCore {
    double oil_level;
    double oil_temperature;
    double oil_pressure;
    double oil_contamination_level;
    double oil_moisture_content;
    double oil_acidity_number;
    double oil_viscosity_change;
    double oil_operating_hours;
    double antioxidant_level;
    double anti_wear_additive_level;
    double corrosion_inhibitor_level;
    double component_wear[6];
    double component_perf[6];
    double lubrication_effectiveness;
    double system_health_factor;
    double operating_hours;
};

// ---- feedwater -------------------------------------------------------------------------------
// FeedwaterPumpState + FeedwaterPump members (feedwater/pump_system.py:62-160,
// primary/coolant/pump_models.py:29-47) and the pump's FeedwaterPumpLubricationSystem
// (feedwater/pump_lubrication.py:202-215).  status: 0 RUNNING 1 STOPPED 2 STARTING 3 STOPPING 4 TRIPPED
struct FWPumpState {
    double speed_percent;
    double flow_rate;
    double status;   // @discrete
    double speed_setpoint;
    double power_consumption;
    double available;   // @discrete
    double trip_active;   // @discrete
    double trip_reason;   // @discrete
    double suction_pressure;
    double discharge_pressure;
    double npsh_available;
    double motor_temperature;
    double motor_current;
    double motor_voltage;
    double vibration_level;
    double differential_pressure;
    double cavitation_intensity;
    double cavitation_damage;
    double cavitation_time;
    double cavitation_noise_level;
    double flow_demand;
    double ic_applied;   // @discrete
    LubCore lub;
    double pump_load_factor;
    double cavitation_lubrication_effect;
    double seal_leakage_rate;
    double pump_efficiency_degradation;
    double pump_flow_degradation;
    double pump_head_degradation;
    double npsh_margin_degradation;
    double vibration_increase;
};

// EnhancedFeedwaterPhysics + ThreeElementControl + PerformanceDiagnostics + FeedwaterProtectionSystem
// (feedwater/physics.py:148-183, level_control.py:131-153, performance_monitoring.py:85-112,402-421,
//  protection_system.py:41-57,152-184)
struct FeedwaterState {
    FWPumpState pump[4];
    double total_flow_rate;
    double total_power_consumption;
    double system_efficiency;
    double system_availability;
    double performance_factor;
    double maintenance_factor;
    double operating_hours;
    double load_demand;
    double n_running_prev;   // @discrete
    double pump_system_available;   // @discrete
    double total_flow_demand;
    double lc_level_errors[3];
    double lc_level_integral_errors[3];
    double lc_previous_level_errors[3];
    double lc_quality_integral_error;
    double lc_control_performance;
    double cav_current_intensity;
    double cav_accumulated_damage;
    double cav_n_events;   // @discrete
    double cav_time_in_cavitation;
    double cav_acoustic_signature;
    double cav_noise_increase;
    double cav_induced_vibration;
    double cav_risk_score;
    double cav_predicted_damage_rate;
    double diag_health_score;
    double prot_npsh_low_alarm_active;   // @discrete
    double prot_npsh_low_low_trip_active;   // @discrete
    double prot_npsh_critical_trip_active;   // @discrete
    double prot_npsh_low_low_timer;
    double prot_timer_low_flow;
    double prot_timer_high_flow;
    double prot_timer_bearing_temp;
    double prot_timer_motor_temp;
    double prot_timer_vibration;
    double prot_system_trip_active;   // @discrete
};

// ---- steam generators ------------------------------------------------------------------------
// SteamGenerator members (steam_generator/steam_generator.py:88-113), its TSPFoulingModel
// (tsp_fouling_model.py:166-184, fouling_model_base.py:71-86; tsp_thickness[level][species] with
// species 0 magnetite 1 copper 2 silica 3 biological) and TubeInteriorFouling
// (tube_interior_fouling.py:62-75).  fouling_stage: 0 normal 1 significant 2 severe 3 critical.
// shutdown_reasons bitmask: 1 fouling 2 heat-transfer 4 pressure-drop 8 maldistribution 16 design-life.
struct SGState {
    double primary_inlet_temp;
    double primary_outlet_temp;
    double secondary_pressure;
    double secondary_temperature;
    double steam_quality;
    double water_level;
    double steam_void_fraction;
    double steam_flow_rate;
    double feedwater_flow_rate;
    double feedwater_temperature;
    double tube_wall_temp;
    double heat_transfer_rate;
    double overall_htc;
    double heat_flux;
    double thermal_efficiency;
    double tsp_operating_years;
    double tsp_last_cleaning_time;
    double tsp_total_cleaning_cycles;   // @discrete
    double tsp_fouling_fraction;
    double tsp_thickness[7][4];
    double tsp_fouling_stage;   // @discrete
    double tsp_heat_transfer_degradation;
    double tsp_pressure_drop_ratio;
    double tsp_flow_maldistribution;
    double tsp_cumulative_power_loss;
    double tsp_shutdown_required;   // @discrete
    double tsp_shutdown_reasons;   // @discrete
    double tsp_replacement_recommended;   // @discrete
    double tif_operating_years;
    double tif_last_cleaning_time;
    double tif_scale_thickness;
    double tif_scale_thermal_resistance;
    double tif_scale_formation_rate;
    double tif_comp[3];
    double tif_fouling_fraction;
    double tif_cumulative_performance_loss;
    double tif_replacement_recommended;   // @discrete
};

// EnhancedSteamGeneratorPhysics members: steam_generator/enhanced_physics.py:133-150
struct SGSystemState {
    SGState sg[3];
    double total_thermal_power;
    double total_steam_flow;
    double average_steam_pressure;
    double average_steam_temperature;
    double average_steam_quality;
    double system_availability;
    double operating_hours;
    double load_demand;
};

// ---- turbine -----------------------------------------------------------------------------------
// TurbineStage members: turbine/stage_system.py:57-96
struct TurbineStageState {
    double inlet_pressure;
    double inlet_temperature;
    double inlet_enthalpy;
    double inlet_entropy;
    double inlet_flow;
    double outlet_pressure;
    double outlet_temperature;
    double outlet_enthalpy;
    double outlet_flow;
    double actual_efficiency;
    double power_output;
    double enthalpy_drop;
    double extraction_flow;
    double extraction_pressure;
    double extraction_enthalpy;
    double blade_condition_factor;
    double fouling_factor;
    double deposit_thickness;
    double blade_wear_factor;
    double operating_hours;
    double efficiency_degradation;
    double loading_factor;
};

// BearingModel members: turbine/rotor_dynamics.py:57-81
struct TurbineBearingState {
    double current_load;
    double metal_temperature;
    double vibration_displacement;
    double operating_hours;
    double wear_factor;
    double efficiency_factor;
    double clearance_increase;
    double oil_temperature;
    double oil_flow_rate;
    double oil_contamination_level;
    double external_oil_temp;   // @discrete
};

// EnhancedTurbinePhysics (turbine/enhanced_physics.py:520-541) with its TurbineStageSystem
// (stage_system.py:693-704), RotorDynamicsModel (rotor_dynamics.py:806-829), VibrationMonitor
// (:597-622), MetalTemperatureTracker (enhanced_physics.py:52-71), TurbineProtectionSystem (:215-236)
// and TurbineBearingLubricationSystem (turbine_bearing_lubrication.py:191-202).
// trip_reasons bitmask: 1 overspeed 2 vibration 4 bearing-temp 8 thrust 16 low-vacuum 32 thermal-stress.
struct TurbineState {
    TurbineStageState stage[14];
    TurbineBearingState bearing[4];
    double ss_total_power_output;
    double ss_total_steam_flow;
    double ss_overall_efficiency;
    double ss_total_extraction_flow;
    double ss_system_efficiency;
    double ss_operating_hours;
    double rotor_speed;
    double rotor_acceleration;
    double friction_torque;
    double net_torque;
    double rotor_temperature;
    double thermal_expansion;
    double thermal_bow;
    double rotor_operating_hours;
    double overspeed_events;   // @discrete
    double vib_displacement_x;
    double vib_displacement_y;
    double vib_velocity_x;
    double vib_velocity_y;
    double vib_acceleration_x;
    double vib_acceleration_y;
    double vib_harmonic[3];
    double vib_displacement_alarm;   // @discrete
    double vib_velocity_alarm;   // @discrete
    double vib_acceleration_alarm;   // @discrete
    double vib_critical_speed_alarm;   // @discrete
    double th_rotor_temperatures[8];
    double th_casing_temperatures[6];
    double th_blade_temperatures[14];
    double th_rotor_gradients[7];
    double th_casing_gradients[5];
    double th_stress_levels[8];
    double th_max_thermal_stress;
    double th_temperature_rates[8];
    double th_thermal_shock_risk;
    double prot_timer_overspeed;
    double prot_timer_vibration;
    double prot_timer_bearing_temp;
    double prot_trip_active;   // @discrete
    double prot_trip_reasons;   // @discrete
    LubCore lub;
    double lub_turbine_efficiency_degradation;
    double lub_vibration_increase;
    double lub_oil_cooling_effectiveness;
    double lub_bearing_housing_temperature;
    double total_power_output;
    double overall_efficiency;
    double steam_rate;
    double heat_rate;
    double performance_factor;
    double availability_factor;
    double operating_hours;
    double load_demand;
};

// ---- condenser ---------------------------------------------------------------------------------
// SteamJetEjector members: condenser/vacuum_pump.py:56-95
struct EjectorState {
    double is_operating;   // @discrete
    double operating_hours;
    double current_capacity;
    double suction_pressure;
    double motive_steam_flow;
    double motive_steam_pressure_actual;
    double motive_steam_temp_actual;
    double first_stage_capacity;
    double second_stage_capacity;
    double intercondenser_load;
    double nozzle_fouling_factor;
    double diffuser_fouling_factor;
    double nozzle_erosion_factor;
    double overall_performance_factor;
    double steam_consumption_rate;
    double compression_ratio_actual;
    double entrainment_ratio;
};

// EnhancedCondenserPhysics (condenser/physics.py:536-560) with TubeDegradationModel (:56-71),
// AdvancedFoulingModel (:151-165), VacuumSystem + VacuumControlLogic (vacuum_system.py:46-58,256-306)
// and its own WaterChemistry (WAT-002).  lead/lag ejector: index, -1 = None.
struct CondenserState {
    WaterChemState wc;
    EjectorState ejector[2];
    double steam_inlet_pressure;
    double steam_inlet_temperature;
    double steam_inlet_flow;
    double steam_inlet_quality;
    double cooling_water_inlet_temp;
    double cooling_water_outlet_temp;
    double cooling_water_flow;
    double heat_rejection_rate;
    double overall_htc;
    double condensate_temperature;
    double condensate_flow;
    double thermal_performance_factor;
    double operating_hours;
    double td_active_tube_count;
    double td_plugged_tube_count;
    double td_average_wall_thickness;
    double td_tube_leak_rate;
    double td_vibration_damage_accumulation;
    double td_corrosion_damage_accumulation;
    double td_operating_hours;
    double td_area_factor;
    double td_pressure_drop_factor;
    double fl_biofouling_thickness;
    double fl_scale_thickness;
    double fl_corrosion_product_thickness;
    double fl_distribution_factor;
    double fl_time_since_cleaning;
    double fl_total_fouling_resistance;
    double vs_condenser_pressure;
    double vs_air_partial_pressure;
    double vs_steam_partial_pressure;
    double vs_total_air_removal_rate;
    double vs_total_steam_consumption;
    double vs_current_air_leakage;
    double vs_air_mass_in_condenser;
    double vs_motive_steam_pressure;
    double vs_motive_steam_temperature;
    double vs_motive_steam_available;   // @discrete
    double vs_system_efficiency;
    double vs_operating_hours;
    double vs_alarm_high_pressure;   // @discrete
    double vs_alarm_low_motive_pressure;   // @discrete
    double vs_alarm_ejector_failure;   // @discrete
    double vs_alarm_excessive_air_leakage;   // @discrete
    double vs_trip_high_pressure;   // @discrete
    double vc_lead_ejector;   // @discrete
    double vc_lag_ejector;   // @discrete
    double vc_rotation_timer;
};

// ---- pH control --------------------------------------------------------------------------------
// PHControllerState + PHController/PHControlSystem internals: secondary/ph_control_system.py:139-190,
// 199-217, 521-532.  control_mode: 0 AUTO 1 MANUAL 2 FAILED 3 MAINTENANCE.
// dev_hist is the sliding window of the last 100 |pH error| values (oldest at dev_head).
struct PHControlState {
    double control_mode;   // @discrete
    double controller_enabled;   // @discrete
    double manual_output;
    double measured_ph;
    double ph_setpoint;
    double ph_error;
    double controller_output;
    double ammonia_dose_rate;
    double morpholine_dose_rate;
    double proportional_term;
    double integral_term;
    double derivative_term;
    double previous_error;
    double integral_sum;
    double ammonia_tank_level;
    double morpholine_tank_level;
    double ammonia_supply_available;   // @discrete
    double morpholine_supply_available;   // @discrete
    double ammonia_pump_status;   // @discrete
    double morpholine_pump_status;   // @discrete
    double ph_sensor_status;   // @discrete
    double ph_low_alarm;   // @discrete
    double ph_high_alarm;   // @discrete
    double low_chemical_alarm;   // @discrete
    double equipment_failure_alarm;   // @discrete
    double control_deviation_rms;
    double chemical_consumption_rate;
    double time_in_control;
    double operating_hours;
    double tic_initialized;   // @discrete
    double tic_sum;
    double tic_total_time;
    double total_chemical_consumed;
    double control_actions_count;   // @discrete
    double dev_count;   // @discrete
    double dev_head;   // @discrete
    double dev_hist[100];
};

// ---- secondary system orchestrator -------------------------------------------------------------
// SecondaryReactorPhysics members: systems/secondary/__init__.py:296-337,384-442
struct SecondaryState {
    double total_steam_flow;
    double total_heat_transfer;
    double electrical_power_output;
    double thermal_efficiency;
    double total_feedwater_flow;
    double load_demand;
    double feedwater_temperature;
    double cooling_water_temperature;
    double operating_hours;
    double total_system_heat_rejection;
    double has_previous_feedwater_temp;   // @discrete
    double previous_feedwater_temp;
    double has_previous_sg_conditions;   // @discrete
    double prev_sg_levels[3];
    double prev_sg_pressures[3];
    double prev_sg_steam_flows[3];
    double prev_sg_steam_qualities[3];
    double sg_avg_pressure;
    double sg_avg_temperature;
    double condenser_pressure;
    double heat_rate_kj_kwh;
    double power_reduction_factor;
};

// ---- report-only quantities ------------------------------------------------------------------------
// Values the reference computes on the fly in its get_state_dict() methods (or keeps in HeatFlowTracker) for the state
// log and nowhere else; nothing on the step path reads them.  They are written on the last fused substep only
// (StepInput.emit_outputs) and refreshed by the maintenance kernel, so a logged row sees what the reference's
// collect_states (simulator/state/state_manager.py:152-211, called after maintenance: sim.py:209-223) would see.
//   sg_*            SteamGenerator.get_state_dict: steam_generator/steam_generator.py:944-986
//   sgs_*           EnhancedSteamGeneratorPhysics.get_state_dict: steam_generator/enhanced_physics.py:689-723
//   fwp_*           FeedwaterPumpLubricationSystem properties: feedwater/pump_lubrication.py:225-233,1619-1620
//   fw_avg_* ...    EnhancedFeedwaterPhysics.get_state_dict: feedwater/physics.py:1142-1145 (the SG conditions it was
//                   LAST GIVEN, i.e. the previous step's: systems/secondary/__init__.py:447-491)
//   fw_diag_*       PerformanceDiagnostics.get_state_dict: feedwater/performance_monitoring.py:650-662
//                   (fw_diag_total_wear: WearTrackingModel is never advanced on the lubrication-system path,
//                   performance_monitoring.py:473-491, so it keeps its initial value; the step does not touch it)
//   fw_prot_*       FeedwaterProtectionSystem.get_state_dict: feedwater/protection_system.py:778
//   hf_*            HeatFlowTracker: systems/secondary/__init__.py:680-744, heat_flow_tracker.py:248-327
struct ReportState {
    double sg_primary_flow_restriction_factor[3];
    double sg_secondary_flow_restriction_factor[3];
    double sg_max_primary_flow_capacity[3];
    double sg_max_steam_flow_capacity[3];
    double sg_max_feedwater_flow_capacity[3];
    double sg_fouling_energy_penalty_mw[3];
    double sg_total_pump_power_mw[3];
    double sgs_total_fouling_impact;
    double sgs_fouling_maintenance_needed;
    double fwp_efficiency_factor[4];
    double fwp_flow_factor[4];
    double fw_avg_sg_level;
    double fw_avg_sg_pressure;
    double fw_total_steam_flow;
    double fw_avg_steam_quality;
    double fw_diag_maintenance_urgency;
    double fw_diag_total_wear;
    double fw_prot_active_alarms_count;   // @discrete
    double hf_steam_enthalpy_flow;
    double hf_turbine_work_output;
    double hf_condenser_heat_rejection;
    double hf_net_electrical_output;
    double hf_overall_efficiency;
    double hf_energy_balance_error;
    double hf_energy_balance_percent;
};

struct PlantState {
    PrimaryState pri;
    SimState sim;
    WaterChemState wc_main;
    FeedwaterState fw;
    SGSystemState sgs;
    TurbineState turb;
    CondenserState cond;
    PHControlState ph;
    SecondaryState sec;
    ReportState rep;
};

// ---- batch-uniform parameters ------------------------------------------------------------
struct PlantParams {
    double dt;
    double heat_source_type;
    double rated_power_mw;
    double noise_enabled;
    double noise_std_percent;
    double noise_filter_time_constant;
    double enable_secondary;
    // feedwater (feedwater/config.py; feedwater/physics.py:101-129)
    double fw_num_sg;
    double fw_design_total_flow;
    double fw_design_sg_level;
    double fw_design_pressure;
    double fw_design_feedwater_temperature;
    double fw_auto_level_control;
    double fw_lc_level_control_weight;
    double fw_lc_feedwater_flow_weight;
    double fw_lc_quality_gain;
    double fwp_rated_flow;
    double fwp_rated_power;
    double fw_prot_low_suction_pressure_trip;
    double fw_prot_high_discharge_pressure_trip;
    double fw_prot_low_flow_trip;
    // steam generators (steam_generator/config.py:179-281) and the never-updated WAT-001 chemistry
    double sg_heat_transfer_area;
    double sg_primary_design_flow;
    double sg_primary_htc;
    double sg_secondary_htc;
    double sg_design_pressure_secondary;
    double sg_tube_wall_thickness;
    double sg_tube_conductivity;
    double sg_design_thermal_power_per_sg;
    double sg_secondary_design_flow;
    double sg_secondary_water_mass;
    double sg_design_steam_flow_per_sg;
    double sg_design_feedwater_flow_per_sg;
    double sg_tube_inner_diameter;
    double sg_tube_count;
    double sg_design_total_steam_flow;
    double sg_auto_load_balancing;
    double sgwc_iron;
    double sgwc_copper;
    double sgwc_silica;
    double sgwc_ph;
    double sgwc_dissolved_oxygen;
    // turbine (turbine/config.py:37-195,260-311; per-stage design tables from stage_system.py:706-758)
    double ts_design_inlet_pressure[14];
    double ts_design_outlet_pressure[14];
    double ts_design_steam_flow[14];
    double ts_design_efficiency[14];
    double ts_has_extraction[14];
    double ts_max_extraction_flow[14];
    double ts_min_extraction_flow[14];
    double ts_fouling_rate;
    double ts_erosion_rate;
    double ts_deposit_buildup_rate;
    double rd_rotor_inertia;
    double rd_max_speed;
    double rd_thermal_expansion_coefficient;
    double rd_rotor_length;
    double rd_thermal_bow_limit;
    double rd_rotor_mass;
    double rd_design_load_capacity;
    double rd_bearing_clearance;
    double rd_bearing_stiffness;
    double rd_bearing_damping;
    double rd_friction_coefficient;
    double rd_first_critical_speed;
    double rd_second_critical_speed;
    double rd_critical_speed_margin;
    double rd_displacement_alarm;
    double rd_velocity_alarm;
    double rd_acceleration_alarm;
    double tt_thermal_time_constant;
    double tt_thermal_expansion_coeff;
    double tt_elastic_modulus;
    double tt_max_thermal_gradient;
    double tt_max_thermal_stress;
    double tp_overspeed_trip;
    double tp_overspeed_delay;
    double tp_vibration_trip;
    double tp_vibration_delay;
    double tp_bearing_temp_trip;
    double tp_bearing_temp_delay;
    double tp_thrust_bearing_trip;
    double tp_low_vacuum_trip;
    double tp_max_thermal_stress;
    double tl_contamination_limit;
    double tl_acidity_limit;
    double tl_moisture_limit;
    double tl_viscosity_change_limit;
    // condenser (condenser/config.py:37-215)
    double cd_design_heat_duty;
    double cd_design_cooling_water_flow;
    double cd_heat_transfer_area;
    double cd_tube_inner_diameter;
    double cd_tube_wall_thickness;
    double cd_steam_side_htc;
    double cd_water_side_htc;
    double cd_tube_wall_conductivity;
    double cd_td_initial_tube_count;
    double cd_td_tube_failure_rate;
    double cd_td_vibration_damage_threshold;
    double cd_td_wall_thickness_initial;
    double cd_td_wall_thickness_minimum;
    double cd_td_corrosion_rate;
    double cd_fl_biofouling_base_rate;
    double cd_fl_biofouling_temp_coefficient;
    double cd_fl_biofouling_nutrient_factor;
    double cd_fl_scale_base_rate;
    double cd_fl_scale_hardness_coefficient;
    double cd_fl_scale_temp_coefficient;
    double cd_fl_corrosion_base_rate;
    double cd_fl_corrosion_oxygen_coefficient;
    double cd_fl_corrosion_ph_optimum;
    double cd_vs_base_air_leakage;
    double cd_vs_auto_start_pressure;
    double cd_vs_auto_stop_pressure;
    double cd_vs_rotation_interval;
    double cd_vs_leakage_degradation_rate;
    double cd_vs_condenser_volume;
    double cd_vs_steam_pressure_drop;
    double cd_vs_high_pressure_alarm;
    double cd_vs_high_pressure_trip;
    double cd_vs_low_motive_pressure_alarm;
    double ej_two_stage[2];
    double ej_design_capacity[2];
    double ej_design_suction_pressure[2];
    double ej_motive_steam_pressure[2];
    double ej_motive_steam_temperature[2];
    double ej_base_steam_consumption[2];
    double ej_steam_consumption_exponent[2];
    double ej_pressure_effect_coefficient[2];
    double ej_min_suction_pressure[2];
    double ej_max_suction_pressure[2];
    double ej_min_motive_pressure[2];
    double ej_intercondenser_pressure[2];
    double ej_nozzle_fouling_rate[2];
    double ej_diffuser_fouling_rate[2];
    double ej_erosion_rate[2];
};

}  // namespace nps
