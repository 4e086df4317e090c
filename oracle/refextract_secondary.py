"""TEST INFRASTRUCTURE — maps the reference's secondary-side object graph onto PlantState /
PlantParams field names (nuclear_sim_b200/_layout.py).  One function per subsystem; each reads
live attributes of the reference objects, nothing is computed here."""

PUMP_STATUS = {"running": 0.0, "stopped": 1.0, "starting": 2.0, "stopping": 3.0, "tripped": 4.0}

FW_LUB_COMPONENTS = ["impeller", "motor_bearings", "pump_bearings", "thrust_bearing", "mechanical_seals",
                     "coupling_system"]

PUMP_TRIP_CODES = {
    "": 0, "Low Flow": 1, "NPSH Violation": 2, "Low Suction Pressure": 3, "High Discharge Pressure": 4,
    "Steam Generator High Level": 5, "Severe Cavitation": 6, "Cavitation Damage Limit": 7,
    "Critical NPSH Violation": 8, "Lubrication: Very Low Oil Level": 9, "Lubrication: Low Oil Level": 10,
    "Lubrication: Oil System Overfill": 11, "Lubrication: Impeller Excessive Wear": 12,
    "Lubrication: Motor Bearings Excessive Wear": 13, "Lubrication: Pump Bearings Excessive Wear": 14,
    "Lubrication: Thrust Bearing Excessive Wear": 15, "Lubrication: Mechanical Seals Excessive Wear": 16,
    "Lubrication: Coupling System Excessive Wear": 17, "Lubrication: Excessive Seal Leakage": 18,
    "Lubrication: Combined Wear Limit": 19, "Lubrication: Performance Degradation": 20,
}


def water_chem(w, pre, d):
    for name in ("ph", "iron_concentration", "copper_concentration", "silica_concentration", "dissolved_oxygen",
                 "hardness", "total_dissolved_solids", "chloride", "alkalinity", "chlorine_residual",
                 "antiscalant_concentration", "corrosion_inhibitor_level", "biocide_concentration",
                 "water_aggressiveness", "particle_content", "scaling_tendency", "corrosion_tendency",
                 "concentration_factor", "treatment_efficiency", "blowdown_rate", "operating_hours",
                 "last_treatment_time", "chemistry_stability_factor"):
        d[pre + name] = float(getattr(w, name))
    pend = getattr(w, "_pending_chemistry_effects", None) or {}
    phc = pend.get("ph_control")
    d[pre + "pending_effects"] = 1.0 if phc else 0.0
    d[pre + "pend_ph_setpoint"] = float(phc.get("ph_setpoint", 0.0)) if phc else 0.0
    d[pre + "pend_ammonia_dose_rate"] = float(phc.get("ammonia_dose_rate", 0.0)) if phc else 0.0
    d[pre + "pend_morpholine_dose_rate"] = float(phc.get("morpholine_dose_rate", 0.0)) if phc else 0.0


def lub_core(L, comps, pre, d):
    for name in ("oil_level", "oil_temperature", "oil_pressure", "oil_contamination_level", "oil_moisture_content",
                 "oil_acidity_number", "oil_viscosity_change", "oil_operating_hours", "antioxidant_level",
                 "anti_wear_additive_level", "corrosion_inhibitor_level", "lubrication_effectiveness",
                 "system_health_factor", "operating_hours"):
        d[pre + name] = float(getattr(L, name))
    for i in range(6):
        if i < len(comps):
            d[f"{pre}component_wear[{i}]"] = float(L.component_wear[comps[i]])
            d[f"{pre}component_perf[{i}]"] = float(L.component_performance_factors[comps[i]])
        else:
            d[f"{pre}component_wear[{i}]"] = 0.0
            d[f"{pre}component_perf[{i}]"] = 0.0


def feedwater(fw, d):
    P = "fw."
    pumps = list(fw.pump_system.pumps.values())
    assert len(pumps) == 4
    for k, pump in enumerate(pumps):
        pre = f"{P}pump[{k}]."
        st = pump.state
        for name in ("speed_percent", "flow_rate", "speed_setpoint", "power_consumption", "suction_pressure",
                     "discharge_pressure", "npsh_available", "motor_temperature", "motor_current", "motor_voltage",
                     "vibration_level", "differential_pressure", "cavitation_intensity", "cavitation_damage",
                     "cavitation_time", "cavitation_noise_level"):
            d[pre + name] = float(getattr(st, name))
        d[pre + "status"] = PUMP_STATUS[st.status.value]
        d[pre + "available"] = float(st.available)
        d[pre + "trip_active"] = float(st.trip_active)
        reason = st.trip_reason
        if reason.startswith("NPSH Violation"):
            reason = "NPSH Violation"
        d[pre + "trip_reason"] = float(PUMP_TRIP_CODES[reason])
        d[pre + "flow_demand"] = float(pump.flow_demand)
        d[pre + "ic_applied"] = float(hasattr(pump, "_initial_conditions_applied"))
        L = pump.lubrication_system
        lub_core(L, FW_LUB_COMPONENTS, pre + "lub.", d)
        for name in ("pump_load_factor", "cavitation_lubrication_effect", "seal_leakage_rate",
                     "pump_efficiency_degradation", "pump_flow_degradation", "pump_head_degradation",
                     "npsh_margin_degradation", "vibration_increase"):
            d[pre + name] = float(getattr(L, name))
    for name in ("total_flow_rate", "total_power_consumption", "system_efficiency", "performance_factor",
                 "maintenance_factor", "operating_hours", "load_demand"):
        d[P + name] = float(getattr(fw, name))
    d[P + "system_availability"] = float(fw.system_availability)
    d[P + "n_running_prev"] = float(len(fw.pump_system.running_pumps))
    d[P + "pump_system_available"] = float(fw.pump_system.system_available)
    lc = fw.level_control
    hist = lc.flow_demand_history
    d[P + "total_flow_demand"] = float(hist[-1]) if hist else 0.0
    for i in range(3):
        d[f"{P}lc_level_errors[{i}]"] = float(lc.level_errors[i])
        d[f"{P}lc_level_integral_errors[{i}]"] = float(lc.level_integral_errors[i])
        d[f"{P}lc_previous_level_errors[{i}]"] = float(lc.previous_level_errors[i])
    d[P + "lc_quality_integral_error"] = float(lc.quality_compensator.quality_integral_error)
    d[P + "lc_control_performance"] = float(lc.control_performance)
    cm = fw.diagnostics.cavitation_model
    d[P + "cav_current_intensity"] = float(cm.current_intensity)
    d[P + "cav_accumulated_damage"] = float(cm.accumulated_damage)
    d[P + "cav_n_events"] = float(len(cm.cavitation_events))
    d[P + "cav_time_in_cavitation"] = float(cm.time_in_cavitation)
    d[P + "cav_acoustic_signature"] = float(cm.acoustic_signature)
    d[P + "cav_noise_increase"] = float(cm.cavitation_noise_increase)
    d[P + "cav_induced_vibration"] = float(cm.cavitation_induced_vibration)
    d[P + "cav_risk_score"] = float(cm.cavitation_risk_score)
    d[P + "cav_predicted_damage_rate"] = float(cm.predicted_damage_rate)
    d[P + "diag_health_score"] = float(fw.diagnostics.overall_health_score)
    pr = fw.protection_system
    npr = pr.npsh_protection
    d[P + "prot_npsh_low_alarm_active"] = float(npr.npsh_low_alarm_active)
    d[P + "prot_npsh_low_low_trip_active"] = float(npr.npsh_low_low_trip_active)
    d[P + "prot_npsh_critical_trip_active"] = float(npr.npsh_critical_trip_active)
    d[P + "prot_npsh_low_low_timer"] = float(npr.npsh_low_low_timer)
    d[P + "prot_timer_low_flow"] = float(pr.trip_timers["low_flow"])
    d[P + "prot_timer_high_flow"] = float(pr.trip_timers["high_flow"])
    d[P + "prot_timer_bearing_temp"] = float(pr.trip_timers["bearing_temp"])
    d[P + "prot_timer_motor_temp"] = float(pr.trip_timers["motor_temp"])
    d[P + "prot_timer_vibration"] = float(pr.trip_timers["vibration"])
    d[P + "prot_system_trip_active"] = float(pr.system_trip_active)


def feedwater_params(fw, d):
    cfg = fw.config
    d["fw_num_sg"] = float(cfg.num_steam_generators)
    d["fw_design_total_flow"] = float(cfg.design_total_flow)
    d["fw_design_sg_level"] = float(cfg.design_sg_level)
    d["fw_design_pressure"] = float(cfg.design_pressure)
    d["fw_design_feedwater_temperature"] = float(cfg.design_feedwater_temperature)
    d["fw_auto_level_control"] = float(cfg.auto_level_control)
    cc = fw.level_control.config
    d["fw_lc_level_control_weight"] = float(getattr(cc, "level_control_weight", 0.4))
    d["fw_lc_feedwater_flow_weight"] = float(getattr(cc, "feedwater_flow_weight", 0.1))
    d["fw_lc_quality_gain"] = float(fw.level_control.quality_compensator.quality_control_gain)
    pump = next(iter(fw.pump_system.pumps.values()))
    d["fwp_rated_flow"] = float(pump.config.rated_flow)
    d["fwp_rated_power"] = float(pump.config.rated_power)
    pc = fw.protection_system.config
    d["fw_prot_low_suction_pressure_trip"] = float(pc.low_suction_pressure_trip)
    d["fw_prot_high_discharge_pressure_trip"] = float(pc.high_discharge_pressure_trip)
    d["fw_prot_low_flow_trip"] = float(pc.low_flow_trip)


FOULING_STAGE = {"normal": 0.0, "significant": 1.0, "severe": 2.0, "critical": 3.0}
SHUTDOWN_BITS = {"critical_tsp_fouling": 1, "excessive_heat_transfer_loss": 2, "excessive_pressure_drop": 4,
                 "severe_flow_maldistribution": 8, "tube_integrity_risk": 0, "design_life_exceeded": 16}


def steam_generators(sgs, d):
    P = "sgs."
    assert len(sgs.steam_generators) == 3
    for i, sg in enumerate(sgs.steam_generators):
        pre = f"{P}sg[{i}]."
        for name in ("primary_inlet_temp", "primary_outlet_temp", "secondary_pressure", "secondary_temperature",
                     "steam_quality", "water_level", "steam_void_fraction", "steam_flow_rate", "feedwater_flow_rate",
                     "feedwater_temperature", "tube_wall_temp", "heat_transfer_rate", "overall_htc", "heat_flux"):
            d[pre + name] = float(getattr(sg, name))
        d[pre + "thermal_efficiency"] = float(sg.heat_transfer_rate / sg.config.design_thermal_power_per_sg)
        t = sg.tsp_fouling
        d[pre + "tsp_operating_years"] = float(t.operating_years)
        d[pre + "tsp_last_cleaning_time"] = float(t.last_cleaning_time)
        d[pre + "tsp_total_cleaning_cycles"] = float(t.total_cleaning_cycles)
        d[pre + "tsp_fouling_fraction"] = float(t.fouling_fraction)
        for lv in range(7):
            d[f"{pre}tsp_thickness[{lv}][0]"] = float(t.deposits.magnetite_thickness[lv])
            d[f"{pre}tsp_thickness[{lv}][1]"] = float(t.deposits.copper_thickness[lv])
            d[f"{pre}tsp_thickness[{lv}][2]"] = float(t.deposits.silica_thickness[lv])
            d[f"{pre}tsp_thickness[{lv}][3]"] = float(t.deposits.biological_thickness[lv])
        d[pre + "tsp_fouling_stage"] = FOULING_STAGE[t.fouling_stage.value]
        d[pre + "tsp_heat_transfer_degradation"] = float(t.heat_transfer_degradation)
        d[pre + "tsp_pressure_drop_ratio"] = float(t.pressure_drop_ratio)
        d[pre + "tsp_flow_maldistribution"] = float(t.flow_maldistribution)
        d[pre + "tsp_cumulative_power_loss"] = float(t.cumulative_power_loss)
        d[pre + "tsp_shutdown_required"] = float(t.shutdown_required)
        d[pre + "tsp_shutdown_reasons"] = float(sum(SHUTDOWN_BITS[r.value] for r in t.shutdown_reasons))
        d[pre + "tsp_replacement_recommended"] = float(t.replacement_recommended)
        f = sg.tube_interior_fouling
        d[pre + "tif_operating_years"] = float(f.operating_years)
        d[pre + "tif_last_cleaning_time"] = float(f.last_cleaning_time)
        d[pre + "tif_scale_thickness"] = float(f.scale_thickness)
        d[pre + "tif_scale_thermal_resistance"] = float(f.scale_thermal_resistance)
        d[pre + "tif_scale_formation_rate"] = float(f.scale_formation_rate)
        d[pre + "tif_comp[0]"] = float(f.scale_composition["iron_oxide"])
        d[pre + "tif_comp[1]"] = float(f.scale_composition["crud_deposits"])
        d[pre + "tif_comp[2]"] = float(f.scale_composition["corrosion_products"])
        d[pre + "tif_fouling_fraction"] = float(f.fouling_fraction)
        d[pre + "tif_cumulative_performance_loss"] = float(f.cumulative_performance_loss)
        d[pre + "tif_replacement_recommended"] = float(f.replacement_recommended)
    for name in ("total_thermal_power", "total_steam_flow", "average_steam_pressure", "average_steam_temperature",
                 "average_steam_quality", "operating_hours", "load_demand"):
        d[P + name] = float(getattr(sgs, name))
    d[P + "system_availability"] = float(sgs.system_availability)


def steam_generator_params(sgs, d):
    c = sgs.steam_generators[0].config
    d["sg_heat_transfer_area"] = float(c.heat_transfer_area_per_sg)
    d["sg_primary_design_flow"] = float(c.primary_design_flow)
    d["sg_primary_htc"] = float(c.primary_htc)
    d["sg_secondary_htc"] = float(c.secondary_htc)
    d["sg_design_pressure_secondary"] = float(c.design_pressure_secondary)
    d["sg_tube_wall_thickness"] = float(c.tube_wall_thickness)
    d["sg_tube_conductivity"] = float(c.tube_material_conductivity)
    d["sg_design_thermal_power_per_sg"] = float(c.design_thermal_power_per_sg)
    d["sg_secondary_design_flow"] = float(c.secondary_design_flow)
    d["sg_secondary_water_mass"] = float(c.secondary_water_mass)
    d["sg_design_steam_flow_per_sg"] = float(c.design_steam_flow_per_sg)
    d["sg_design_feedwater_flow_per_sg"] = float(c.design_feedwater_flow_per_sg)
    d["sg_tube_inner_diameter"] = float(c.tube_inner_diameter)
    d["sg_tube_count"] = float(c.tube_count_per_sg)
    d["sg_design_total_steam_flow"] = float(sgs.config.design_total_steam_flow)
    d["sg_auto_load_balancing"] = float(sgs.config.auto_load_balancing)
    w = sgs.steam_generators[0].tsp_fouling.water_chemistry
    d["sgwc_iron"] = float(w.iron_concentration)
    d["sgwc_copper"] = float(w.copper_concentration)
    d["sgwc_silica"] = float(w.silica_concentration)
    d["sgwc_ph"] = float(w.ph)
    d["sgwc_dissolved_oxygen"] = float(w.dissolved_oxygen)


TURB_LUB_COMPONENTS = ["hp_journal_bearing", "lp_journal_bearing", "thrust_bearing", "seal_oil_system", "oil_coolers"]
TURB_TRIP_BITS = {"Overspeed": 1, "High Vibration": 2, "High Bearing Temperature": 4,
                  "Thrust Bearing Displacement": 8, "Low Vacuum": 16, "High Thermal Stress": 32}


def _sorted_stages(turb):
    st = sorted(turb.stage_system.stages.items(), key=lambda x: x[1].config.design_inlet_pressure, reverse=True)
    ids = [k for k, _ in st]
    assert ids == [f"HP-{i+1}" for i in range(8)] + [f"LP-{i+1}" for i in range(6)], ids
    return [v for _, v in st]


def turbine(turb, d):
    P = "turb."
    for k, stg in enumerate(_sorted_stages(turb)):
        pre = f"{P}stage[{k}]."
        for name in ("inlet_pressure", "inlet_temperature", "inlet_enthalpy", "inlet_entropy", "inlet_flow",
                     "outlet_pressure", "outlet_temperature", "outlet_enthalpy", "outlet_flow", "actual_efficiency",
                     "power_output", "enthalpy_drop", "extraction_flow", "extraction_pressure", "extraction_enthalpy",
                     "blade_condition_factor", "fouling_factor", "deposit_thickness", "blade_wear_factor",
                     "operating_hours", "efficiency_degradation", "loading_factor"):
            d[pre + name] = float(getattr(stg, name))
    rd = turb.rotor_dynamics
    bearings = list(rd.bearings.items())
    assert [b for b, _ in bearings] == ["TB-001", "TB-002", "TB-003", "TB-004"]
    assert [b.config.bearing_type for _, b in bearings] == ["journal", "journal", "thrust", "journal"]
    for b, (_, br) in enumerate(bearings):
        pre = f"{P}bearing[{b}]."
        for name in ("current_load", "metal_temperature", "vibration_displacement", "operating_hours", "wear_factor",
                     "efficiency_factor", "clearance_increase", "oil_temperature", "oil_flow_rate",
                     "oil_contamination_level"):
            d[pre + name] = float(getattr(br, name))
        d[pre + "external_oil_temp"] = float(br.external_oil_temp)
    ss = turb.stage_system
    d[P + "ss_total_power_output"] = float(ss.total_power_output)
    d[P + "ss_total_steam_flow"] = float(ss.total_steam_flow)
    d[P + "ss_overall_efficiency"] = float(ss.overall_efficiency)
    d[P + "ss_total_extraction_flow"] = float(ss.total_extraction_flow)
    d[P + "ss_system_efficiency"] = float(ss.system_efficiency)
    d[P + "ss_operating_hours"] = float(ss.operating_hours)
    for name in ("rotor_speed", "rotor_acceleration", "friction_torque", "net_torque", "rotor_temperature",
                 "thermal_expansion", "thermal_bow", "overspeed_events"):
        d[P + name] = float(getattr(rd, name))
    d[P + "rotor_operating_hours"] = float(rd.operating_hours)
    vm = rd.vibration_monitor
    for a, b in (("vib_displacement_x", "displacement_x"), ("vib_displacement_y", "displacement_y"),
                 ("vib_velocity_x", "velocity_x"), ("vib_velocity_y", "velocity_y"),
                 ("vib_acceleration_x", "acceleration_x"), ("vib_acceleration_y", "acceleration_y"),
                 ("vib_displacement_alarm", "displacement_alarm"), ("vib_velocity_alarm", "velocity_alarm"),
                 ("vib_acceleration_alarm", "acceleration_alarm"), ("vib_critical_speed_alarm", "critical_speed_alarm")):
        d[P + a] = float(getattr(vm, b))
    for i in range(3):
        d[f"{P}vib_harmonic[{i}]"] = float(vm.harmonic_amplitudes[i])
    th = turb.thermal_tracker
    for i in range(8):
        d[f"{P}th_rotor_temperatures[{i}]"] = float(th.rotor_temperatures[i])
        d[f"{P}th_stress_levels[{i}]"] = float(th.thermal_stress_levels[i])
        d[f"{P}th_temperature_rates[{i}]"] = float(th.temperature_rates[i])
    for i in range(6):
        d[f"{P}th_casing_temperatures[{i}]"] = float(th.casing_temperatures[i])
    for i in range(14):
        d[f"{P}th_blade_temperatures[{i}]"] = float(th.blade_temperatures[i])
    for i in range(7):
        d[f"{P}th_rotor_gradients[{i}]"] = float(th.rotor_gradients[i])
    for i in range(5):
        d[f"{P}th_casing_gradients[{i}]"] = float(th.casing_gradients[i])
    d[P + "th_max_thermal_stress"] = float(th.max_thermal_stress)
    d[P + "th_thermal_shock_risk"] = float(th.thermal_shock_risk)
    pr = turb.protection_system
    d[P + "prot_timer_overspeed"] = float(pr.trip_timers["overspeed"])
    d[P + "prot_timer_vibration"] = float(pr.trip_timers["vibration"])
    d[P + "prot_timer_bearing_temp"] = float(pr.trip_timers["bearing_temp"])
    d[P + "prot_trip_active"] = float(pr.trip_active)
    d[P + "prot_trip_reasons"] = float(sum(TURB_TRIP_BITS[r] for r in pr.trip_reasons))
    L = turb.bearing_lubrication_system
    lub_core(L, TURB_LUB_COMPONENTS, P + "lub.", d)
    d[P + "lub_turbine_efficiency_degradation"] = float(L.turbine_efficiency_degradation)
    d[P + "lub_vibration_increase"] = float(L.vibration_increase)
    d[P + "lub_oil_cooling_effectiveness"] = float(L.oil_cooling_effectiveness)
    d[P + "lub_bearing_housing_temperature"] = float(L.bearing_housing_temperature)
    for name in ("total_power_output", "overall_efficiency", "steam_rate", "heat_rate", "performance_factor",
                 "availability_factor", "operating_hours", "load_demand"):
        d[P + name] = float(getattr(turb, name))


def turbine_params(turb, d):
    for k, stg in enumerate(_sorted_stages(turb)):
        c = stg.config
        d[f"ts_design_inlet_pressure[{k}]"] = float(c.design_inlet_pressure)
        d[f"ts_design_outlet_pressure[{k}]"] = float(c.design_outlet_pressure)
        d[f"ts_design_steam_flow[{k}]"] = float(c.design_steam_flow)
        d[f"ts_design_efficiency[{k}]"] = float(c.design_efficiency)
        d[f"ts_has_extraction[{k}]"] = float(c.has_extraction)
        d[f"ts_max_extraction_flow[{k}]"] = float(c.max_extraction_flow)
        d[f"ts_min_extraction_flow[{k}]"] = float(c.min_extraction_flow)
        assert getattr(c, "min_stage_efficiency", 0.7) == 0.7
    c0 = _sorted_stages(turb)[0].config
    d["ts_fouling_rate"] = float(c0.fouling_rate)
    d["ts_erosion_rate"] = float(c0.erosion_rate)
    d["ts_deposit_buildup_rate"] = float(c0.deposit_buildup_rate)
    rc = turb.rotor_dynamics.config
    for a, b in (("rd_rotor_inertia", "rotor_inertia"), ("rd_max_speed", "max_speed"),
                 ("rd_thermal_expansion_coefficient", "thermal_expansion_coefficient"), ("rd_rotor_length", "rotor_length"),
                 ("rd_thermal_bow_limit", "thermal_bow_limit"), ("rd_rotor_mass", "rotor_mass"),
                 ("rd_first_critical_speed", "first_critical_speed"), ("rd_second_critical_speed", "second_critical_speed"),
                 ("rd_critical_speed_margin", "critical_speed_margin"), ("rd_displacement_alarm", "displacement_alarm"),
                 ("rd_velocity_alarm", "velocity_alarm"), ("rd_acceleration_alarm", "acceleration_alarm")):
        d[a] = float(getattr(rc, b))
    b0 = next(iter(turb.rotor_dynamics.bearings.values())).config
    d["rd_design_load_capacity"] = float(b0.design_load_capacity)
    d["rd_bearing_clearance"] = float(b0.bearing_clearance)
    d["rd_bearing_stiffness"] = float(b0.stiffness_coefficient)
    d["rd_bearing_damping"] = float(b0.damping_coefficient)
    d["rd_friction_coefficient"] = float(b0.friction_coefficient)
    tc = turb.thermal_tracker.config
    d["tt_thermal_time_constant"] = float(tc.thermal_time_constant)
    d["tt_thermal_expansion_coeff"] = float(tc.thermal_expansion_coeff)
    d["tt_elastic_modulus"] = float(tc.elastic_modulus)
    d["tt_max_thermal_gradient"] = float(tc.max_thermal_gradient)
    d["tt_max_thermal_stress"] = float(tc.max_thermal_stress)
    pc = turb.protection_system.config
    for a, b in (("tp_overspeed_trip", "overspeed_trip"), ("tp_overspeed_delay", "overspeed_delay"),
                 ("tp_vibration_trip", "vibration_trip"), ("tp_vibration_delay", "vibration_delay"),
                 ("tp_bearing_temp_trip", "bearing_temp_trip"), ("tp_bearing_temp_delay", "bearing_temp_delay"),
                 ("tp_thrust_bearing_trip", "thrust_bearing_trip"), ("tp_low_vacuum_trip", "low_vacuum_trip"),
                 ("tp_max_thermal_stress", "max_thermal_stress")):
        d[a] = float(getattr(pc, b))
    lc = turb.bearing_lubrication_system.config
    d["tl_contamination_limit"] = float(lc.contamination_limit)
    d["tl_acidity_limit"] = float(lc.acidity_limit)
    d["tl_moisture_limit"] = float(lc.moisture_limit)
    d["tl_viscosity_change_limit"] = float(lc.viscosity_change_limit)


def condenser(cond, d):
    P = "cond."
    water_chem(cond.water_chemistry, P + "wc.", d)
    vs = cond.vacuum_system
    ej_ids = list(vs.ejectors.keys())
    assert len(ej_ids) == 2
    for i, eid in enumerate(ej_ids):
        e = vs.ejectors[eid]
        pre = f"{P}ejector[{i}]."
        d[pre + "is_operating"] = float(e.is_operating)
        for name in ("operating_hours", "current_capacity", "suction_pressure", "motive_steam_flow",
                     "motive_steam_pressure_actual", "motive_steam_temp_actual", "nozzle_fouling_factor",
                     "diffuser_fouling_factor", "nozzle_erosion_factor", "overall_performance_factor",
                     "steam_consumption_rate", "compression_ratio_actual", "entrainment_ratio"):
            d[pre + name] = float(getattr(e, name))
        for name in ("first_stage_capacity", "second_stage_capacity", "intercondenser_load"):
            d[pre + name] = float(getattr(e, name, 0.0))
        assert e.motive_steam_available
    for name in ("steam_inlet_pressure", "steam_inlet_temperature", "steam_inlet_flow", "steam_inlet_quality",
                 "cooling_water_inlet_temp", "cooling_water_outlet_temp", "cooling_water_flow", "heat_rejection_rate",
                 "overall_htc", "condensate_temperature", "condensate_flow", "thermal_performance_factor",
                 "operating_hours"):
        d[P + name] = float(getattr(cond, name))
    td = cond.tube_degradation
    for a, b in (("td_active_tube_count", "active_tube_count"), ("td_plugged_tube_count", "plugged_tube_count"),
                 ("td_average_wall_thickness", "average_wall_thickness"), ("td_tube_leak_rate", "tube_leak_rate"),
                 ("td_vibration_damage_accumulation", "vibration_damage_accumulation"),
                 ("td_corrosion_damage_accumulation", "corrosion_damage_accumulation"),
                 ("td_operating_hours", "operating_hours"), ("td_area_factor", "effective_heat_transfer_area_factor"),
                 ("td_pressure_drop_factor", "tube_side_pressure_drop_factor")):
        d[P + a] = float(getattr(td, b))
    fl = cond.fouling_model
    for a, b in (("fl_biofouling_thickness", "biofouling_thickness"), ("fl_scale_thickness", "scale_thickness"),
                 ("fl_corrosion_product_thickness", "corrosion_product_thickness"),
                 ("fl_distribution_factor", "fouling_distribution_factor"),
                 ("fl_time_since_cleaning", "time_since_cleaning"),
                 ("fl_total_fouling_resistance", "total_fouling_resistance")):
        d[P + a] = float(getattr(fl, b))
    for a, b in (("vs_condenser_pressure", "condenser_pressure"), ("vs_air_partial_pressure", "air_partial_pressure"),
                 ("vs_steam_partial_pressure", "steam_partial_pressure"),
                 ("vs_total_air_removal_rate", "total_air_removal_rate"),
                 ("vs_total_steam_consumption", "total_steam_consumption"),
                 ("vs_current_air_leakage", "current_air_leakage"), ("vs_air_mass_in_condenser", "air_mass_in_condenser"),
                 ("vs_motive_steam_pressure", "motive_steam_pressure"),
                 ("vs_motive_steam_temperature", "motive_steam_temperature"),
                 ("vs_motive_steam_available", "motive_steam_available"), ("vs_system_efficiency", "system_efficiency"),
                 ("vs_operating_hours", "operating_hours")):
        d[P + a] = float(getattr(vs, b))
    d[P + "vs_alarm_high_pressure"] = float(vs.alarms["high_pressure"])
    d[P + "vs_alarm_low_motive_pressure"] = float(vs.alarms["low_motive_pressure"])
    d[P + "vs_alarm_ejector_failure"] = float(vs.alarms["ejector_failure"])
    d[P + "vs_alarm_excessive_air_leakage"] = float(vs.alarms["excessive_air_leakage"])
    d[P + "vs_trip_high_pressure"] = float(vs.trips["high_pressure_trip"])
    cl = vs.control_logic
    assert not cl.manual_override
    d[P + "vc_lead_ejector"] = float(ej_ids.index(cl.lead_ejector_id)) if cl.lead_ejector_id is not None else -1.0
    d[P + "vc_lag_ejector"] = float(ej_ids.index(cl.lag_ejector_id)) if cl.lag_ejector_id is not None else -1.0
    d[P + "vc_rotation_timer"] = float(cl.rotation_timer)


def condenser_params(cond, d):
    c = cond.config
    ht = c.heat_transfer
    d["cd_design_heat_duty"] = float(c.design_heat_duty)
    d["cd_design_cooling_water_flow"] = float(c.design_cooling_water_flow)
    d["cd_heat_transfer_area"] = float(ht.heat_transfer_area)
    d["cd_tube_inner_diameter"] = float(ht.tube_inner_diameter)
    d["cd_tube_wall_thickness"] = float(ht.tube_wall_thickness)
    d["cd_steam_side_htc"] = float(ht.steam_side_htc)
    d["cd_water_side_htc"] = float(ht.water_side_htc)
    d["cd_tube_wall_conductivity"] = float(ht.tube_wall_conductivity)
    tc = cond.tube_degradation.config
    for a, b in (("cd_td_initial_tube_count", "initial_tube_count"), ("cd_td_tube_failure_rate", "tube_failure_rate"),
                 ("cd_td_vibration_damage_threshold", "vibration_damage_threshold"),
                 ("cd_td_wall_thickness_initial", "wall_thickness_initial"),
                 ("cd_td_wall_thickness_minimum", "wall_thickness_minimum"), ("cd_td_corrosion_rate", "corrosion_rate")):
        d[a] = float(getattr(tc, b))
    fc = cond.fouling_model.config
    for a, b in (("cd_fl_biofouling_base_rate", "biofouling_base_rate"),
                 ("cd_fl_biofouling_temp_coefficient", "biofouling_temp_coefficient"),
                 ("cd_fl_biofouling_nutrient_factor", "biofouling_nutrient_factor"),
                 ("cd_fl_scale_base_rate", "scale_base_rate"),
                 ("cd_fl_scale_hardness_coefficient", "scale_hardness_coefficient"),
                 ("cd_fl_scale_temp_coefficient", "scale_temp_coefficient"),
                 ("cd_fl_corrosion_base_rate", "corrosion_base_rate"),
                 ("cd_fl_corrosion_oxygen_coefficient", "corrosion_oxygen_coefficient"),
                 ("cd_fl_corrosion_ph_optimum", "corrosion_ph_optimum")):
        d[a] = float(getattr(fc, b))
    vc = cond.vacuum_system.config
    assert vc.control_strategy == "lead_lag", vc.control_strategy
    for a, b in (("cd_vs_base_air_leakage", "base_air_leakage"), ("cd_vs_auto_start_pressure", "auto_start_pressure"),
                 ("cd_vs_auto_stop_pressure", "auto_stop_pressure"), ("cd_vs_rotation_interval", "rotation_interval"),
                 ("cd_vs_leakage_degradation_rate", "leakage_degradation_rate"),
                 ("cd_vs_condenser_volume", "condenser_volume"), ("cd_vs_steam_pressure_drop", "steam_pressure_drop"),
                 ("cd_vs_high_pressure_alarm", "high_pressure_alarm"), ("cd_vs_high_pressure_trip", "high_pressure_trip"),
                 ("cd_vs_low_motive_pressure_alarm", "low_motive_pressure_alarm")):
        d[a] = float(getattr(vc, b))
    for i, e in enumerate(cond.vacuum_system.ejectors.values()):
        ec = e.config
        d[f"ej_two_stage[{i}]"] = float(ec.ejector_type == "two_stage")
        assert ec.ejector_type in ("two_stage", "single_stage")
        assert e.discharge_pressure == 0.101
        for a, b in (("ej_design_capacity", "design_capacity"), ("ej_design_suction_pressure", "design_suction_pressure"),
                     ("ej_motive_steam_pressure", "motive_steam_pressure"),
                     ("ej_motive_steam_temperature", "motive_steam_temperature"),
                     ("ej_base_steam_consumption", "base_steam_consumption"),
                     ("ej_steam_consumption_exponent", "steam_consumption_exponent"),
                     ("ej_pressure_effect_coefficient", "pressure_effect_coefficient"),
                     ("ej_min_suction_pressure", "min_suction_pressure"), ("ej_max_suction_pressure", "max_suction_pressure"),
                     ("ej_min_motive_pressure", "min_motive_pressure"), ("ej_intercondenser_pressure", "intercondenser_pressure"),
                     ("ej_nozzle_fouling_rate", "nozzle_fouling_rate"), ("ej_diffuser_fouling_rate", "diffuser_fouling_rate"),
                     ("ej_erosion_rate", "erosion_rate")):
            d[f"{a}[{i}]"] = float(getattr(ec, b))


PH_MODE = {"AUTO": 0.0, "MANUAL": 1.0, "FAILED": 2.0, "MAINTENANCE": 3.0}


def ph_control(ph, d):
    P = "ph."
    c = ph.controller
    st = c.state
    d[P + "control_mode"] = PH_MODE[st.control_mode.value]
    for name in ("controller_enabled", "manual_output", "measured_ph", "ph_setpoint", "ph_error", "controller_output",
                 "ammonia_dose_rate", "morpholine_dose_rate", "proportional_term", "integral_term", "derivative_term",
                 "previous_error", "integral_sum", "ammonia_tank_level", "morpholine_tank_level",
                 "ammonia_supply_available", "morpholine_supply_available", "ammonia_pump_status",
                 "morpholine_pump_status", "ph_sensor_status", "ph_low_alarm", "ph_high_alarm", "low_chemical_alarm",
                 "equipment_failure_alarm", "control_deviation_rms", "chemical_consumption_rate", "time_in_control",
                 "operating_hours"):
        d[P + name] = float(getattr(st, name))
    assert not c._output_history
    d[P + "tic_initialized"] = float(hasattr(c, "_time_in_control_sum"))
    d[P + "tic_sum"] = float(getattr(c, "_time_in_control_sum", 0.0))
    d[P + "tic_total_time"] = float(getattr(c, "_total_time", 0.0))
    d[P + "total_chemical_consumed"] = float(ph.total_chemical_consumed)
    d[P + "control_actions_count"] = float(ph.control_actions_count)
    h = c._deviation_history
    d[P + "dev_count"] = float(len(h))
    d[P + "dev_head"] = 0.0
    for i in range(100):
        d[f"{P}dev_hist[{i}]"] = float(h[i]) if i < len(h) else 0.0


def secondary(sec, d):
    P = "sec."
    for name in ("total_steam_flow", "total_heat_transfer", "electrical_power_output", "thermal_efficiency",
                 "total_feedwater_flow", "load_demand", "feedwater_temperature", "cooling_water_temperature",
                 "operating_hours", "total_system_heat_rejection"):
        d[P + name] = float(getattr(sec, name))
    d[P + "has_previous_feedwater_temp"] = float(hasattr(sec, "_previous_feedwater_temp"))
    d[P + "previous_feedwater_temp"] = float(getattr(sec, "_previous_feedwater_temp", 0.0))
    prev = getattr(sec, "_previous_sg_conditions", None)
    d[P + "has_previous_sg_conditions"] = float(prev is not None)
    for i in range(3):
        d[f"{P}prev_sg_levels[{i}]"] = float(prev["levels"][i]) if prev else 0.0
        d[f"{P}prev_sg_pressures[{i}]"] = float(prev["pressures"][i]) if prev else 0.0
        d[f"{P}prev_sg_steam_flows[{i}]"] = float(prev["steam_flows"][i]) if prev else 0.0
        d[f"{P}prev_sg_steam_qualities[{i}]"] = float(prev["steam_qualities"][i]) if prev else 0.0
    last = getattr(sec, "_nps_last_result", None)
    d[P + "sg_avg_pressure"] = float(last["sg_avg_pressure"]) if last else 0.0
    d[P + "sg_avg_temperature"] = float(last["sg_avg_temperature"]) if last else 0.0
    d[P + "condenser_pressure"] = float(last["condenser_pressure"]) if last else 0.0
    d[P + "heat_rate_kj_kwh"] = float(last["heat_rate_kj_kwh"]) if last else 0.0
    tnet = float(last["turbine_electrical_power_net"]) if last else 0.0
    # power_reduction_factor is a local of update_system (systems/secondary/__init__.py:864-921): electrical = net * factor.
    # With a tripped turbine the net output is 0 and the quotient says nothing; the factor then follows from the same
    # three conditions on the reference's own outputs (the energy-conservation branch cannot fire: thermal_power_mw IS
    # the primary thermal power there).
    if tnet:
        factor = float(last["electrical_power_mw"]) / tnet
    elif not last:
        factor = 1.0
    else:
        factor = 1.0
        if float(last["feedwater_total_flow"]) < 300.0:
            factor = 0.0
        else:
            if float(last["total_steam_flow"]) < 150.0:
                factor *= 0.1
            if float(last["sg_avg_pressure"]) < 0.5:
                factor *= 0.1
    d[P + "power_reduction_factor"] = factor


def report(sec, d):
    """ReportState: what the reference's get_state_dict() methods / HeatFlowTracker report at this moment."""
    P = "rep."
    sgs = sec.steam_generator_system
    for i, sg in enumerate(sgs.steam_generators):
        sd = sg.get_state_dict()
        for name in ("primary_flow_restriction_factor", "secondary_flow_restriction_factor",
                     "max_primary_flow_capacity", "max_steam_flow_capacity", "max_feedwater_flow_capacity",
                     "fouling_energy_penalty_mw", "total_pump_power_mw"):
            d[f"{P}sg_{name}[{i}]"] = float(sd[name])
    sd = sgs.get_state_dict()
    d[P + "sgs_total_fouling_impact"] = float(sd["system_total_fouling_impact"])
    d[P + "sgs_fouling_maintenance_needed"] = float(sd["system_fouling_maintenance_needed"])
    fw = sec.feedwater_system
    for k, pump in enumerate(fw.pump_system.pumps.values()):
        L = pump.lubrication_system
        d[f"{P}fwp_efficiency_factor[{k}]"] = float(L.pump_efficiency_factor)
        d[f"{P}fwp_flow_factor[{k}]"] = float(L.pump_flow_factor)
    fd = fw.get_state_dict()
    d[P + "fw_avg_sg_level"] = float(fd["feedwater_avg_sg_level"])
    d[P + "fw_avg_sg_pressure"] = float(fd["feedwater_avg_sg_pressure"])
    d[P + "fw_total_steam_flow"] = float(fd["feedwater_total_steam_flow"])
    d[P + "fw_avg_steam_quality"] = float(fd["feedwater_avg_steam_quality"])
    dd = fw.diagnostics.get_state_dict()
    d[P + "fw_diag_maintenance_urgency"] = float(dd["diagnostics_maintenance_urgency"])
    d[P + "fw_diag_total_wear"] = float(dd["diagnostics_total_wear"])
    d[P + "fw_prot_active_alarms_count"] = float(len(fw.protection_system.active_alarms))
    hs = sec.heat_flow_tracker.heat_flow_state
    d[P + "hf_steam_enthalpy_flow"] = float(hs.steam_enthalpy_flow)
    d[P + "hf_turbine_work_output"] = float(hs.turbine_work_output)
    d[P + "hf_condenser_heat_rejection"] = float(hs.condenser_heat_rejection)
    d[P + "hf_net_electrical_output"] = float(hs.net_electrical_output)
    d[P + "hf_overall_efficiency"] = float(hs.overall_thermal_efficiency)
    d[P + "hf_energy_balance_error"] = float(hs.energy_balance_error)
    d[P + "hf_energy_balance_percent"] = float(hs.energy_balance_percent_error)


def extract(sim, d):
    if not (sim.enable_secondary and sim.secondary_physics is not None):
        return
    sec = sim.secondary_physics
    report(sec, d)
    water_chem(sec.water_chemistry, "wc_main.", d)
    feedwater(sec.feedwater_system, d)
    steam_generators(sec.steam_generator_system, d)
    turbine(sec.turbine, d)
    condenser(sec.condenser, d)
    ph_control(sec.ph_control_system, d)
    secondary(sec, d)


def extract_params(sim, d):
    if not (sim.enable_secondary and sim.secondary_physics is not None):
        return
    sec = sim.secondary_physics
    feedwater_params(sec.feedwater_system, d)
    steam_generator_params(sec.steam_generator_system, d)
    turbine_params(sec.turbine, d)
    condenser_params(sec.condenser, d)
