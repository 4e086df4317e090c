// Feedwater system: three-element level control, four pumps (3 running + 1 spare) with their
// lubrication systems, performance diagnostics and the protection system.
// Restates EnhancedFeedwaterPhysics.update_state
// (reference: nuclear_simulator/systems/secondary/feedwater/physics.py:662-863) and its callees.
#pragma once
#include "hd.h"
#include "state.h"
#include "prefetch.h"
#include "lubrication.h"
#include "water_chemistry.h"

namespace nps {

enum PumpStatus : int { PUMP_RUNNING = 0, PUMP_STOPPED = 1, PUMP_STARTING = 2, PUMP_STOPPING = 3, PUMP_TRIPPED = 4 };

enum PumpTrip : int {
    TRIP_NONE = 0, TRIP_LOW_FLOW = 1, TRIP_NPSH = 2, TRIP_LOW_SUCTION = 3, TRIP_HIGH_DISCHARGE = 4,
    TRIP_SG_HIGH_LEVEL = 5, TRIP_SEVERE_CAVITATION = 6, TRIP_CAVITATION_DAMAGE = 7, TRIP_CRITICAL_NPSH = 8,
    TRIP_LUB_VERY_LOW_OIL = 9, TRIP_LUB_LOW_OIL = 10, TRIP_LUB_OVERFILL = 11, TRIP_LUB_WEAR_BASE = 12,
    TRIP_LUB_SEAL_LEAKAGE = 18, TRIP_LUB_COMBINED_WEAR = 19, TRIP_LUB_PERFORMANCE = 20
};

// Component order: impeller, motor_bearings, pump_bearings, thrust_bearing, mechanical_seals,
// coupling_system (feedwater/pump_lubrication.py:105-196)
enum { FWL_IMPELLER = 0, FWL_MOTOR_BRG = 1, FWL_PUMP_BRG = 2, FWL_THRUST_BRG = 3, FWL_SEALS = 4, FWL_COUPLING = 5, FWL_NCOMP = 6 };

NPS_HD LubComponent fw_lub_component(int i) {
    //                 base    load speed contam wearperf lubperf alarm trip
    switch (i) {
        case FWL_IMPELLER:   return {0.002, 1.8, 2.0, 1.5, 0.025, 0.0, 10.0, 25.0};
        case FWL_MOTOR_BRG:  return {0.004, 1.8, 2.0, 2.5, 0.015, 0.05, 20.0, 60.0};
        case FWL_PUMP_BRG:   return {0.006, 2.2, 1.8, 3.0, 0.02, 0.1, 15.0, 50.0};
        case FWL_THRUST_BRG: return {0.008, 2.4, 1.6, 3.5, 0.025, 0.15, 12.0, 40.0};
        case FWL_SEALS:      return {0.01, 2.0, 1.4, 4.0, 0.03, 0.2, 15.0, 50.0};
        default:             return {0.0003, 1.3, 1.0, 1.5, 0.01, 0.05, 15.0, 35.0};
    }
}

// pump_efficiency_factor / pump_flow_factor / pump_head_factor properties: pump_lubrication.py:224-237
NPS_HD double fwp_efficiency_factor(const FWPumpState& u) { return py_max(0.5, 1.0 - (u.pump_efficiency_degradation / 100.0)); }
NPS_HD double fwp_flow_factor(const FWPumpState& u) { return py_max(0.5, 1.0 - (u.pump_flow_degradation / 100.0)); }
NPS_HD double fwp_head_factor(const FWPumpState& u) { return py_max(0.7, 1.0 - (u.pump_head_degradation / 100.0)); }

struct PumpSysCond { double feedwater_temperature, suction_pressure, discharge_pressure; double sg_levels[3]; };

// _calculate_dynamic_npsh_required: feedwater/pump_system.py:362-420
NPS_HD double fwp_dynamic_npsh_required(const FWPumpState& u, const PlantParams& p) {
    const double base_npsh = 12.0;
    double impeller_wear = u.lub.component_wear[FWL_IMPELLER];
    double impeller_pen = impeller_wear * 0.1;
    double cav_pen = u.cavitation_damage * 0.2;
    double speed_pen = py_max(0.0, (u.speed_percent - 100.0) * 0.02);
    double flow_ratio = u.flow_rate / p.fwp_rated_flow;
    double flow_pen = py_max(0.0, (flow_ratio - 1.0) * 1.5);
    double max_brg = py_max3(u.lub.component_wear[FWL_MOTOR_BRG], u.lub.component_wear[FWL_PUMP_BRG],
                             u.lub.component_wear[FWL_THRUST_BRG]);
    double brg_pen = max_brg * 0.05;
    double coupling_pen = (impeller_wear * max_brg / 10000.0) * 0.3;
    double total = (base_npsh + impeller_pen + cav_pen + speed_pen + flow_pen + brg_pen + coupling_pen);
    return py_max(base_npsh, total);
}

// FeedwaterPumpLubricationSystem.calculate_component_wear: pump_lubrication.py:275-396
// (the impeller receives no conditions from the wrapper, so it sees the defaults).
NPS_HD double fwp_component_wear_rate(const FWPumpState& u, int comp, double load_factor, double speed_factor,
                                      double temperature, double cav, double electrical_load_factor,
                                      double head_factor, double pressure_factor) {
    const LubComponent c = fw_lub_component(comp);
    const double impeller_wear = u.lub.component_wear[FWL_IMPELLER];
    const double max_brg = py_max3(u.lub.component_wear[FWL_MOTOR_BRG], u.lub.component_wear[FWL_PUMP_BRG],
                                   u.lub.component_wear[FWL_THRUST_BRG]);
    double wear_rate;
    if (comp == FWL_IMPELLER) {
        double cav_factor = 1.0 + cav * 3.0;
        double temp_factor = py_max(1.0, (temperature - 80.0) / 40.0);
        double brg_cpl = 1.0 + (max_brg / 100.0) * 0.3;
        wear_rate = (c.base_wear_rate * py_pow(load_factor, c.load_wear_exponent) * py_pow(speed_factor, c.speed_wear_exponent) *
                     cav_factor * temp_factor * brg_cpl);
    } else if (comp == FWL_MOTOR_BRG) {
        double temp_factor = py_max(1.0, (temperature - 60.0) / 25.0);
        double imp_cpl = 1.0 + (impeller_wear / 100.0) * 0.2;
        wear_rate = (c.base_wear_rate * py_pow(electrical_load_factor, c.load_wear_exponent) *
                     py_pow(speed_factor, c.speed_wear_exponent) * temp_factor * imp_cpl);
    } else if (comp == FWL_PUMP_BRG) {
        double cav_factor = 1.0 + cav * 2.0;
        double temp_factor = py_max(1.0, (temperature - 50.0) / 30.0);
        double imp_cpl = 1.0 + (impeller_wear / 100.0) * 0.4;
        wear_rate = (c.base_wear_rate * py_pow(load_factor, c.load_wear_exponent) * py_pow(speed_factor, c.speed_wear_exponent) *
                     cav_factor * temp_factor * imp_cpl);
    } else if (comp == FWL_THRUST_BRG) {
        double axial = head_factor * load_factor;
        double imp_cpl = 1.0 + (impeller_wear / 100.0) * 0.25;
        wear_rate = (c.base_wear_rate * py_pow(axial, c.load_wear_exponent) * py_pow(speed_factor, c.speed_wear_exponent) * imp_cpl);
    } else if (comp == FWL_SEALS) {
        double cav_seal = 1.0 + cav * 5.0;
        double imp_cpl = 1.0 + (impeller_wear / 100.0) * 0.15;
        double brg_cpl = 1.0 + (max_brg / 100.0) * 0.2;
        wear_rate = (c.base_wear_rate * py_pow(pressure_factor, c.load_wear_exponent) * 1.0 * cav_seal * imp_cpl * brg_cpl);
    } else {
        double brg_cpl = 1.0 + (max_brg / 100.0) * 0.3;
        wear_rate = (c.base_wear_rate * 1.0 * 1.0 * py_pow(load_factor, c.load_wear_exponent) * brg_cpl);
    }
    wear_rate *= 1.0;  // chemistry_wear_factor default
    return wear_rate;
}

// _calculate_pump_performance_factors: pump_lubrication.py:1412-1478
NPS_HD void fwp_performance_factors(FWPumpState& u, double cavitation_damage) {
    double mb = u.lub.component_wear[FWL_MOTOR_BRG], pb = u.lub.component_wear[FWL_PUMP_BRG], tb = u.lub.component_wear[FWL_THRUST_BRG];
    double brg_loss = ((mb / 100.0) * 0.01 + (pb / 100.0) * 0.015 + (tb / 100.0) * 0.02);
    double seal_loss = (u.lub.component_wear[FWL_SEALS] / 100.0) * 0.01;
    double lub_loss = (1.0 - u.lub.lubrication_effectiveness) * 0.02;
    double cav_eff_loss = py_min(0.3, cavitation_damage * 0.01);
    double cav_flow_loss = cav_eff_loss * 0.5;
    double imp_flow_loss = cavitation_damage * 0.02;
    double imp_eff_loss = cavitation_damage * 0.015;
    double total_eff = (brg_loss + seal_loss + lub_loss + cav_eff_loss + imp_eff_loss);
    double total_flow = (cav_flow_loss + imp_flow_loss + brg_loss * 0.3);
    double total_head = (imp_flow_loss * 0.8 + cav_eff_loss * 0.4);
    u.pump_efficiency_degradation = py_min(50.0, total_eff * 100.0);
    u.pump_flow_degradation = py_min(50.0, total_flow * 100.0);
    u.pump_head_degradation = py_min(30.0, total_head * 100.0);
    u.npsh_margin_degradation = cavitation_damage * 0.5;
    double total_brg = mb + pb + tb;
    u.vibration_increase = total_brg * 0.1 + cavitation_damage * 0.05;
}

// Lubrication part of update_with_lubrication: pump_lubrication.py:1659-1830 (uses the pump state
// left by the previous step), followed by update_pump_lubrication_effects (:570-623).
NPS_HD void fwp_update_lubrication(FWPumpState& u, const PlantParams& p, const PumpSysCond& sc, double dt) {
    NPS_TOUCH(u.flow_rate); NPS_TOUCH(u.speed_percent); NPS_TOUCH(u.power_consumption); NPS_TOUCH(u.cavitation_intensity);
    NPS_TOUCH(u.differential_pressure); NPS_TOUCH(u.lub.lubrication_effectiveness); NPS_TOUCH(u.lub.oil_level);
    for (int c = 0; c < FWL_NCOMP; ++c) NPS_TOUCH(u.lub.component_wear[c]);
    NPS_TOUCH(u.cavitation_damage); NPS_TOUCH(u.lub.oil_contamination_level); NPS_TOUCH(u.lub.oil_temperature);
    double load_factor = (p.fwp_rated_flow > 0) ? u.flow_rate / p.fwp_rated_flow : 0.0;
    double speed_factor = u.speed_percent / 100.0;
    double elf = (p.fwp_rated_power > 0) ? u.power_consumption / p.fwp_rated_power : 0.0;
    double cav = u.cavitation_intensity;
    double pressure_factor = u.differential_pressure / 7.5;

    double base_temp = 40.0 + load_factor * 10.0;
    double motor_heat = elf * 2.0;
    double fw_heat = (sc.feedwater_temperature - 200.0) * 0.01;
    double pressure_ratio = (sc.suction_pressure > 0) ? sc.discharge_pressure / sc.suction_pressure : 16.0;
    double pressure_heat = py_max(0.0, (pressure_ratio - 12.0) * 0.5);
    double cav_heat = cav * 3.0;
    double oil_temp = base_temp + motor_heat + fw_heat + pressure_heat + cav_heat;
    oil_temp = py_max(35.0, py_min(75.0, oil_temp));

    double base_contam = load_factor * 0.002;
    double mbw = u.lub.component_wear[FWL_MOTOR_BRG], pbw = u.lub.component_wear[FWL_PUMP_BRG];
    double tbw = u.lub.component_wear[FWL_THRUST_BRG], sw = u.lub.component_wear[FWL_SEALS];
    double brg_contam = (mbw + pbw + tbw) * 0.0025;
    double seal_contam = sw * 0.004;
    double cav_contam = cav * 0.01;
    double temp_contam = (oil_temp > 70.0) ? (oil_temp - 70.0) * 0.0025 : 0.0;
    double lqf = py_max(0.3, u.lub.lubrication_effectiveness);
    double scaling = 2.0 - (lqf * 0.7);
    double total_contam = (base_contam + brg_contam + seal_contam + cav_contam + temp_contam) * scaling;
    total_contam = py_min(0.5, py_max(0.0005, total_contam));

    const LubLimits lim = {15.0, 1.6, 0.08, 10.0};
    const double dth = dt / 60.0;
    lub_update_oil_quality(u.lub, FWL_NCOMP, lim, oil_temp, total_contam, 0.0001, dth);

    // update_component_wear (dict order; each rate sees wear already updated for earlier components)
    for (int c = 0; c < FWL_NCOMP; ++c) {
        double rate;
        switch (c) {
            case FWL_IMPELLER:   rate = fwp_component_wear_rate(u, c, 1.0, 1.0, 55.0, 0.0, 1.0, 1.0, 1.0); break;
            case FWL_MOTOR_BRG:  rate = fwp_component_wear_rate(u, c, elf, speed_factor, 60.0 + elf * 25.0, 0.0, elf, 1.0, 1.0); break;
            case FWL_PUMP_BRG:   rate = fwp_component_wear_rate(u, c, load_factor, speed_factor, 50.0 + load_factor * 30.0, cav, 1.0, 1.0, 1.0); break;
            case FWL_THRUST_BRG: rate = fwp_component_wear_rate(u, c, load_factor, speed_factor, 45.0 + load_factor * 30.0, 0.0, 1.0, 1.0, 1.0); break;
            case FWL_SEALS:      rate = fwp_component_wear_rate(u, c, load_factor, speed_factor, 40.0 + load_factor * 30.0, cav, 1.0, 1.0, pressure_factor); break;
            default:             rate = fwp_component_wear_rate(u, c, load_factor, speed_factor, 50.0 + load_factor * 20.0, 0.0, 1.0, 1.0, 1.0); break;
        }
        lub_apply_component_wear(u.lub, c, fw_lub_component(c), rate, dth);
    }
    lub_update_health(u.lub, FWL_NCOMP);

    // update_pump_lubrication_effects(pump_conditions, dt [minutes])
    u.pump_load_factor = load_factor;
    u.cavitation_lubrication_effect = py_max(0.3, 1.0 - cav * 0.5);
    double total_leak = 0.001 + u.lub.component_wear[FWL_SEALS] * 0.2 + cav * 0.1;
    u.seal_leakage_rate = py_min(total_leak, 0.05);
    if (u.seal_leakage_rate > 0) {
        double lost_l = u.seal_leakage_rate * dt;
        double loss_pct = (lost_l / 150.0) * 100.0;
        u.lub.oil_level = py_max(0.0, u.lub.oil_level - loss_pct * 0.5);
    }
    u.lub.oil_level = py_min(100.0, py_max(0.0, u.lub.oil_level));
    fwp_performance_factors(u, 0.0);  // the wrapper passes no 'cavitation_damage'
}

NPS_HD void fwp_trip(FWPumpState& u, int reason) {
    u.status = PUMP_TRIPPED; u.trip_active = 1.0; u.trip_reason = (double)reason; u.available = 0.0;
}

// FeedwaterPump.set_flow_demand: feedwater/pump_system.py:422-447
NPS_HD void fwp_set_flow_demand(FWPumpState& u, const PlantParams& p, double flow_demand) {
    u.flow_demand = np_clip(flow_demand, 0.0, p.fwp_rated_flow * 1.2);
    if (flow_demand > 0) {
        double eff_cap = p.fwp_rated_flow * fwp_flow_factor(u);
        double sp = (eff_cap > 0) ? sqrt(flow_demand / eff_cap) * 100.0 : 100.0;
        u.speed_setpoint = np_clip(sp, 0.0, 100.0);
    } else {
        u.speed_setpoint = 0.0;
    }
}

// _vapor_pressure: pump_system.py:239-248
NPS_HD double fwp_vapor_pressure(double t) {
    return (t <= 100) ? 0.001 + (t / 100.0) * 0.01 : 0.05 + (t - 100) * 0.001;
}

// _simulate_sensors: pump_system.py:637-744
NPS_HD void fwp_simulate_sensors(FWPumpState& u, const PlantParams& p, const PumpSysCond& sc) {
    if (!is_true(u.ic_applied)) {
        u.suction_pressure = sc.suction_pressure;
        u.discharge_pressure = sc.discharge_pressure;
    }
    u.differential_pressure = u.discharge_pressure - u.suction_pressure;
    if (!is_true(u.ic_applied)) {
        double npsh = (u.suction_pressure - fwp_vapor_pressure(sc.feedwater_temperature)) * 100.0;
        u.npsh_available = py_max(0.0, npsh);
    }
    double lf = (p.fwp_rated_flow > 0) ? u.flow_rate / p.fwp_rated_flow : 0.0;
    u.motor_current = 200.0 + 100.0 * lf;
    u.motor_voltage = 6.6;
    u.motor_temperature = 60.0 + 20.0 * lf;
    double base_vib = 1.0 + 0.05 * u.speed_percent;
    u.vibration_level = base_vib + u.vibration_increase + u.cavitation_intensity * 2.0;
    u.vibration_level = py_max(0.5, py_min(15.0, u.vibration_level));
}

// FeedwaterPump.update_pump (pump_system.py:449-464) over BasePump.update_pump
// (primary/coolant/pump_models.py:104-146); dt in minutes as passed by the feedwater system.
NPS_HD void fwp_update(FWPumpState& u, const PlantParams& p, const PumpSysCond& sc, double dt) {
    const double max_speed = 110.0, speed_ramp_rate = 15.0, startup_time = 20.0, coastdown_time = 60.0;
    NPS_TOUCH(u.speed_setpoint); NPS_TOUCH(u.flow_demand); NPS_TOUCH(u.ic_applied); NPS_TOUCH(u.suction_pressure);
    NPS_TOUCH(u.npsh_available); NPS_TOUCH(u.discharge_pressure); NPS_TOUCH(u.pump_flow_degradation);
    int status = (int)u.status;
    // _update_pump_dynamics
    if (status == PUMP_RUNNING) {
        double err = u.speed_setpoint - u.speed_percent;
        double max_change = speed_ramp_rate * dt;
        if (fabs(err) <= max_change) u.speed_percent = u.speed_setpoint;
        else u.speed_percent += max_change * np_sign(err);
    } else if (status == PUMP_STARTING) {
        double acc = 100.0 / startup_time;
        u.speed_percent += acc * dt;
        if (u.speed_percent >= u.speed_setpoint * 0.95) { status = PUMP_RUNNING; u.speed_percent = u.speed_setpoint; }
    } else if (status == PUMP_STOPPING) {
        double dec = 100.0 / coastdown_time;
        u.speed_percent -= dec * dt;
        if (u.speed_percent <= 5.0) { u.speed_percent = 0.0; status = PUMP_STOPPED; }
    }
    u.status = (double)status;
    u.speed_percent = np_clip(u.speed_percent, 0.0, max_speed);

    // _calculate_flow_rate: pump_system.py:172-216
    if (status == PUMP_RUNNING || status == PUMP_STARTING) {
        double speed_ratio = u.speed_percent / 100.0;
        if (u.flow_demand > 0) {
            if (speed_ratio > 0.8) u.flow_rate = u.flow_demand;
            else u.flow_rate = py_min(u.flow_demand, p.fwp_rated_flow * speed_ratio);
        } else {
            u.flow_rate = p.fwp_rated_flow * speed_ratio;
        }
        {   // _apply_system_effects: pump_system.py:218-237
            double temp_factor = 1.0 - (sc.feedwater_temperature - 227.0) * 0.0002;
            if (!is_true(u.ic_applied)) {
                u.suction_pressure = sc.suction_pressure;
                double npsh = (u.suction_pressure - fwp_vapor_pressure(sc.feedwater_temperature)) * 100.0;
                u.npsh_available = py_max(0.0, npsh);
            }
            u.flow_rate *= temp_factor;
        }
        u.flow_rate *= fwp_flow_factor(u);
        if (status == PUMP_RUNNING && u.speed_percent < 20.0) u.flow_rate = py_max(u.flow_rate, p.fwp_rated_flow * 0.05);
        u.flow_rate = py_min(u.flow_rate, p.fwp_rated_flow * 1.2);
    } else {
        u.flow_rate = 0.0;
    }
    // _calculate_power_consumption: pump_system.py:250-275
    if (status == PUMP_RUNNING || status == PUMP_STARTING) {
        double speed_ratio = u.speed_percent / 100.0;
        double flow_ratio = u.flow_rate / p.fwp_rated_flow;
        double head_ratio = py_pow(speed_ratio, 2.0);
        double base_power = p.fwp_rated_power * (flow_ratio * head_ratio);
        u.power_consumption = base_power / fwp_efficiency_factor(u);
        if (status == PUMP_STARTING) u.power_consumption = py_max(u.power_consumption, p.fwp_rated_power * 0.2);
    } else {
        u.power_consumption = 0.0;
    }
    fwp_simulate_sensors(u, p, sc);

    // _check_protection_systems: pump_models.py:243-259 then pump_system.py:277-359
    do {
        if (status == PUMP_STOPPED) { u.trip_active = 0.0; u.trip_reason = 0.0; }
        else if (status == PUMP_RUNNING && u.flow_rate < 25.0) { fwp_trip(u, TRIP_LOW_FLOW); }
        if (is_true(u.trip_active)) break;
        if (status == PUMP_STARTING) break;
        double npsh_req = fwp_dynamic_npsh_required(u, p);
        if (status == PUMP_RUNNING && u.npsh_available < npsh_req) { fwp_trip(u, TRIP_NPSH); break; }
        if (status == PUMP_RUNNING && u.suction_pressure < 0.2) { fwp_trip(u, TRIP_LOW_SUCTION); break; }
        if (sc.discharge_pressure > 10.0) { fwp_trip(u, TRIP_HIGH_DISCHARGE); break; }
        double max_level = py_max3(sc.sg_levels[0], sc.sg_levels[1], sc.sg_levels[2]);
        if (max_level > 16.0) { fwp_trip(u, TRIP_SG_HIGH_LEVEL); break; }
        if (u.cavitation_intensity > 0.7) { fwp_trip(u, TRIP_SEVERE_CAVITATION); break; }
        if (u.cavitation_damage > 10.0) { fwp_trip(u, TRIP_CAVITATION_DAMAGE); break; }
        if (u.npsh_available < 8.0 * 0.5) { fwp_trip(u, TRIP_CRITICAL_NPSH); break; }
        // lubrication_system.check_protection_trips: pump_lubrication.py:1536-1580
        if (u.lub.oil_level < 5.0) { fwp_trip(u, TRIP_LUB_VERY_LOW_OIL); break; }
        if (u.lub.oil_level < 10.0) { fwp_trip(u, TRIP_LUB_LOW_OIL); break; }
        if (u.lub.oil_level > 105.0) { fwp_trip(u, TRIP_LUB_OVERFILL); break; }
        bool tripped = false;
        double total_wear = 0.0;
        for (int c = 0; c < FWL_NCOMP; ++c) {
            if (!tripped && u.lub.component_wear[c] > fw_lub_component(c).wear_trip_threshold) {
                fwp_trip(u, TRIP_LUB_WEAR_BASE + c); tripped = true;
            }
            total_wear += u.lub.component_wear[c];
        }
        if (tripped) break;
        if (u.seal_leakage_rate > 10.0) { fwp_trip(u, TRIP_LUB_SEAL_LEAKAGE); break; }
        if (total_wear > 40.0) { fwp_trip(u, TRIP_LUB_COMBINED_WEAR); break; }
        if ((1.0 - fwp_efficiency_factor(u)) * 100.0 > 25.0) { fwp_trip(u, TRIP_LUB_PERFORMANCE); break; }
    } while (0);
    status = (int)u.status;

    fwp_simulate_sensors(u, p, sc);  // second call in FeedwaterPump.update_pump

    // _simulate_cavitation: pump_system.py:556-601
    if (!(status == PUMP_RUNNING || status == PUMP_STARTING)) {
        u.cavitation_intensity = 0.0; u.cavitation_time = 0.0; u.cavitation_noise_level = 0.0;
    } else {
        double thr = fwp_dynamic_npsh_required(u, p) + 2.0;
        if (u.npsh_available < thr) {
            double deficit = thr - u.npsh_available;
            double severity = py_min(1.0, deficit / thr);
            double fr = u.flow_rate / p.fwp_rated_flow;
            double flow_factor = py_pow(fr, 2.0);
            u.cavitation_intensity = severity * flow_factor;
            u.cavitation_time += dt * 60.0;
            u.cavitation_noise_level = 20.0 + u.cavitation_intensity * 30.0;
            u.vibration_level += u.cavitation_intensity * 2.0;
        } else {
            u.cavitation_intensity = 0.0;
            u.cavitation_noise_level = 0.0;
            u.cavitation_time = py_max(0.0, u.cavitation_time - dt * 6.0);
        }
        double damage_rate = py_pow(u.cavitation_intensity, 2.0) * dt / 60.0;
        u.cavitation_damage += damage_rate;
    }
    // _simulate_mechanical_wear: pump_system.py:603-617
    if (status == PUMP_RUNNING && u.cavitation_intensity > 0.1) {
        double damage_rate = py_pow(u.cavitation_intensity, 2.0) * dt / 60.0;
        u.cavitation_damage += damage_rate;
    }
}

struct FeedwaterResult {
    double total_flow_rate, total_power_consumption, num_running_pumps, system_availability;
    double sg_flow[3];
};

// EnhancedFeedwaterPhysics.update_state: feedwater/physics.py:662-863
NPS_HD void feedwater_update(FeedwaterState& fw, WaterChemState& wc, const PlantParams& p,
                             const double* sg_levels, const double* sg_steam_flows, const double* sg_steam_qualities,
                             double manual_total_flow, double fw_temperature, double suction_pressure,
                             double discharge_pressure, double dt, FeedwaterResult& out,
                             const SGState* prefetch_next = nullptr, ReportState* rep = nullptr) {
    const int nsg = 3;
    const MakeupWater mk = {7.2, 100.0, 300.0, 30.0, 8.0};
    NPS_TOUCH(fw.lc_quality_integral_error); NPS_TOUCH(fw.n_running_prev); NPS_TOUCH(fw.cav_n_events); NPS_TOUCH(fw.cav_accumulated_damage); NPS_TOUCH(fw.operating_hours); NPS_TOUCH(fw.cav_time_in_cavitation);
    for (int i = 0; i < 3; ++i) { NPS_TOUCH(fw.lc_level_integral_errors[i]); NPS_TOUCH(fw.lc_previous_level_errors[i]); }
    wc_update(wc, true, mk, 0.02, dt);

    // ThreeElementControl.calculate_flow_demands: feedwater/level_control.py:157-363
    double individual[3];
    double total_demand = 0.0;
    const double design_flow_per_sg = p.fw_design_total_flow / p.fw_num_sg;
    if (is_true(p.fw_auto_level_control)) {
        double qcorr[3];
        for (int i = 0; i < nsg; ++i) {   // SteamQualityCompensator: level_control.py:58-106
            double qe = 0.99 - sg_steam_qualities[i];
            if (fabs(qe) < 0.005) qe = 0.0;
            double prop = qe * p.fw_lc_quality_gain * sg_steam_flows[i];
            fw.lc_quality_integral_error += qe * dt;
            double integ = fw.lc_quality_integral_error * 0.1 * sg_steam_flows[i];
            qcorr[i] = np_clip(prop + integ, -50.0, 50.0);
        }
        double abs_err_sum = 0.0;
        for (int i = 0; i < nsg; ++i) {
            double level = sg_levels[i], steam_flow = sg_steam_flows[i];
            double level_error = p.fw_design_sg_level - level;
            fw.lc_level_errors[i] = level_error;
            double prop = 1.0 * level_error * 0.2;
            fw.lc_level_integral_errors[i] += level_error * dt;
            double integ = fw.lc_level_integral_errors[i] * 0.0005 * 0.5;
            double rate = (level_error - fw.lc_previous_level_errors[i]) / dt;
            double deriv = rate * 0.0002 * 0.5;
            double absolute_minimum = design_flow_per_sg * 0.05;
            double feedforward = (steam_flow < absolute_minimum) ? absolute_minimum : steam_flow;
            double level_corr = prop + integ + deriv;
            double max_corr = feedforward * 0.01;
            level_corr = np_clip(level_corr, -max_corr, max_corr);
            double flow_error = feedforward - design_flow_per_sg;  // previous_feedwater_flows never updated
            double flow_fb = flow_error * p.fw_lc_feedwater_flow_weight;
            double level_contrib = level_corr * p.fw_lc_level_control_weight;
            double flow_contrib = flow_fb * p.fw_lc_feedwater_flow_weight;
            double td = feedforward + level_contrib + flow_contrib + qcorr[i];
            td = np_clip(td, design_flow_per_sg * 0.05, design_flow_per_sg * 2.0);
            individual[i] = td;
            fw.lc_previous_level_errors[i] = level_error;
            abs_err_sum += fabs(level_error);
        }
        total_demand = 0.0 + individual[0]; total_demand += individual[1]; total_demand += individual[2];
        double avg_err = abs_err_sum / 3.0;
        fw.lc_control_performance = py_max(0.0, 1.0 - avg_err / 2.0);
        fw.total_flow_demand = total_demand;   // ThreeElementControl.flow_demand_history[-1]: appended by the controller only
    } else {
        total_demand = manual_total_flow;      // manual mode (feedwater/physics.py:735-741): the level controller is not called
    }

    PumpSysCond sc;
    sc.feedwater_temperature = fw_temperature;
    sc.suction_pressure = suction_pressure;
    sc.discharge_pressure = discharge_pressure;
    for (int i = 0; i < 3; ++i) sc.sg_levels[i] = sg_levels[i];

    // FeedwaterPumpSystem.update_system: feedwater/pump_system.py:1235-1329
    const int n_prev = (int)fw.n_running_prev;
    double flow_per_pump = (n_prev > 0) ? total_demand / n_prev : 0.0;
    double total_flow = 0.0, total_power = 0.0;
    int n_running = 0;
    double speed_sum = 0.0, perf_sum = 0.0;
    NPS_UNIT_LOOP
    for (int k = 0; k < 4; ++k) {
        FWPumpState& u = fw.pump[k];
        NPS_PREFETCH_SELF(u);
        if (k < 3) NPS_PREFETCH_FAR(fw.pump[k + 1]);
        else if (prefetch_next) NPS_PREFETCH_FAR(*prefetch_next);   // what runs after the feedwater system
        if ((int)u.status == PUMP_RUNNING && n_prev > 0) {
            if (n_running < n_prev) {
                if (!(flow_per_pump < p.fwp_rated_flow * 0.2)) fwp_set_flow_demand(u, p, flow_per_pump);
            }
        }
        fwp_update_lubrication(u, p, sc, dt);
        fwp_update(u, p, sc, dt);
        if ((int)u.status == PUMP_RUNNING) {
            total_flow += u.flow_rate;
            total_power += u.power_consumption;
            speed_sum += u.speed_percent;
            perf_sum += fwp_flow_factor(u) * fwp_efficiency_factor(u);
            n_running += 1;
        }
    }
    fw.n_running_prev = (double)n_running;
    fw.pump_system_available = as_flag(n_running >= 3);
    double avg_perf = (n_running > 0) ? perf_sum / n_running : 0.0;
    (void)speed_sum;

    // PerformanceDiagnostics.update_diagnostics: feedwater/performance_monitoring.py:423-542
    double tot_risk = 0.0, tot_wear = 0.0, tot_vib = 0.0;
    double wear_for_protection = 0.0;
    NPS_UNIT_LOOP
    for (int k = 0; k < 4; ++k) {
        FWPumpState& u = fw.pump[k];
        double npsh_req = fwp_dynamic_npsh_required(u, p);
        double thr = npsh_req + 2.0;
        if (u.npsh_available < thr) {   // CavitationModel.update_cavitation_monitoring :113-202
            double deficit = thr - u.npsh_available;
            double severity = py_min(1.0, deficit / thr);
            double ff = py_pow(u.flow_rate / 555.0, 2.0);
            double sf = py_pow(u.speed_percent / 100.0, 1.5);
            fw.cav_current_intensity = severity * ff * sf;
            fw.cav_time_in_cavitation += dt;
            if (fw.cav_current_intensity > 0.1) fw.cav_n_events = py_min(fw.cav_n_events + 1.0, 100.0);
        } else {
            fw.cav_current_intensity = 0.0;
        }
        if (fw.cav_current_intensity > 0.1)
            fw.cav_accumulated_damage += (py_pow(fw.cav_current_intensity, 2.0) * 0.01) * dt;
        fw.cav_noise_increase = fw.cav_current_intensity * 30.0;
        fw.cav_acoustic_signature = 20.0 + fw.cav_noise_increase;
        fw.cav_induced_vibration = fw.cav_current_intensity * 2.0;
        double ir = py_min(1.0, fw.cav_current_intensity / 0.5);
        double dr = py_min(1.0, fw.cav_accumulated_damage / 10.0);
        double fr = py_min(1.0, fw.cav_n_events / 50.0);
        fw.cav_risk_score = (ir * 0.4 + dr * 0.4 + fr * 0.2);
        fw.cav_predicted_damage_rate = (fw.cav_current_intensity > 0) ? py_pow(fw.cav_current_intensity, 2.0) * 0.01 : 0.0;
        tot_risk += fw.cav_risk_score;
        double max_brg = py_max3(u.lub.component_wear[FWL_MOTOR_BRG], u.lub.component_wear[FWL_PUMP_BRG],
                                 u.lub.component_wear[FWL_THRUST_BRG]);
        double tw = (max_brg + u.lub.component_wear[FWL_SEALS]);
        tot_wear += tw;
        wear_for_protection += tw;
        tot_vib += u.vibration_level;
    }
    double avg_risk = tot_risk / 4, avg_wear = tot_wear / 4, avg_vib = tot_vib / 4;
    {
        double ch = py_max(0.0, 1.0 - avg_risk);
        double wh = py_max(0.0, 1.0 - avg_wear / 50.0);
        double vh = py_max(0.0, 1.0 - avg_vib / 10.0);
        fw.diag_health_score = (ch * 0.3 + wh * 0.4 + vh * 0.2 + 1.0 * 0.1);
    }

    // FeedwaterProtectionSystem.check_protection_systems: feedwater/protection_system.py:378-480
    int n_trips = 0;
    {
        const double dt_seconds = dt * 60.0;
        double prot_total_flow = 0.0;
        for (int k = 0; k < 4; ++k) prot_total_flow += fw.pump[k].flow_rate;
        const double npsh_trip = p.fw_prot_low_suction_pressure_trip;  // getattr fallbacks resolve to this field
        for (int k = 0; k < 4; ++k) {   // NPSHProtection.update_npsh_protection :59-124 (one shared instance)
            double npsh = fw.pump[k].npsh_available;
            fw.prot_npsh_low_alarm_active = as_flag(npsh < 18.0);
            if (npsh < npsh_trip) {
                fw.prot_npsh_low_low_timer += dt_seconds;
                if (fw.prot_npsh_low_low_timer >= 5.0) fw.prot_npsh_low_low_trip_active = 1.0;
            } else {
                fw.prot_npsh_low_low_timer = 0.0;
                fw.prot_npsh_low_low_trip_active = 0.0;
            }
            fw.prot_npsh_critical_trip_active = as_flag(npsh < npsh_trip);
            if (is_true(fw.prot_npsh_critical_trip_active)) n_trips++;
            else if (is_true(fw.prot_npsh_low_low_trip_active)) n_trips++;
        }
        for (int k = 0; k < 4; ++k) {   // _check_pressure_protection
            if (fw.pump[k].suction_pressure < p.fw_prot_low_suction_pressure_trip) n_trips++;
            if (fw.pump[k].discharge_pressure > p.fw_prot_high_discharge_pressure_trip) n_trips++;
        }
        {   // _check_flow_protection
            double lf = p.fw_prot_low_flow_trip;
            double low_trip = (lf < 1.0) ? lf * 1500.0 : lf;
            double high_trip = 1.3 * 1500.0;
            if (prot_total_flow < low_trip) {
                fw.prot_timer_low_flow += dt_seconds;
                if (fw.prot_timer_low_flow >= 10.0) n_trips++;
            } else fw.prot_timer_low_flow = 0.0;
            if (prot_total_flow > high_trip) {
                fw.prot_timer_high_flow += dt_seconds;
                if (fw.prot_timer_high_flow >= 2.0) n_trips++;
            } else fw.prot_timer_high_flow = 0.0;
        }
        for (int i = 0; i < 3; ++i) if (sg_levels[i] > 16.5) n_trips++;   // _check_sg_level_protection
        for (int k = 0; k < 4; ++k) {   // _check_equipment_protection (timers shared across pumps)
            const FWPumpState& u = fw.pump[k];
            if (u.vibration_level > 10.0) {
                fw.prot_timer_vibration += dt_seconds;
                if (fw.prot_timer_vibration >= 10.0) n_trips++;
            } else fw.prot_timer_vibration = 0.0;
            double bearing_temp = u.lub.oil_temperature + 5.0;
            if (bearing_temp > 120.0) {
                fw.prot_timer_bearing_temp += dt_seconds;
                if (fw.prot_timer_bearing_temp >= 30.0) n_trips++;
            } else fw.prot_timer_bearing_temp = 0.0;
            if (u.motor_temperature > 130.0) {
                fw.prot_timer_motor_temp += dt_seconds;
                if (fw.prot_timer_motor_temp >= 60.0) n_trips++;
            } else fw.prot_timer_motor_temp = 0.0;
        }
        // _check_diagnostic_protection
        if (fw.diag_health_score < 0.3) n_trips++;
        if (avg_risk > 0.8) n_trips++;
        if (wear_for_protection / 4 > 85.0) n_trips++;
        fw.prot_system_trip_active = as_flag(n_trips > 0);

        if (rep) {   // len(active_alarms): the alarm branches of the same checks (protection_system.py:59-124,483-678)
            int n_alarms = 0;
            for (int k = 0; k < 4; ++k) {
                const FWPumpState& u = fw.pump[k];
                if (u.npsh_available < 18.0) n_alarms++;
                if (!(u.suction_pressure < p.fw_prot_low_suction_pressure_trip) && u.suction_pressure < 0.3) n_alarms++;
                if (!(u.discharge_pressure > p.fw_prot_high_discharge_pressure_trip) && u.discharge_pressure > 9.0) n_alarms++;
                if (u.vibration_level > 5.0) n_alarms++;
                if (u.lub.oil_temperature + 5.0 > 80.0) n_alarms++;
                if (u.motor_temperature > 100.0) n_alarms++;
            }
            {
                double lf = p.fw_prot_low_flow_trip;
                double low_alarm = (lf < 1.0) ? (lf * 1500.0) * 2.0 : 100.0;
                double high_alarm = (1.3 * 1500.0) * 0.9;
                if (prot_total_flow < low_alarm) n_alarms++;
                if (prot_total_flow > high_alarm) n_alarms++;
            }
            for (int i = 0; i < 3; ++i) {
                double level = sg_levels[i];
                if (!(level > 16.5) && level > 15.5) n_alarms++;
                if (level < 10.0) n_alarms++;
                else if (level < 11.0) n_alarms++;
            }
            if (!(fw.diag_health_score < 0.3) && fw.diag_health_score < 0.5) n_alarms++;
            if (!(avg_risk > 0.8) && avg_risk > 0.5) n_alarms++;
            if (!(wear_for_protection / 4 > 85.0) && wear_for_protection / 4 > 50.0) n_alarms++;
            rep->fw_prot_active_alarms_count = (double)n_alarms;
        }
    }

    fw.total_flow_rate = total_flow;
    fw.total_power_consumption = total_power;
    fw.system_availability = as_flag(is_true(fw.pump_system_available) && !is_true(fw.prot_system_trip_active));
    if (fw.total_power_consumption > 0) {
        double hyd = (fw.total_flow_rate * (p.fw_design_pressure - suction_pressure) * 1e6 * 1000 * 9.81) / 1e6;
        fw.system_efficiency = hyd / fw.total_power_consumption;
    } else {
        fw.system_efficiency = 0.0;
    }
    double wq_factor = 1.0 - wc.water_aggressiveness * 0.1;
    fw.performance_factor = avg_perf * wq_factor * fw.diag_health_score;
    fw.maintenance_factor = 1.0;
    fw.operating_hours += dt / 60;

    out.total_flow_rate = fw.total_flow_rate;
    out.total_power_consumption = fw.total_power_consumption;
    out.num_running_pumps = (double)n_running;
    out.system_availability = fw.system_availability;
    double per_sg = (total_flow > 0) ? total_flow / p.fw_num_sg : 0.0;
    for (int i = 0; i < 3; ++i) out.sg_flow[i] = per_sg;
}

}  // namespace nps
