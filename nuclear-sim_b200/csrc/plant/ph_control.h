// pH control system: sensor lag + noise + drift, PID, dosing, tank depletion, random equipment
// failures and the 100-sample RMS deviation metric.
// Restates PHControlSystem.update_system / PHController.update_controller
// (reference: nuclear_simulator/systems/secondary/ph_control_system.py:534-565, 219-478);
// the controller is always built with the default PHControllerConfig (:84-136).
// Random draws are host-supplied: one standard normal z and up to three uniforms u[] per step,
// consumed with the reference's short-circuit order.
#pragma once
#include "hd.h"
#include "state.h"

namespace nps {

// numpy float64 add.reduce over n <= 128 contiguous values (pairwise_sum block for n >= 8)
template <class Get>
NPS_HD double np_sum_small(int n, Get get) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += get(i);
        return r;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = get(j);
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] += get(i + j);
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += get(i);
    return 0.0 + res;
}

NPS_HD void ph_control_update(PHControlState& s, double current_ph, double dt, double z, const double* u,
                              bool emit_outputs = true) {
    NPS_TOUCH(s.ph_sensor_status); NPS_TOUCH(s.measured_ph); NPS_TOUCH(s.ph_setpoint); NPS_TOUCH(s.ammonia_tank_level); NPS_TOUCH(s.morpholine_tank_level); NPS_TOUCH(s.control_mode); NPS_TOUCH(s.controller_enabled); NPS_TOUCH(s.integral_sum); NPS_TOUCH(s.previous_error); NPS_TOUCH(s.ammonia_supply_available); NPS_TOUCH(s.ammonia_pump_status); NPS_TOUCH(s.dev_count); NPS_TOUCH(s.dev_head); NPS_TOUCH(s.tic_initialized); NPS_TOUCH(s.tic_sum); NPS_TOUCH(s.tic_total_time); NPS_TOUCH(s.total_chemical_consumed); NPS_TOUCH(s.control_actions_count);
    const double dt_minutes = dt * 60.0;
    s.operating_hours += dt;
    // _apply_sensor_dynamics: :274-291
    if (!is_true(s.ph_sensor_status)) {
        s.measured_ph = s.measured_ph + (0.0 + 0.01 * z);
    } else {
        double alpha = dt_minutes / (1.0 + dt_minutes);
        double filtered = s.measured_ph + alpha * (current_ph - s.measured_ph);
        double noise = 0.0 + 0.005 * z;
        double drift = 0.001 * dt_minutes / 60.0;
        s.measured_ph = filtered + noise + drift;
    }
    s.ph_error = s.ph_setpoint - s.measured_ph;
    // _update_alarms_and_trips: :424-441
    s.ph_low_alarm = as_flag(s.measured_ph < 8.8);
    s.ph_high_alarm = as_flag(s.measured_ph > 9.6);
    s.low_chemical_alarm = as_flag(s.ammonia_tank_level < 20.0 || s.morpholine_tank_level < 20.0);
    bool ph_trip = (s.measured_ph < 8.5 || s.measured_ph > 10.0);
    if (ph_trip && (int)s.control_mode == 0) { s.control_mode = 2.0; s.controller_enabled = 0.0; }
    double output;
    if ((int)s.control_mode == 0 && is_true(s.controller_enabled)) {   // _calculate_pid_output: :293-325
        double error = s.ph_error;
        s.proportional_term = 2.0 * error;
        s.integral_sum += error * dt_minutes;
        s.integral_sum = np_clip(s.integral_sum, -50.0, 50.0);
        s.integral_term = 0.1 * s.integral_sum;
        if (dt_minutes > 0) s.derivative_term = 0.5 * ((error - s.previous_error) / dt_minutes);
        else s.derivative_term = 0.0;
        double o = (s.proportional_term + s.integral_term + s.derivative_term);
        s.previous_error = error;
        output = np_clip(o, 0.0, 100.0);
    } else if ((int)s.control_mode == 1) {
        output = s.manual_output;
    } else {
        output = 0.0;
    }
    // _apply_rate_limiting is inert: its history list is never seeded (:329-330)
    s.controller_output = output;
    // _calculate_dosing_rates: :350-377
    if (is_true(s.ammonia_supply_available) && is_true(s.ammonia_pump_status) && output > 0)
        s.ammonia_dose_rate = (output / 100.0) * 5.0 * 0.95;
    else s.ammonia_dose_rate = 0.0;
    if (!is_true(s.ammonia_supply_available) && is_true(s.morpholine_supply_available) &&
        is_true(s.morpholine_pump_status) && output > 0)
        s.morpholine_dose_rate = (output / 100.0) * 10.0 * 0.98;
    else s.morpholine_dose_rate = 0.0;
    s.chemical_consumption_rate = (s.ammonia_dose_rate + s.morpholine_dose_rate);
    // _update_chemical_supplies: :379-401
    if (s.ammonia_dose_rate > 0) {
        double dec = ((s.ammonia_dose_rate * dt) / 1000.0) * 100.0;
        s.ammonia_tank_level = py_max(0.0, s.ammonia_tank_level - dec);
    }
    if (s.morpholine_dose_rate > 0) {
        double dec = ((s.morpholine_dose_rate * dt) / 2000.0) * 100.0;
        s.morpholine_tank_level = py_max(0.0, s.morpholine_tank_level - dec);
    }
    s.ammonia_supply_available = as_flag(s.ammonia_tank_level > 5.0);
    s.morpholine_supply_available = as_flag(s.morpholine_tank_level > 5.0);
    // _update_equipment_status: :403-422 (draws consumed only while the equipment is healthy)
    {
        double fp = dt / 8760.0;
        int k = 0;
        if (is_true(s.ammonia_pump_status)) { if (u[k++] < fp) { s.ammonia_pump_status = 0.0; s.equipment_failure_alarm = 1.0; } }
        if (is_true(s.morpholine_pump_status)) { if (u[k++] < fp) { s.morpholine_pump_status = 0.0; s.equipment_failure_alarm = 1.0; } }
        double sp = fp * 0.5;
        if (is_true(s.ph_sensor_status)) { if (u[k++] < sp) { s.ph_sensor_status = 0.0; s.equipment_failure_alarm = 1.0; } }
    }
    // _update_performance_metrics: :443-466
    {
        double deviation = fabs(s.ph_error);
        int n = (int)s.dev_count, head = (int)s.dev_head;
        if (n < 100) { s.dev_hist[(head + n) % 100] = deviation; n += 1; }
        else { s.dev_hist[head] = deviation; head = (head + 1) % 100; }
        s.dev_count = (double)n; s.dev_head = (double)head;
        if (emit_outputs) {   // logged only (ph_control_system.py:455); reads the whole 100-sample window
            const double* h = s.dev_hist;
            double sum = np_sum_small(n, [&](int i) { double v = h[(head + i) % 100]; return v * v; });
            s.control_deviation_rms = sqrt(sum / n);
        }
        bool in_control = deviation <= 0.05;
        if (is_true(s.tic_initialized)) {
            s.tic_sum += in_control ? dt : 0.0;
            s.tic_total_time += dt;
            s.time_in_control = (s.tic_sum / s.tic_total_time) * 100.0;
        } else {
            s.tic_initialized = 1.0;
            s.tic_sum = in_control ? dt : 0.0;
            s.tic_total_time = dt;
            s.time_in_control = in_control ? 100.0 : 0.0;
        }
    }
    // PHControlSystem.update_system: :534-565
    s.total_chemical_consumed += (s.ammonia_dose_rate + s.morpholine_dose_rate) * dt;
    if (s.controller_output > 0) s.control_actions_count += 1.0;
}

}  // namespace nps
