// Shared oil-quality and component-wear model used by the four feedwater-pump lubrication systems
// and the turbine bearing lubrication system.
// Restates BaseLubricationSystem.update_oil_quality / update_component_wear
// (reference: nuclear_simulator/systems/secondary/lubrication_base.py:186-401).
#pragma once
#include "hd.h"
#include "state.h"

namespace nps {

// LubricationComponent constants (lubrication_base.py:76-103)
struct LubComponent {
    double base_wear_rate, load_wear_exponent, speed_wear_exponent, contamination_wear_factor;
    double wear_performance_factor, lubrication_performance_factor;
    double wear_alarm_threshold, wear_trip_threshold;
};

// BaseLubricationConfig limits (lubrication_base.py:54-57); identical for pump and turbine systems
struct LubLimits { double contamination_limit, acidity_limit, moisture_limit, viscosity_change_limit; };

// update_oil_quality: lubrication_base.py:186-352
NPS_HD_SHARED void lub_update_oil_quality(LubCore& L, int n_comp, const LubLimits& lim, double operating_temperature,
                                   double contamination_input, double moisture_input, double dt) {
    // Every field this function needs is read into a local FIRST, as one group of independent loads: the record
    // streams from HBM, and a load issued at its point of use costs a full DRAM round trip each.
    double oil_temperature = L.oil_temperature, oil_contamination_level = L.oil_contamination_level;
    double oil_moisture_content = L.oil_moisture_content, oil_acidity_number = L.oil_acidity_number;
    double oil_viscosity_change = L.oil_viscosity_change, antioxidant_level = L.antioxidant_level;
    double anti_wear_additive_level = L.anti_wear_additive_level, corrosion_inhibitor_level = L.corrosion_inhibitor_level;
    double oil_operating_hours = L.oil_operating_hours, operating_hours = L.operating_hours;
    double wear[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) wear[i] = (i < n_comp) ? L.component_wear[i] : 0.0;

    double temp_change = (operating_temperature - oil_temperature) / 0.5 * dt;
    double max_temp_change = 10.0 * dt;
    temp_change = py_max(-max_temp_change, py_min(max_temp_change, temp_change));
    oil_temperature += temp_change;
    oil_temperature = py_max(20.0, py_min(120.0, oil_temperature));

    double temp_diff = oil_temperature - 60.0;
    temp_diff = py_max(-50.0, py_min(200.0, temp_diff));
    double activation_factor = py_max(0.1, py_min(1.5, 1.0 + temp_diff / 50.0));
    double thermal_degradation_rate = 0.00001 * activation_factor;

    double filter_loading_factor = py_max(0.3, 1.0 - (oil_contamination_level / 50.0));
    double temp_factor = py_max(0.5, 1.0 - (oil_temperature - 60.0) / 60.0);
    double eff_filtration = 0.60 * filter_loading_factor * temp_factor;
    double contamination_removal_rate = oil_contamination_level * eff_filtration * 0.005;
    double base_thermal_contamination = thermal_degradation_rate * 0.75;
    double wear_sum = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) if (i < n_comp) wear_sum += wear[i];
    double avg_wear = wear_sum / n_comp;
    double thermal_contamination_input = base_thermal_contamination * (1.0 + (avg_wear / 20.0));

    double contamination_change = contamination_input - contamination_removal_rate + thermal_contamination_input;
    oil_contamination_level += contamination_change * dt;
    oil_contamination_level = py_max(1.0, oil_contamination_level);

    double moisture_change;
    if (oil_temperature > 70.0) {
        double evaporation_rate = (oil_temperature - 70.0) * 0.001;
        moisture_change = moisture_input - evaporation_rate;
    } else {
        moisture_change = moisture_input;
    }
    oil_moisture_content += moisture_change * dt;
    oil_moisture_content = py_max(0.001, oil_moisture_content);

    double contamination_factor = 1.0 + oil_contamination_level / 50.0;
    double acidity_increase_rate = thermal_degradation_rate * contamination_factor * 0.1;
    oil_acidity_number += acidity_increase_rate * dt;

    double viscosity_change_rate = thermal_degradation_rate * 0.5 + contamination_change * 0.01;
    oil_viscosity_change += viscosity_change_rate * dt;

    double antioxidant_rate = thermal_degradation_rate * 10.0;
    antioxidant_level = py_max(0.0, antioxidant_level - antioxidant_rate * dt * 100.0);
    double aw_rate = (contamination_input * 0.1) * 0.5;
    anti_wear_additive_level = py_max(0.0, anti_wear_additive_level - aw_rate * dt * 100.0);
    double ci_rate = oil_moisture_content * 2.0;
    corrosion_inhibitor_level = py_max(0.0, corrosion_inhibitor_level - ci_rate * dt * 100.0);

    contamination_factor = py_max(0.1, 1.0 - oil_contamination_level / lim.contamination_limit);
    double acidity_factor = py_max(0.1, 1.0 - oil_acidity_number / lim.acidity_limit);
    double moisture_factor = py_max(0.1, 1.0 - oil_moisture_content / lim.moisture_limit);
    double viscosity_factor = py_max(0.1, 1.0 - fabs(oil_viscosity_change) / lim.viscosity_change_limit);
    double antioxidant_factor = antioxidant_level / 100.0;
    double aw_factor = anti_wear_additive_level / 100.0;
    double critical = py_pow(contamination_factor * antioxidant_factor * aw_factor, 1.0 / 3);
    double secondary = (0.0 + acidity_factor + moisture_factor + viscosity_factor) / 3;
    double effectiveness = critical * 0.7 + secondary * 0.3;
    effectiveness = py_max(0.3, py_min(1.0, effectiveness));

    L.oil_temperature = oil_temperature; L.oil_contamination_level = oil_contamination_level;
    L.oil_moisture_content = oil_moisture_content; L.oil_acidity_number = oil_acidity_number;
    L.oil_viscosity_change = oil_viscosity_change; L.antioxidant_level = antioxidant_level;
    L.anti_wear_additive_level = anti_wear_additive_level; L.corrosion_inhibitor_level = corrosion_inhibitor_level;
    L.lubrication_effectiveness = effectiveness;
    L.oil_operating_hours = oil_operating_hours + dt;
    L.operating_hours = operating_hours + dt;
}

// One iteration of the loop in update_component_wear (lubrication_base.py:368-391)
NPS_HD void lub_apply_component_wear(LubCore& L, int i, const LubComponent& c, double wear_rate, double dt) {
    double lubrication_wear_factor = 1.0 + (1.0 - L.lubrication_effectiveness) * c.contamination_wear_factor;
    double actual = wear_rate * lubrication_wear_factor;
    L.component_wear[i] += actual * dt;
    double wear_loss = L.component_wear[i] * c.wear_performance_factor;
    double lub_loss = (1.0 - L.lubrication_effectiveness) * c.lubrication_performance_factor;
    L.component_perf[i] = py_max(0.1, 1.0 - (wear_loss + lub_loss));
}

// Tail of update_component_wear (lubrication_base.py:398-399)
NPS_HD void lub_update_health(LubCore& L, int n_comp) {
    double s = 0.0;
    for (int i = 0; i < n_comp; ++i) s += L.component_perf[i];
    L.system_health_factor = (s / n_comp) * L.lubrication_effectiveness;
}

}  // namespace nps
