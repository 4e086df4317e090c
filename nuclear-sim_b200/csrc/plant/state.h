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
    double feedwater_pump_status;
    double feedwater_pump_speed;
    double feedwater_system_available;
    double feedwater_pump_power;
    double feedwater_num_running_pumps;
    double xenon_concentration;
    double iodine_concentration;
    double samarium_concentration;
    double burnable_poison_worth;
    double fuel_burnup;
    double power_level;
    double scram_status;
    // PrimaryReactorPhysics members: systems/primary/__init__.py:168-171
    double thermal_power_mw;
    double total_reactivity_pcm;
    double scram_activated;
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
    double has_last_heat_removal_factor;
    double last_heat_removal_factor;
    double last_load_factor;
    double last_feedwater_flow_factor;
    double last_pump_reliability_factor;
};

struct PlantState {
    PrimaryState pri;
    SimState sim;
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
};

}  // namespace nps
