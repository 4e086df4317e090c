// One plant, one timestep: NuclearPlantSimulator.step
// (reference: nuclear_simulator/simulator/core/sim.py:130-258), physics part.
// Maintenance/threshold monitoring (sim.py:209-223) runs in the flag kernel, not here.
#pragma once
#include "hd.h"
#include "state.h"
#include "primary.h"

namespace nps {

NPS_HD void plant_step(PlantState& st, const PlantParams& p, const StepInput& in) {
    const double dt = p.dt;
    primary_update(st.pri, p, in, dt);
    // state_manager.advance_time(dt) / self.time += dt : sim.py:183-194 (minutes)
    st.sim.time_minutes += dt;
}

// get_observation: sim.py:290-333 (first 12 entries are primary-only)
NPS_HD void plant_observe_primary(const PlantState& st, double* obs) {
    const PrimaryState& s = st.pri;
    obs[0] = s.neutron_flux / 1e12;
    obs[1] = s.fuel_temperature / 1000;
    obs[2] = s.coolant_temperature / 300;
    obs[3] = s.coolant_pressure / 20;
    obs[4] = s.coolant_flow_rate / 50000;
    obs[5] = s.steam_temperature / 300;
    obs[6] = s.steam_pressure / 10;
    obs[7] = s.steam_flow_rate / 3000;
    obs[8] = s.control_rod_position / 100;
    obs[9] = s.steam_valve_position / 100;
    obs[10] = s.power_level / 100;
    obs[11] = is_true(s.scram_status) ? 1.0 : 0.0;
}

}  // namespace nps
