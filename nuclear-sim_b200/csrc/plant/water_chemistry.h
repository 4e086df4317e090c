// Unified water chemistry (three instances per plant: WAT-003 shared with feedwater and updated
// twice per step, WAT-002 owned by the condenser, WAT-001 owned by the SG system and never updated).
// Restates systems/secondary/water_chemistry.py:277-438, 659-760.  All instances are built with
// the default WaterChemistryConfig (water_chemistry.py:70-133), whose values appear as literals.
#pragma once
#include "hd.h"
#include "state.h"

namespace nps {

struct MakeupWater { double ph, hardness, tds, chloride, dissolved_oxygen; };

// _calculate_composite_parameters: water_chemistry.py:277-320
NPS_HD_SHARED void wc_composite(WaterChemState& w) {
    double iron_effect = w.iron_concentration * 0.5;
    double chloride_effect = w.chloride / 100.0;
    double ph_effect = fabs(w.ph - 7.0) * 0.2;
    double hardness_effect = py_max(0.0, (w.hardness - 150.0) / 150.0) * 0.3;
    w.water_aggressiveness = 1.0 + iron_effect + chloride_effect + ph_effect + hardness_effect;
    w.water_aggressiveness = np_clip(w.water_aggressiveness, 0.5, 3.0);
    double tds_factor = w.total_dissolved_solids / 500.0;
    double iron_particle = w.iron_concentration * 2.0;
    double silica_particle = w.silica_concentration / 20.0;
    w.particle_content = 1.0 + (tds_factor + iron_particle + silica_particle) * 0.1;
    w.particle_content = np_clip(w.particle_content, 0.5, 2.0);
    double A = (nps_log10(w.total_dissolved_solids) - 1) / 10;
    double B = -13.12 * log10(25.0 + 273) + 34.55;   // constant argument: inlined so that it folds
    double C = nps_log10(w.hardness) - 0.4;
    double D = nps_log10(w.alkalinity);
    double ph_sat = (9.3 + A + B) - (C + D);
    w.scaling_tendency = w.ph - ph_sat;
    w.corrosion_tendency = 2 * ph_sat - w.ph;
    double ph_stab = 1.0 - fabs(w.ph - 9.2) / 2.0;
    double conc_stab = 1.0 - fabs(w.concentration_factor - 2.0) / 3.0;
    w.chemistry_stability_factor = (ph_stab + w.treatment_efficiency + conc_stab) / 3.0;
    w.chemistry_stability_factor = np_clip(w.chemistry_stability_factor, 0.1, 1.0);
}

// update_chemistry: water_chemistry.py:322-389.  `has_makeup` mirrors `if makeup_water_quality:`.
NPS_HD_SHARED void wc_update(WaterChemState& w, bool has_makeup, const MakeupWater& mk, double blowdown, double dt) {
    NPS_TOUCH(w.operating_hours); NPS_TOUCH(w.last_treatment_time); NPS_TOUCH(w.pending_effects); NPS_TOUCH(w.pend_ammonia_dose_rate); NPS_TOUCH(w.pend_morpholine_dose_rate); NPS_TOUCH(w.ph); NPS_TOUCH(w.pend_ph_setpoint); NPS_TOUCH(w.antiscalant_concentration); NPS_TOUCH(w.corrosion_inhibitor_level); NPS_TOUCH(w.hardness); NPS_TOUCH(w.total_dissolved_solids); NPS_TOUCH(w.chloride); NPS_TOUCH(w.chlorine_residual); NPS_TOUCH(w.iron_concentration); NPS_TOUCH(w.silica_concentration); NPS_TOUCH(w.alkalinity);
    double dt_hours;
    if (dt > 100) dt_hours = dt / 3600.0;
    else if (dt > 1) dt_hours = dt / 60.0;
    else dt_hours = dt;
    w.operating_hours += dt_hours;
    w.last_treatment_time += dt_hours;

    if (is_true(w.pending_effects)) {
        // 'ph_control' -> _apply_ph_control_effects: water_chemistry.py:683-718
        double ammonia = w.pend_ammonia_dose_rate, morpholine = w.pend_morpholine_dose_rate;
        if (ammonia > 0) {
            double inc_ppm = (ammonia / 3600.0) / (1000.0 * 1000.0) * 1e6;
            w.ph = py_min(w.ph + inc_ppm * 0.1, 9.6);
        }
        if (morpholine > 0) {
            double inc_ppm = (morpholine / 3600.0) / (1000.0 * 1000.0) * 1e6;
            w.ph = py_min(w.ph + inc_ppm * 0.05, 9.6);
        }
        if (fabs(w.ph - w.pend_ph_setpoint) > 0.01) w.ph += (w.pend_ph_setpoint - w.ph) * 0.3;
        // 'chemical_additions' -> _apply_chemical_additions: water_chemistry.py:720-740
        // (rates were queued as dose/3600 kg/s at systems/secondary/__init__.py:660-663)
        w.antiscalant_concentration += (ammonia / 3600.0) * 3600.0 * 0.1;
        w.corrosion_inhibitor_level += (morpholine / 3600.0) * 3600.0 * 0.05;
        w.pending_effects = 0.0;
    }
    if (has_makeup) {  // _update_from_makeup_water: water_chemistry.py:391-412
        double blend = 0.05 * dt_hours * 0.1;
        blend = py_min(blend, 0.5);
        w.ph += (mk.ph - w.ph) * blend;
        w.hardness += (mk.hardness - w.hardness) * blend;
        w.total_dissolved_solids += (mk.tds - w.total_dissolved_solids) * blend;
        w.chloride += (mk.chloride - w.chloride) * blend;
        w.dissolved_oxygen = mk.dissolved_oxygen * 0.8;
    }
    w.concentration_factor = 1.0 / (blowdown + 0.01);
    w.concentration_factor = py_min(w.concentration_factor, 5.0);
    if (w.concentration_factor > 1.1) {
        double inc = (w.concentration_factor - 1.0) * 0.1 * dt_hours;
        w.total_dissolved_solids += inc * 50.0;
        w.hardness += inc * 10.0;
        w.chloride += inc * 5.0;
    }
    {   // _update_chemical_treatment: water_chemistry.py:414-438
        double dose_rate = 0.5 * dt_hours;
        w.antiscalant_concentration += (5.0 - w.antiscalant_concentration) * dose_rate;
        w.corrosion_inhibitor_level += (10.0 - w.corrosion_inhibitor_level) * dose_rate;
        double decay = 0.1 * dt_hours;
        w.chlorine_residual *= nps_exp(-decay);
        w.chlorine_residual += (1.0 - w.chlorine_residual) * dose_rate;
        double ce = (w.chlorine_residual > 0.2) ? 1.0 : 0.5;
        double ae = (w.antiscalant_concentration > 2.0) ? 1.0 : 0.7;
        double ke = (w.corrosion_inhibitor_level > 5.0) ? 1.0 : 0.8;
        w.treatment_efficiency = (ce * ae * ke * 0.95);
    }
    wc_composite(w);
}

// update_chemistry_effects: water_chemistry.py:659-682 (queues for the next update_chemistry)
NPS_HD void wc_queue_effects(WaterChemState& w, double ph_setpoint, double ammonia, double morpholine) {
    w.pending_effects = 1.0;
    w.pend_ph_setpoint = ph_setpoint;
    w.pend_ammonia_dose_rate = ammonia;
    w.pend_morpholine_dose_rate = morpholine;
}

}  // namespace nps
