"""TEST INFRASTRUCTURE — generate committed fixtures from the LIVE reference (/root/reference).

Run in the build container (the reference is absent on the GPU box):
    python oracle/make_golden.py            # all scenarios
    python oracle/make_golden.py cfg1 cfg4  # selected

Writes
  nuclear-sim_b200/data/<snapshot>.npz   initial PlantState + PlantParams vectors (engine starting points)
  tests/golden/<scenario>.npz            inputs + reference states/observations at checkpoint steps

Every array is produced by stepping the unmodified reference ``NuclearPlantSimulator.step``
(nuclear_simulator/simulator/core/sim.py:130-258) with host-controlled random streams
(oracle/refplant.py); nothing here goes through the C/CUDA restatement.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
sys.path.insert(0, _REPO)
from oracle import refplant as R  # noqa: E402

GOLDEN = os.path.join(_REPO, "tests", "golden")
DATA = os.path.join(_REPO, "nuclear-sim_b200", "data")


def _names():
    L = R._layout()
    return np.array(L.field_names("PlantState")), np.array(L.field_names("PlantParams"))


def save_snapshot(name, sim):
    os.makedirs(DATA, exist_ok=True)
    sn, pn = _names()
    np.savez_compressed(os.path.join(DATA, name + ".npz"), state=R.extract_state(sim), params=R.extract_params(sim),
                        state_names=sn, param_names=pn)


def run_scenario(name, plants, T, checkpoints, policy, *, inject=None, seed_base=1000, strict=True):
    """plants: list of ReferencePlant; policy(p, t, sim) -> (action, magnitude, setpoint_or_nan)."""
    P = len(plants)
    L = R._layout()
    NS = L.N_STATE
    params = [R.extract_params(rp.sim, strict) for rp in plants]
    for q in params[1:]:
        assert np.array_equal(q, params[0]), "plants of one golden must share PlantParams"
    state0 = np.stack([R.extract_state(rp.sim, strict) for rp in plants])
    actions = np.full((T, P), 8, dtype=np.int8)
    mags = np.ones((T, P))
    noise = np.zeros((T, P, 5))
    setp = np.full((T, P), np.nan)
    inj = np.full((T, P, 2), np.nan)   # (field index, value) written to the state BEFORE step t
    cps = sorted(set(int(c) for c in checkpoints if c <= T))
    states = np.zeros((len(cps), P, NS))
    obs = np.zeros((len(cps), P, 22))
    rew = np.zeros((len(cps), P))
    done_step = np.full(P, -1, dtype=np.int64)
    power = np.zeros((T, P))
    elec = np.zeros((T, P))
    ix = L.field_index()
    t0 = time.time()
    for p, rp in enumerate(plants):
        rng = np.random.RandomState(seed_base + p)
        sim = rp.sim
        for t in range(T):
            a, m, sp = policy(p, t, sim)
            z = np.array([rng.standard_normal(), rng.standard_normal(), rng.random_sample(), rng.random_sample(),
                          rng.random_sample()])
            actions[t, p], mags[t, p], noise[t, p], setp[t, p] = a, m, z, sp
            if inject is not None:
                ev = inject(p, t)
                if ev is not None:
                    fname, val = ev
                    assert fname.startswith("pri.")
                    setattr(sim.primary_physics.state, fname[4:], val)
                    inj[t, p] = (ix[fname], val)
            if not np.isnan(sp):
                sim.primary_physics.heat_source.set_power_setpoint(sp)
            out = rp.step(int(a), float(m), z)
            power[t, p] = sim.state.power_level
            elec[t, p] = sim.secondary_physics.electrical_power_output if sim.secondary_physics else 0.0
            if out["done"] and done_step[p] < 0:
                done_step[p] = t
            if (t + 1) in cps:
                c = cps.index(t + 1)
                states[c, p] = R.extract_state(sim, strict)
                o = np.asarray(out["observation"], dtype=np.float64)      # 12 entries without a secondary side: stored with a zero tail
                obs[c, p, :len(o)] = o
                rew[c, p] = out["reward"]
    os.makedirs(GOLDEN, exist_ok=True)
    sn, pn = _names()
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), state0=state0, params=params[0], actions=actions,
                        magnitudes=mags, noise=noise, setpoint=setp, inject=inj, checkpoints=np.array(cps),
                        states=states, obs=obs, reward=rew, done_step=done_step, power_level=power, electrical=elec,
                        state_names=sn, param_names=pn)
    print(f"[golden] {name}: P={P} T={T} checkpoints={cps} done_step={done_step.tolist()} ({time.time()-t0:.1f}s)")


NO = (8, 1.0, np.nan)


def cfg1():
    """BASELINE config #1: oil_top_off, dt = 5 min, 12 steps, ConstantHeatSource noise 0.1 % (seed-42 stream),
    runner-style setpoint ramp 90 % -> target at 0.02 %/step (maintenance_scenario_runner.py:651-671)."""
    cfg = R.compose_config("oil_top_off")
    rp = R.make_reference_plant(cfg, dt=5.0, heat_source="constant", noise_enabled=True, noise_std_percent=0.1)
    save_snapshot("pwr3000_oil_top_off_dt5", rp.sim)
    rp.sim.primary_physics.heat_source.set_power_setpoint(90.0)

    def policy(p, t, sim):
        return 8, 1.0, 90.0 + 0.02 * (t + 1)
    run_scenario("cfg1_oil_top_off", [rp], 12, range(1, 13), policy, seed_base=42)


def _plants(actions, **kw):
    out = []
    for a in actions:
        out.append(R.make_reference_plant(R.compose_config(a), **kw))
    return out


IC_ACTIONS = ["oil_top_off", "tsp_chemical_cleaning", "scale_removal", "oil_change"]


def cfg2():
    """BASELINE config #2: steady 100 %, constant heat source (noise off), NO_ACTION, 1 h at dt = 1.0 (3 600 steps)."""
    plants = _plants(IC_ACTIONS[:2], dt=1.0, heat_source="constant", noise_enabled=False)
    save_snapshot("pwr3000_steady_dt1", plants[0].sim)
    run_scenario("cfg2_steady", plants, 3600, [1, 2, 10, 100, 1000, 3600], lambda p, t, sim: NO)


def cfg3():
    """BASELINE config #3: ReactorHeatSource at create_equilibrium_state(); load-following rods
    (data/gen_training_data.py:401-406), power ramp 20xINSERT then 30xWITHDRAW (tests/test_scenarios.py:56-74),
    boron / flow / valve actions mixed in; dt = 1.0, 3 600 steps; per-plant ICs from the action catalog."""
    plants = _plants(IC_ACTIONS[:3], dt=1.0, heat_source="reactor")
    save_snapshot("pwr3000_reactor_dt1", plants[0].sim)

    def policy(p, t, sim):
        mag = 0.2 + 0.8 * ((p * 37 + 11) % 100) / 100.0
        if p == 0:
            s = np.sin(t / 100.0)
            return (1 if s > 0.5 else (0 if s < -0.5 else 8)), mag, np.nan
        if p == 1:
            u = t % 200
            return (0 if u < 20 else (1 if u < 50 else 8)), mag, np.nan
        seq = [10, 8, 8, 9, 8, 2, 3, 4, 5, 6, 7, 8]
        return seq[(t // 25) % len(seq)], mag, np.nan
    run_scenario("cfg3_loadfollow", plants, 3600, [1, 10, 100, 500, 1000, 2000, 3600], policy)


def cfg4():
    """BASELINE config #4: scram / shutdown transients — injected coolant pressure 17.3 MPa (> 17.2 trip,
    scram_logic.py:20), injected fuel temperature 1600 C (tests/test_scenarios.py:98-110), flow reduction to the
    5 000 kg/s clamp (scram_logic.py:21,37: the clamp value itself does not trip) and a boration shutdown
    (power falls 10 %/s through the point-kinetics rate clamp, no scram)."""
    plants = _plants(["oil_top_off"] * 4, dt=1.0, heat_source="reactor")

    def policy(p, t, sim):
        if p == 2 and t >= 20:
            return 3, 1.0, np.nan
        if p == 3 and 30 <= t < 60:
            return 10, 1.0, np.nan
        return NO

    def inject(p, t):
        if p == 0 and t == 50:
            return ("pri.coolant_pressure", 17.3)
        if p == 1 and t == 100:
            return ("pri.fuel_temperature", 1600.0)
        return None
    run_scenario("cfg4_scram", plants, 300, [1, 50, 52, 54, 56, 60, 100, 101, 102, 150, 300], policy, inject=inject)


def cfg5():
    """BASELINE config #5 (one day of it): long-horizon degradation at dt = 5 min, 288 steps, ICs near
    maintenance thresholds, physics only (work-order effects are a 'next' row)."""
    plants = _plants(["tsp_chemical_cleaning", "oil_change", "scale_removal"], dt=5.0, heat_source="constant",
                     noise_enabled=False)
    run_scenario("cfg5_degradation", plants, 288, [1, 12, 144, 288], lambda p, t, sim: NO)


def cfg6():
    """Secondary-side trips and alarms (BASELINE config #4's "safety-system trips with divergent per-plant control flow"
    on the secondary side): each plant starts from the standard plant with ONE degraded condition written into the
    reference's own objects before the first step, so every protection / trip branch runs in the reference itself."""
    plants = _plants(["oil_top_off"] * 8, dt=1.0, heat_source="constant", noise_enabled=False)

    def pumps(rp):
        return list(rp.sim.secondary_physics.feedwater_system.pump_system.pumps.values())
    # 0: lubrication oil nearly gone on FWP-1 (level alarms / lubrication trip path, pump_lubrication.py:1536-1580)
    pumps(plants[0])[0].lubrication_system.oil_level = 6.0
    # 1: NPSH collapse on FWP-2 (frozen at its IC value by the _initial_conditions_applied quirk): cavitation + NPSH trips
    pumps(plants[1])[1].state.npsh_available = 6.0
    # 2: SG-1 level above the 16.0 m pump-trip limit (pump_system.py:277-333)
    plants[2].sim.secondary_physics.steam_generator_system.steam_generators[1].water_level = 16.3
    # 3: hot turbine bearing (bearing-temperature trip with its delay timer, enhanced_physics.py:348-437)
    list(plants[3].sim.secondary_physics.turbine.rotor_dynamics.bearings.values())[1].metal_temperature = 135.0
    # 4: condenser air in-leakage x60 (vacuum alarms / low-vacuum turbine trip, vacuum_system.py:407-423)
    plants[4].sim.secondary_physics.condenser.vacuum_system.current_air_leakage *= 60.0
    # 5: badly contaminated, wet, acidic oil on FWP-3 (oil-quality alarms and trips, lubrication_base.py:434-500)
    L = pumps(plants[5])[2].lubrication_system
    L.oil_contamination_level, L.oil_moisture_content, L.oil_acidity_number = 48.0, 0.5, 3.5
    # 6: feedwater pH far out of band (pH controller FAILED mode, ph_control_system.py:396-441)
    plants[6].sim.secondary_physics.water_chemistry.ph = 7.2
    # 7: heavy wear on every FWP-4 component (performance degradation, vibration, wear trips)
    W = pumps(plants[7])[3].lubrication_system.component_wear
    for k in W:
        W[k] = 45.0
    run_scenario("cfg6_secondary_trips", plants, 240, [1, 2, 3, 5, 8, 12, 20, 40, 80, 160, 240], lambda p, t, sim: NO)


def cfg7():
    """Turbine protection trips and severe fouling states, prepared the same way as cfg6."""
    plants = _plants(["oil_top_off"] * 6, dt=1.0, heat_source="constant", noise_enabled=False)
    sec = lambda i: plants[i].sim.secondary_physics
    # 0: rotor overspeed (overspeed trip + overspeed event counter, enhanced_physics.py:348-437, rotor_dynamics.py:897-902)
    sec(0).turbine.rotor_dynamics.rotor_speed = 4100.0
    # 1: all four bearings hot (bearing-temperature trip after its delay)
    for b in sec(1).turbine.rotor_dynamics.bearings.values():
        b.metal_temperature = 150.0
    # 2: thick TSP deposits on SG-0 (severe / critical fouling stage, shutdown and replacement flags, tsp_fouling_model.py:394-445)
    d = sec(2).steam_generator_system.steam_generators[0].tsp_fouling.deposits
    for lv in range(7):
        d.magnetite_thickness[lv], d.copper_thickness[lv], d.silica_thickness[lv], d.biological_thickness[lv] = 3.6, 1.8, 2.7, 0.9
    # 3: thick tube-interior scale on SG-2 (replacement flag, thermal resistance saturation)
    t = sec(3).steam_generator_system.steam_generators[2].tube_interior_fouling
    t.scale_thickness = 2.6
    t.scale_composition = {k: v * 2.6 / max(1e-9, sum(t.scale_composition.values())) for k, v in t.scale_composition.items()} \
        if sum(t.scale_composition.values()) > 0 else t.scale_composition
    # 4: a quarter of the condenser tubes plugged, thick fouling (heat-transfer limit, vacuum degradation)
    c = sec(4).condenser
    c.tube_degradation.plugged_tube_count = 21000
    c.tube_degradation.active_tube_count = 63000
    c.fouling_model.biofouling_thickness, c.fouling_model.scale_thickness = 1.5, 0.8
    # 5: large rotor vibration / thrust displacement
    rd = sec(5).turbine.rotor_dynamics
    for b in rd.bearings.values():
        b.vibration_displacement = 30.0
    rd.thrust_bearing_displacement = 1.4 if hasattr(rd, "thrust_bearing_displacement") else 0.0
    run_scenario("cfg7_turbine_trips_fouling", plants, 240, [1, 2, 3, 5, 8, 12, 20, 40, 80, 160, 240], lambda p, t, sim: NO)


def _set(obj, **kw):
    for k, v in kw.items():
        assert hasattr(obj, k), k
        setattr(obj, k, v)


# Protection limits moved so that the NORMAL operating point violates them: every trip / alarm branch of the turbine
# protection, the rotor vibration monitor, the condenser vacuum system and the feedwater protection runs in the reference
# itself and stays latched for the rest of the run.  The limits are PlantParams, so each case is its own one-plant fixture.
TRIP_CASES = {
    "overspeed": lambda sp: _set(sp.turbine.protection_system.config, overspeed_trip=3500.0),
    "vibration": lambda sp: _set(sp.turbine.protection_system.config, vibration_trip=-1.0, vibration_delay=3.0),
    "bearing_temp": lambda sp: _set(sp.turbine.protection_system.config, bearing_temp_trip=35.0, bearing_temp_delay=4.0),
    "thrust_bearing": lambda sp: _set(sp.turbine.protection_system.config, thrust_bearing_trip=1e-4),
    "low_vacuum": lambda sp: _set(sp.turbine.protection_system.config, low_vacuum_trip=0.004),
    "thermal_stress": lambda sp: _set(sp.turbine.protection_system.config, max_thermal_stress=1.0e3),
    "rotor_vibration_alarms": lambda sp: _set(sp.turbine.rotor_dynamics.config, displacement_alarm=-1.0, velocity_alarm=-1.0,
                                              acceleration_alarm=-1.0, first_critical_speed=3500.0),
    "vacuum_alarms": lambda sp: _set(sp.condenser.vacuum_system.config, high_pressure_alarm=0.003, high_pressure_trip=0.0035,
                                     low_motive_pressure_alarm=5.0),
    "fw_low_suction": lambda sp: _set(sp.feedwater_system.protection_system.config, low_suction_pressure_trip=5.0),
    "fw_high_discharge": lambda sp: _set(sp.feedwater_system.protection_system.config, high_discharge_pressure_trip=2.0),
    "fw_low_flow": lambda sp: _set(sp.feedwater_system.protection_system.config, low_flow_trip=5000.0),
    # the NPSH protection object is shared by the four pumps and keeps what the LAST pump left in it
    "fw_npsh_alarm": lambda sp: _set(list(sp.feedwater_system.pump_system.pumps.values())[3].state, npsh_available=6.0),
    "fw_npsh_trip": lambda sp: (_set(list(sp.feedwater_system.pump_system.pumps.values())[3].state, npsh_available=6.0),
                                _set(sp.feedwater_system.protection_system.config, low_suction_pressure_trip=8.0)),
    # cooling-water velocity above the tube vibration-damage threshold; steam demand split evenly instead of by primary flow
    "cond_tube_vibration": lambda sp: _set(sp.condenser.tube_degradation.config, vibration_damage_threshold=0.5),
    "sg_no_load_balancing": lambda sp: _set(sp.steam_generator_system.config, auto_load_balancing=False),
    # feedwater level control in manual mode (total flow demand taken as given); rotor held below 100 rpm by its speed limit
    # (thermal-bow accumulation of a rotor at rest)
    "fw_manual_flow": lambda sp: _set(sp.feedwater_system.config, auto_level_control=False),
    "rotor_slow": lambda sp: _set(sp.turbine.rotor_dynamics.config, max_speed=90.0),
    # condenser pressure outside the lead ejector's working range: no capacity, vacuum decays
    "ejector_out_of_range": lambda sp: [_set(e.config, min_suction_pressure=0.05) for e in
                                        (sp.condenser.vacuum_system.ejectors.values() if isinstance(sp.condenser.vacuum_system.ejectors, dict)
                                         else sp.condenser.vacuum_system.ejectors)],
    # lag ejector started by the pressure rule, then lead / lag rotation after 20 s
    "vacuum_lag_rotation": lambda sp: _set(sp.condenser.vacuum_system.config, auto_start_pressure=0.003, auto_stop_pressure=0.002,
                                           rotation_interval=20.0 / 3600.0),
}


def cfg8(only=None):
    """One fixture per protection path (tests/golden/trip_<case>.npz): 60 steps at dt = 1 s."""
    for name, tweak in TRIP_CASES.items():
        if only and name not in only:
            continue
        plants = _plants(["oil_top_off"], dt=1.0, heat_source="constant", noise_enabled=False)
        tweak(plants[0].sim.secondary_physics)
        run_scenario("trip_" + name, plants, 60, [1, 2, 3, 4, 5, 6, 8, 12, 20, 40, 60], lambda p, t, sim: NO)


def cfg9():
    """Pump-level trips, pump start / stop dynamics, the un-frozen sensor path and pH-controller modes that no other
    fixture reaches (found with gcov on the host build of the restatement): one forced condition per plant, written into
    the reference's own objects before the first step."""
    plants = _plants(["oil_top_off"] * 19, dt=1.0, heat_source="constant", noise_enabled=False)

    def pumps(i):
        return list(plants[i].sim.secondary_physics.feedwater_system.pump_system.pumps.values())
    # 0: suction pressure (frozen at its IC value) below the 0.2 MPa pump trip
    pumps(0)[0].state.suction_pressure = 0.15
    # 1 / 2: oil reservoir nearly empty / overfilled
    pumps(1)[1].lubrication_system.oil_level = 3.0
    pumps(2)[2].lubrication_system.oil_level = 106.0
    # 3: seal leakage above its trip
    pumps(3)[0].lubrication_system.seal_leakage_rate = 12.0
    # 4: every component moderately worn: no single wear trip, the combined-wear trip
    for k in pumps(4)[1].lubrication_system.component_wear:
        pumps(4)[1].lubrication_system.component_wear[k] = 7.5
    # 5: mild NPSH deficit on a running pump: cavitation without a trip (intensity, damage, noise, vibration)
    pumps(5)[2].state.npsh_available = 13.0
    # 6: NPSH below the critical 4 m
    pumps(6)[0].state.npsh_available = 3.0
    # 7: the standby pump started, a running pump stopped (STARTING / STOPPING ramps)
    pumps(7)[3].start_pump()
    pumps(7)[0].stop_pump()
    # 8: no pump carries the initial-conditions flag: suction pressure and NPSH follow the system every step
    for q in pumps(8):
        if hasattr(q, "_initial_conditions_applied"):
            del q._initial_conditions_applied
    # 9: hot pump motor and oil (equipment-protection timers and alarms)
    pumps(9)[1].lubrication_system.oil_temperature = 118.0
    # 10 / 11 / 12: pH controller in manual mode / low chemical tanks / measured pH above the trip band
    ph = lambda i: plants[i].sim.secondary_physics.ph_control_system.controller
    st10 = ph(10).state
    st10.control_mode = type(st10.control_mode)("MANUAL") if not hasattr(type(st10.control_mode), "MANUAL") else type(st10.control_mode).MANUAL
    st10.manual_output = 35.0
    ph(11).state.ammonia_tank_level, ph(11).state.morpholine_tank_level = 12.0, 15.0
    plants[12].sim.secondary_physics.water_chemistry.ph = 10.4
    # 13: feedwater diagnostics health collapsed (system-level diagnostic trip)
    plants[13].sim.secondary_physics.feedwater_system.diagnostics.overall_health_score = 0.2
    # 14: rotor almost at rest (thermal-bow accumulation below 100 rpm)
    plants[14].sim.secondary_physics.turbine.rotor_dynamics.rotor_speed = 40.0
    # 15: a running pump far below its speed setpoint (rate-limited speed ramp)
    pumps(15)[1].state.speed_percent = 55.0
    # 16: a NaN written into the fuel temperature before step 3: the silent reset of thermal_hydraulics.py:257-269
    def inject(p, t):
        return ("pri.fuel_temperature", float("nan")) if (p == 16 and t == 3) else None
    # 17: every pump out of oil: all four trip, total loss of feedwater (no pump power, steam generators boiling down,
    # electrical output gated to zero)
    for q in pumps(17):
        q.lubrication_system.oil_level = 3.0
    # 18: NPSH just above the required value on a running pump: cavitation strong enough to accumulate damage, no trip
    pumps(18)[1].state.npsh_available = 12.05
    run_scenario("cfg9_pump_trips_modes", plants, 90, [1, 2, 3, 4, 5, 6, 8, 12, 20, 30, 45, 60, 90], lambda p, t, sim: NO,
                 inject=inject)


def cfg10():
    """Primary side alone (NuclearPlantSimulator(enable_secondary=False)): reactor heat source under load-following rods and a
    mixed action sequence, constant heat source with setpoint steps; the secondary half of the observation is zero."""
    plants = [R.make_reference_plant(None, dt=1.0, heat_source="reactor", enable_secondary=False),
              R.make_reference_plant(None, dt=1.0, heat_source="reactor", enable_secondary=False)]

    def policy(p, t, sim):
        if p == 0:
            s = np.sin(t / 60.0)
            return (1 if s > 0.5 else (0 if s < -0.5 else 8)), 0.7, np.nan
        seq = [10, 8, 9, 8, 2, 3, 4, 5, 6, 7, 0, 1]
        return seq[(t // 20) % len(seq)], 0.5, np.nan
    # strict=False: a plant without secondary side has no secondary members; those state fields and parameters are stored as 0
    run_scenario("cfg10_primary_only", plants, 600, [1, 2, 10, 100, 300, 600], policy, strict=False)
    plants = [R.make_reference_plant(None, dt=1.0, heat_source="constant", noise_enabled=True, noise_std_percent=0.5,
                                     enable_secondary=False)]
    run_scenario("cfg10_primary_only_constant", plants, 600, [1, 2, 10, 100, 300, 600],
                 lambda p, t, sim: (8, 1.0, (80.0 if t == 100 else (95.0 if t == 300 else np.nan))), strict=False)


ALL = {"cfg10": cfg10, "cfg9": cfg9, "cfg8": cfg8, "cfg1": cfg1, "cfg2": cfg2, "cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5, "cfg6": cfg6, "cfg7": cfg7}

if __name__ == "__main__":
    if not R.reference_available():
        sys.exit("reference not found at " + R.REF_ROOT)
    sel = sys.argv[1:] or list(ALL)
    for k in sel:
        ALL[k]()
