"""The scalar NuclearPlantSimulator facade (reference API, sim.py:27-258) on the CPU stand-in engine: cfg1 replay
(noise stream + setpoint ramp), a maintenance scenario through the runner-facing members, and the CSV exports."""
import csv
import json
import os
import types

import numpy as np

from tests import _util as U


def _facade(state0, params, dt, heat_source=None, config=None, **kw):
    from nuclear_sim_b200.plant_simulator import NuclearPlantSimulator
    p = np.array(params, copy=True)
    from nuclear_sim_b200 import field_index
    p[field_index("PlantParams")["dt"]] = dt
    eng = U.OracleSim(state0, p)
    sim = NuclearPlantSimulator(dt=dt, heat_source=heat_source, secondary_config=config, initial_state=state0, params=p,
                                engine=eng, **kw)
    return sim, eng


def test_cfg1_replay_through_step_api():
    """BASELINE config #1 through the reference call sequence: set_power_setpoint(%) then step(NO_ACTION); the host
    draws RandomState(42) normals exactly like ConstantHeatSource.rng (constant_heat_source.py:60,178)."""
    g = U.load_golden("cfg1_oil_top_off")
    hs = types.SimpleNamespace(rated_power_mw=3000.0, noise_enabled=True, noise_std_percent=0.1, noise_seed=42,
                               noise_filter_time_constant=30.0)
    sim, eng = _facade(g["state0"][0], g["params"], 5.0, heat_source=hs, enable_state_management=False)
    # replay the fixture's own draws (one generator interleaved heat and pH draws when the fixture was made)
    T = g["actions"].shape[0]
    for t in range(T):
        z = g["noise"][t, 0]
        sim._heat_rng = types.SimpleNamespace(standard_normal=lambda z=z: float(z[0]))
        sim._ph_rng = types.SimpleNamespace(standard_normal=lambda z=z: float(z[1]),
                                            random_sample=lambda z=z, it=iter([2, 3, 4]): float(z[next(it)]))
        sim.primary_physics.heat_source.set_power_setpoint(float(g["setpoint"][t, 0]))
        out = sim.step(action=None)
        U.assert_states_close(eng.state_numpy(), g["states"][t], U.TOL_STEP * (t + 1), f"facade cfg1 step {t}")
        assert U.rel_err(out["observation"], g["obs"][t, 0]).max() <= 1e-9
        assert abs(out["reward"] - g["reward"][t, 0]) <= 1e-9 * max(1.0, abs(g["reward"][t, 0]))
        assert out["done"] is False
        assert out["info"]["time"] == 5.0 * (t + 1)
    assert abs(sim.state.power_level - g["power_level"][T - 1, 0]) <= 1e-9 * 100
    assert isinstance(sim.state.scram_status, bool)


def test_maintenance_scenario_through_facade(tmp_path):
    g = np.load(os.path.join(U.GOLDEN, "maint_oil_top_off.npz"), allow_pickle=False)
    log = json.loads(str(g["log"]))
    cfg = {"maintenance_system": log["maintenance_system"]}
    sim, eng = _facade(g["state0"], g["params"], 5.0, config=None)
    # what MaintenanceScenarioRunner._initialize_maintenance_monitoring does (maintenance_scenario_runner.py:286-299)
    sim.state_manager.config = cfg
    sim.maintenance_system.setup_monitoring_from_state_manager(sim.state_manager, aggressive_mode=True)
    assert sim.maintenance_system.check_interval_hours == 0.25
    draws = g["noise"]
    created_steps = []
    for t in range(g["states"].shape[0]):
        z = draws[t]
        sim._ph_rng = types.SimpleNamespace(standard_normal=lambda z=z: float(z[1]),
                                            random_sample=lambda z=z, it=iter([2, 3, 4]): float(z[next(it)]))
        sim.step(action=None)
        if sim.maintenance_system.current_update_work_orders:
            created_steps.append(t)
        U.assert_states_close(eng.state_numpy(), g["states"][t][None, :], U.TOL_STEP * (t + 1), f"facade maint step {t}")
    assert created_steps == [c["step"] for c in log["created"]]
    recent = sim.maintenance_system.get_recent_work_orders(limit=10)
    assert [(w["component_id"], w["maintenance_actions"][0]["action_type"], w["status"]) for w in recent] == \
           [(c["component"], c["action"], "completed") for c in log["created"]]
    assert "FWP-1" in sim.state_manager.get_current_threshold_violations()
    snap = sim.state_manager.get_component_state_snapshot("FWP-1")
    assert snap["oil_level"] > 90.0      # topped off at step 30
    # exports in the reference schema
    sec = tmp_path / "secondary.csv"; fwp = tmp_path / "fwp.csv"
    sim.state_manager.export_by_category("secondary", str(sec))
    sim.state_manager.export_by_subcategory("secondary", "feedwater_FWP-1", str(fwp))
    rows = list(csv.reader(open(fwp)))
    assert rows[0][0] == "time" and "secondary.feedwater_FWP-1.oil_level" in rows[0]
    assert len(rows) == 1 + g["states"].shape[0]
    j = rows[0].index("secondary.feedwater_FWP-1.oil_level")
    levels = [float(r[j]) for r in rows[1:]]
    assert levels[26] > 58.0 > levels[27] and levels[30] > 90.0
    hdr = next(csv.reader(open(sec)))
    assert len(hdr) > 600 and all(h == "time" or h.startswith("secondary.") for h in hdr)


def test_ring_buffer_export_equals_row_store_export(tmp_path):
    """export_ring_to_csv (the batched engine's ring buffer -> reference CSV) writes what TrajectoryStore writes for the
    same states: same header, same values, same timestamps - on a synthetic ring built from fixture states."""
    import datetime as dt
    from nuclear_sim_b200.export import ColumnSchema, TrajectoryStore, export_ring_to_csv
    g = U.load_golden("cfg1_oil_top_off")
    states = [g["states"][c][0] for c in range(5)]
    schema = ColumnSchema()
    which = schema.select("secondary.feedwater_FWP-1.")
    ids = schema.logged_fields(which)
    ring = np.stack([np.stack([np.stack([s[f], s[f] * 0 - 1.0]) for f in ids]) for s in states])     # [rows, n_logged, 2 plants]
    start = dt.datetime(2024, 1, 1)
    n = export_ring_to_csv(str(tmp_path / "ring.csv"), ring, ids, 0, start, 5.0, first_row_step=1, schema=schema, which=which)
    store = TrajectoryStore(start)
    for k, s in enumerate(states):
        store.add_row(start + dt.timedelta(minutes=5.0 * (k + 1)), s)
    store.export_by_subcategory("secondary", "feedwater_FWP-1", str(tmp_path / "store.csv"))
    assert n == 5
    assert open(tmp_path / "ring.csv").read() == open(tmp_path / "store.csv").read()
