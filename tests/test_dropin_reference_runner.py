"""CPU, needs /root/reference (skipped on the GPU box): the reference's OWN scenario driver
(data_gen/runners/maintenance_scenario_runner.py, unmodified) run twice on the same seeded inputs — once on the
reference simulator, once with ``NuclearPlantSimulator`` swapped for this repo's scalar facade — must report the same
work-order events at the same simulated times and the same plant trajectory.  The facade's engine here is the CPU
stand-in (tests/_util.OracleSim: no GPU in this container); the GPU tests cover the same facade logic on the CUDA engine
through the maintenance fixtures."""
import contextlib
import io
import warnings

import numpy as np
import pytest

from tests import _util as U

try:
    from oracle import refplant as R
    HAVE_REF = R.reference_available()
except Exception:   # pragma: no cover
    HAVE_REF = False

pytestmark = pytest.mark.skipif(not HAVE_REF, reason="live reference not present")


def _run(use_ours: bool, action: str, hours: float):
    R.setup_paths()
    import runners.maintenance_scenario_runner as msr
    import systems.secondary.ph_control_system as phmod
    cfg = R.compose_config(action, duration_hours=hours)
    R._clear_registries()
    if hasattr(phmod.np, "_real"):       # undo the harness' stream patch: the plain global np.random, as shipped
        phmod.np = phmod.np._real
    np.random.seed(123)                  # runner power profile + the reference's pH-controller draws
    orig_cls = msr.NuclearPlantSimulator
    if use_ours:
        from nuclear_sim_b200.plant_simulator import NuclearPlantSimulator as Ours

        def factory(heat_source=None, dt=1.0, enable_secondary=True, enable_state_management=True, secondary_config=None, **kw):
            # reference-side binding: the reference's constructors build the plant, two flat vectors
            # come out, the engine takes over the stepping
            ref = orig_cls(heat_source=heat_source, dt=dt, enable_secondary=True, enable_state_management=False,
                           secondary_config=secondary_config)
            s0, p0 = R.extract_state(ref), R.extract_params(ref)
            R._clear_registries()
            sim = Ours(dt=dt, heat_source=heat_source, initial_state=s0, params=p0, engine=U.OracleSim(s0, p0))
            sim._ph_rng = np.random                 # the reference's pH controller draws from the GLOBAL numpy stream
            sim._ph_rng.random_sample = np.random.random
            return sim
        msr.NuclearPlantSimulator = factory
    try:
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            runner = msr.MaintenanceScenarioRunner(cfg, verbose=False, enable_plotting=False)
            res = runner.run_scenario()
    finally:
        msr.NuclearPlantSimulator = orig_cls
        R._clear_registries()
    return runner, res


@pytest.mark.parametrize("action,hours", [("oil_top_off", 3.0), ("tsp_chemical_cleaning", 1.0)])
def test_reference_runner_on_facade_equals_reference_runner(action, hours):
    r_ref, res_ref = _run(False, action, hours)
    r_our, res_our = _run(True, action, hours)
    for k in ("success", "work_orders_created", "work_orders_executed", "maintenance_events", "simulation_data_points",
              "scenario_tracked_work_orders", "total_work_orders_before_filter"):
        assert res_ref[k] == res_our[k], k
    assert abs(res_ref["final_power_level"] - res_our["final_power_level"]) <= 1e-9 * 100
    ev = lambda r: [(e["event_type"], e.get("component_id"), e.get("time_hours"), e.get("work_order_id")) for e in r.work_order_events]
    assert ev(r_ref) == ev(r_our) and len(ev(r_ref)) >= 1
    me = lambda r: [(e["time_hours"], e["component_id"], e["action_type"], e["work_order_id"], e["success"]) for e in r.maintenance_events]
    assert me(r_ref) == me(r_our)
    # every exportable logged column equals the reference's DataFrame column, row by row
    df = r_ref.simulator.state_manager.data
    store = r_our.simulator.state_manager.store
    assert len(df) == len(store.rows)
    sc = store.schema
    which = sc.select()
    ours = [sc.row(row, which, ev) for row, ev in zip(store.rows, store.events)]
    bad = []
    for j, i in enumerate(which):
        mine = [r[j] for r in ours]
        if isinstance(mine[0], str):       # PumpStatus strings
            if list(df[sc.names[i]]) != mine:
                bad.append(sc.names[i])
            continue
        ref_col = df[sc.names[i]].to_numpy(dtype=float)
        if not np.all(np.abs(np.array(mine, dtype=float) - ref_col) <= 1e-9 * np.maximum(1e-6, np.abs(ref_col))):
            bad.append(sc.names[i])
    assert not bad, f"{len(bad)} exported columns differ from the reference, e.g. {bad[:6]}"
    assert len(which) == 788 and not sc.unavailable      # every column of the reference's state log
    for a, b in zip(r_ref.simulation_data, r_our.simulation_data):
        for k in ("time_minutes", "maintenance_events", "threshold_violations_count", "maintenance_history_count"):
            assert a[k] == b[k], k
        for k in ("target_power", "actual_power", "fuel_temperature", "coolant_temperature", "control_rod_position", "feedwater_flow"):
            assert abs(float(a[k]) - float(b[k])) <= 1e-9 * max(1.0, abs(float(a[k]))), k
