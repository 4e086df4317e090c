"""Maintenance: action effects (host oracle vs live-reference fixtures), threshold binding, orchestrator decisions.

tests/golden/maint_effects.npz and maint_<scenario>.npz come from the unmodified reference (oracle/make_golden_maint.py).
"""
import ctypes
import json
import os

import numpy as np
import pytest

from tests import _util as U


def _maint():
    from nuclear_sim_b200 import maintenance
    return maintenance


def test_action_names_match_library():
    from nuclear_sim_b200 import _clib
    M = _maint()
    L = _clib.lib()
    n = L.nps_n_maintenance_actions()
    assert n == len(M.ACTION_NAMES)
    assert tuple(L.nps_maintenance_action_name(i).decode() for i in range(n)) == M.ACTION_NAMES


def _load_effects():
    z = np.load(os.path.join(U.GOLDEN, "maint_effects.npz"), allow_pickle=False)
    from nuclear_sim_b200 import field_names
    assert tuple(str(s) for s in z["state_names"]) == field_names("PlantState")
    return z


def test_effects_match_reference(oracle_lib):
    """component.perform_maintenance state mutations, one (component, action) pair at a time."""
    M = _maint()
    z = _load_effects()
    params = np.ascontiguousarray(z["params"])
    oracle_lib.nps_oracle_apply_maintenance.restype = ctypes.c_int
    for i in range(len(z["component"])):
        cid, action, sub = str(z["component"][i]), str(z["action"][i]), str(z["arg"][i])
        st = np.ascontiguousarray(z["before"][i]).copy()
        arg = M.BEARING_ARG.get(sub, 0) if action == "bearing_replacement" else 0
        rc = oracle_lib.nps_oracle_apply_maintenance(U.ptr(st), U.ptr(params), M.target_code(cid), M.action_code(action), arg)
        assert rc in (0, 1), f"{cid} {action}: unsupported target"
        assert bool(rc) == bool(z["success"][i]), f"{cid} {action}: success flag {rc} vs reference {z['success'][i]}"
        U.assert_states_close(st[None, :], z["after"][i][None, :], 1e-13, f"{cid} {action} {sub}")


def test_apply_maintenance_error_conventions(oracle_lib):
    """A bearing selector outside 0..3 makes bearing_replacement fail without touching the state (the reference's
    'Unknown bearing component' branch, pump_lubrication.py:756-805); a target id without a restated perform_maintenance
    reports status 2 and changes nothing."""
    M = _maint()
    z = _load_effects()
    params = np.ascontiguousarray(z["params"])
    oracle_lib.nps_oracle_apply_maintenance.restype = ctypes.c_int
    st0 = np.ascontiguousarray(z["before"][0])
    st = st0.copy()
    assert oracle_lib.nps_oracle_apply_maintenance(U.ptr(st), U.ptr(params), M.target_code("FWP-1"), M.action_code("bearing_replacement"), 7) == 0
    rep = [i for i, n in enumerate(z["state_names"]) if str(n).startswith("rep.")]
    keep = np.setdiff1d(np.arange(len(st)), rep)          # the report columns are refreshed after every request
    assert np.array_equal(st[keep], st0[keep], equal_nan=True)
    st = st0.copy()
    assert oracle_lib.nps_oracle_apply_maintenance(U.ptr(st), U.ptr(params), 99, M.action_code("oil_change"), 0) == 2
    assert np.array_equal(st[keep], st0[keep], equal_nan=True)
    with pytest.raises(KeyError):
        M.target_code("SECONDARY-COMP-001-FW")


def _load_sweep():
    z = np.load(os.path.join(U.GOLDEN, "maint_effects_sweep.npz"), allow_pickle=False)
    from nuclear_sim_b200 import field_names
    assert tuple(str(s) for s in z["state_names"]) == field_names("PlantState")
    return z


def sweep_cases(z):
    """(target id, action, reference success flag, reference state after the call) in the order the reference ran them;
    the calls of one segment (a target, one pass over the action list) act on one evolving plant state that starts
    from z['base'][segment]."""
    cur, last = None, -1
    for i in range(len(z["action"])):
        t = int(z["target"][i])
        if t != last:
            cur, last = z["base"][t].copy(), t
        lo, hi = int(z["start"][i]), int(z["start"][i + 1])
        after = cur.copy()
        after[z["index"][lo:hi]] = z["value"][lo:hi]
        yield str(z["targets"][t]), str(z["action"][i]), bool(z["success"][i]), cur, after
        cur = after


def test_catalogue_sweep_matches_reference(oracle_lib):
    """All 127 catalogued action names (+ 10 uncatalogued) x 11 components of the 9 classes a work order can address, in
    catalogue order and again in reverse order: success flag and state change of every call equal the live reference's
    (3 014 calls, oracle/make_golden_maint.py sweep)."""
    M = _maint()
    z = _load_sweep()
    assert len(z["action"]) == 11 * 2 * 137 and len(set(str(t) for t in z["targets"])) == 11
    params = np.ascontiguousarray(z["params"])
    oracle_lib.nps_oracle_apply_maintenance.restype = ctypes.c_int
    n_changed = 0
    for cid, action, ok, before, after in sweep_cases(z):
        st = np.ascontiguousarray(before).copy()
        rc = oracle_lib.nps_oracle_apply_maintenance(U.ptr(st), U.ptr(params), M.target_code(cid), M.action_code(action), 0)
        assert rc in (0, 1), f"{cid} {action}: unsupported target"
        assert bool(rc) == ok, f"{cid} {action}: success flag {rc} vs reference {ok}"
        U.assert_states_close(st[None, :], after[None, :], 1e-13, f"{cid} {action}")
        n_changed += int(not np.array_equal(before, after, equal_nan=True))
    assert n_changed >= 30


def _template_maintenance_config():
    z = np.load(os.path.join(U.GOLDEN, "maint_oil_top_off.npz"), allow_pickle=False)
    return json.loads(str(z["log"]))["maintenance_system"]


def test_threshold_binding_matches_reference_probe():
    """331 threshold rows, 90 of them bound to a logged column; the rest are inert exactly
    as in the reference, where _find_parameter_in_row_data finds no column."""
    M = _maint()
    tab = M.ThresholdTable(_template_maintenance_config())
    assert len(tab) == 331
    assert len(tab.bound()) == 90
    assert not tab.unsupported()
    by = {(r.component_id, r.parameter): r for r in tab.rows}
    assert by[("FWP-1", "oil_level")].field == "fw.pump[0].lub.oil_level"
    assert by[("FWP-3", "thrust_bearing_wear")].field == "fw.pump[2].lub.component_wear[3]"
    assert by[("FWP-3", "thrust_bearing_wear")].sub_component == "thrust_bearing"
    assert by[("FWP-2", "sum_wear_level")].derived == "pump_sum_wear"
    assert by[("SG-1", "tube_wall_temperature")].field == "sgs.sg[1].tube_wall_temp"
    assert by[("LP-6", "efficiency")].field == "turb.stage[13].actual_efficiency"
    assert by[("SECONDARY-COMP-001-TURB", "efficiency")].field == "turb.stage[13].actual_efficiency"   # last stage wins
    assert by[("FEE-001", "oil_level")].field is None and by[("FEE-001", "oil_level")].column is None
    assert by[("SECONDARY-COMP-001-COND", "fouling_resistance")].field == "cond.fl_total_fouling_resistance"
    # order of components = order the reference emits events in
    assert [r.component_id for r in tab.rows][0] == "FWP-1" and tab.rows[-1].component_id == "SECONDARY-COMP-001-COND"


def test_orchestrator_decisions():
    M = _maint()
    v = lambda p, val, a: {"parameter": p, "value": val, "action": a}
    assert M.orchestrate("FWP-1", [v("oil_level", 57.9, "oil_top_off")], "oil_top_off") == "oil_top_off"
    # promotion: oil_top_off -> oil_change when the contamination violation is in the batch
    assert M.orchestrate("FWP-1", [v("oil_level", 57.9, "oil_top_off"), v("oil_contamination_level", 15.5, "oil_change")],
                         "oil_top_off") == "oil_change"
    # two encompassed major actions -> component_overhaul
    assert M.orchestrate("FWP-2", [v("motor_bearing_wear", 9.0, "bearing_replacement"), v("seal_wear", 17.0, "seal_replacement")],
                         "bearing_replacement") == "component_overhaul"
    # coordination keeps the base action
    assert M.orchestrate("FWP-2", [v("motor_bearing_wear", 9.0, "bearing_replacement"), v("npsh_available", 17.0, "npsh_analysis")],
                         "bearing_replacement") == "bearing_replacement"
    assert M.orchestrate("SG-0", [v("tsp_fouling_fraction", 0.31, "tsp_chemical_cleaning")], "tsp_chemical_cleaning") == "tsp_chemical_cleaning"
    # 'HP-1' contains neither 'tb' nor 'turbine': unknown type, no hierarchy
    assert M.infer_component_type("HP-1") == "unknown"
    assert M.infer_component_type("SECONDARY-COMP-001-COND") == "unknown"     # neither 'cd' nor 'condenser' occurs in the id
    assert M.infer_component_type("SECONDARY-COMP-001-SG") == "steam_generator"


def test_orchestrator_sweep_matches_reference():
    """2 700 random sets of 1-5 simultaneous violations (every component of the template + two ids outside it), decided by
    the reference's own MaintenanceOrchestrator (oracle/make_golden_maint.py orchestrator): same selected action."""
    import gzip
    M = _maint()
    with gzip.open(os.path.join(U.GOLDEN, "orchestrator_sweep.json.gz"), "rt") as fh:
        cases = json.load(fh)["cases"]
    assert len(cases) >= 2000
    changed = 0
    for c in cases:
        got = M.orchestrate(c["component"], c["violations"], c["requested"])
        assert got == c["selected"], (c["component"], c["requested"], [(v["parameter"], v["value"]) for v in c["violations"]])
        changed += int(c["selected"] != c["requested"])
    assert changed >= 200          # promotions / comprehensive actions do occur in the sweep


SCENARIOS = ["oil_top_off", "tsp_chemical_cleaning", "oil_change", "scale_removal"]


def compare_logs(maint, log, plant=0):
    """Discrete events must match bit-exactly: which step, which component, which violations, which action."""
    ev_ref = log["events"]
    ev = [e for e in maint.event_log if e["plant"] == plant]
    assert [(e["t"], e["component"], e["action"]) for e in ev] == [(e["t"], e["component"], e["action"]) for e in ev_ref]
    for a, b in zip(ev, ev_ref):
        assert [(v["parameter"], v["action"], v["priority"], v["component_id"]) for v in a["violations"]] == \
               [(v[0], v[2], v[3], v[4]) for v in b["violations"]]
        for va, vb in zip(a["violations"], b["violations"]):
            assert abs(va["value"] - vb[1]) <= 1e-9 * max(1.0, abs(vb[1]))
    cr = [w for w in maint.created_log if w.plant == plant]
    assert [(w.created, w.work_order_id, w.component_id, w.action, w.priority, w.planned_start, w.sub_component) for w in cr] == \
           [(c["t"], c["wo"], c["component"], c["action"], c["priority"], c["planned_start"], c["extracted"]) for c in log["created"]]
    ex = [w for w in maint.executed_log if w.plant == plant]
    assert [(w.executed_at, w.work_order_id, w.component_id, w.action, w.success) for w in ex] == \
           [(e["t"], e["wo"], e["component"], e["action"], e["success"]) for e in log["executed"]]


@pytest.mark.parametrize("name", SCENARIOS)
def test_scenario_host_logic_on_oracle(name):
    """Threshold events, work-order issuance and execution times equal the live reference's, and the state after every
    step (including the steps where perform_maintenance mutated it) stays within the per-step tolerance."""
    M = _maint()
    g = np.load(os.path.join(U.GOLDEN, f"maint_{name}.npz"), allow_pickle=False)
    sim = U.OracleSim(g["state0"], g["params"])
    worst = [0.0]

    def check_state(t, got):
        worst[0] = max(worst[0], U.assert_states_close(got, g["states"][t][None, :], U.TOL_STEP * (t + 1), f"{name} step {t}"))
    maint, log = U.replay_maintenance_scenario(
        sim, g, lambda s, cfg: M.BatchedAutoMaintenance(s, M.ThresholdTable(cfg), aggressive=True), check_state)
    compare_logs(maint, log)
    assert len(log["created"]) >= 1 and len(log["executed"]) >= 1


@pytest.mark.parametrize("name", SCENARIOS)
@pytest.mark.parametrize("max_k", [64, 2])
def test_advance_with_fused_launches_equals_single_steps(name, max_k):
    """BatchedAutoMaintenance.advance(): thresholds evaluated after every substep inside the launch, launches cut at the
    15-minute gate steps.  Events, work orders (created / executed times) and the final state must equal the K = 1
    replay and the live-reference log, whatever the launch length."""
    M = _maint()
    g = np.load(os.path.join(U.GOLDEN, f"maint_{name}.npz"), allow_pickle=False)
    log = json.loads(str(g["log"]))
    T = g["states"].shape[0]
    sim = U.OracleSim(g["state0"], g["params"])
    maint = M.BatchedAutoMaintenance(sim, M.ThresholdTable(log["maintenance_system"]), aggressive=True)
    noise = np.ascontiguousarray((g["noise"][:, None, :] if g["noise"].ndim == 2 else g["noise"]).transpose(0, 2, 1))  # [T, 5, P]
    half = T // 2
    maint.advance(half, noise=noise[:half], max_k=max_k)
    maint.advance(T - half, noise=noise[half:], max_k=max_k)
    compare_logs(maint, log)
    U.assert_states_close(sim.state_numpy(), g["states"][T - 1][None, :], U.TOL_STEP * T, f"{name} final state")


@pytest.mark.parametrize("cls", ["ColumnarAutoMaintenance", "NativeAutoMaintenance"])
@pytest.mark.parametrize("name", SCENARIOS)
def test_columnar_bookkeeping_on_reference_scenarios(name, cls):
    """ColumnarAutoMaintenance (numpy columns) and NativeAutoMaintenance (the library's work-order table) through
    advance(): same events, orders and state as the live reference."""
    M = _maint()
    g = np.load(os.path.join(U.GOLDEN, f"maint_{name}.npz"), allow_pickle=False)
    log = json.loads(str(g["log"]))
    T = g["states"].shape[0]
    sim = U.OracleSim(g["state0"], g["params"])
    maint = getattr(M, cls)(sim, M.ThresholdTable(log["maintenance_system"]), aggressive=True)
    noise = np.ascontiguousarray((g["noise"][:, None, :] if g["noise"].ndim == 2 else g["noise"]).transpose(0, 2, 1))
    maint.advance(T, noise=noise)
    maint.materialize_logs()
    compare_logs(maint, log)
    U.assert_states_close(sim.state_numpy(), g["states"][T - 1][None, :], U.TOL_STEP * T, f"{name} final state")


@pytest.mark.parametrize("cls", ["ColumnarAutoMaintenance", "NativeAutoMaintenance"])
@pytest.mark.parametrize("head_quirks", [True, False])
def test_columnar_equals_object_bookkeeping_on_a_busy_batch(head_quirks, cls):
    """192 plants with oil levels, contamination and bearing wear staggered around their thresholds (several components
    and several violations per component fire in the same check): the columnar and the object implementation create and
    execute the same work orders at the same times and leave the same state."""
    M = _maint()
    from nuclear_sim_b200 import field_index
    g = np.load(os.path.join(U.GOLDEN, "maint_oil_top_off.npz"), allow_pickle=False)
    cfg = json.loads(str(g["log"]))["maintenance_system"]
    ix = field_index()
    n = 192
    st = np.tile(g["state0"], (n, 1))
    rng = np.random.RandomState(11)
    st[:, ix["fw.pump[0].lub.oil_level"]] = 58.0 + rng.uniform(-0.3, 1.0, n)
    st[:, ix["fw.pump[2].lub.oil_level"]] = 58.0 + rng.uniform(0.0, 2.0, n)
    st[:, ix["fw.pump[1].lub.oil_contamination_level"]] = 15.2 - rng.uniform(-0.01, 0.02, n)
    st[:, ix["fw.pump[0].lub.oil_contamination_level"]] = 15.2 - rng.uniform(-0.05, 0.02, n)
    st[::3, ix["fw.pump[3].lub.component_wear[1]"]] = 8.6          # motor bearing wear above 8.5: bearing_replacement
    st[::5, ix["fw.pump[3].lub.component_wear[4]"]] = 16.5         # seal wear above 16: seal_replacement (-> overhaul with the above)
    sims = [U.OracleSim(st, g["params"]), U.OracleSim(st, g["params"])]
    maints = [M.BatchedAutoMaintenance(sims[0], M.ThresholdTable(cfg), aggressive=True, head_quirks=head_quirks),
              getattr(M, cls)(sims[1], M.ThresholdTable(cfg), aggressive=True, head_quirks=head_quirks)]
    for m in maints:
        m.advance(24)
    maints[1].materialize_logs()
    key = lambda w: (w.created, w.plant, w.component_id, w.action, w.work_order_id, w.priority, w.planned_start, w.sub_component,  # noqa: E731
                     w.executed_at, w.success)
    a, b = sorted(key(w) for w in maints[0].created_log), sorted(key(w) for w in maints[1].created_log)
    assert a == b and len(a) > 200
    assert len({k[3] for k in a}) >= 4                     # several distinct actions, promotions included
    ex = lambda m: sorted((w.executed_at, w.plant, w.work_order_id, w.action, w.success) for w in m.executed_log)   # noqa: E731
    assert ex(maints[0]) == ex(maints[1]) and len(maints[1].executed_log) > 100
    assert [(e["t"], e["plant"], e["component"], e["action"]) for e in maints[0].event_log] == \
           [(e["t"], e["plant"], e["component"], e["action"]) for e in maints[1].event_log]
    np.testing.assert_array_equal(sims[0].state_numpy(), sims[1].state_numpy())


@pytest.mark.parametrize("cls", ["BatchedAutoMaintenance", "ColumnarAutoMaintenance", "NativeAutoMaintenance"])
def test_bookkeeping_state_dict_round_trip(cls):
    """A run cut in the middle (work orders pending), its books moved to a fresh bookkeeping object through state_dict /
    pickle, continues to the same work orders and state as the uninterrupted run — for both bookkeeping classes."""
    import pickle
    M = _maint()
    from nuclear_sim_b200 import field_index
    g = np.load(os.path.join(U.GOLDEN, "maint_oil_top_off.npz"), allow_pickle=False)
    cfg = json.loads(str(g["log"]))["maintenance_system"]
    ix = field_index()
    n = 24
    st = np.tile(g["state0"], (n, 1))
    st[:, ix["fw.pump[0].lub.oil_level"]] = 58.0 + np.linspace(-0.2, 0.8, n)
    st[::4, ix["fw.pump[1].lub.oil_contamination_level"]] = 15.21
    mk = lambda sim: getattr(M, cls)(sim, M.ThresholdTable(cfg), aggressive=True)   # noqa: E731
    full_sim, part_sim = U.OracleSim(st, g["params"]), U.OracleSim(st, g["params"])
    full, part = mk(full_sim), mk(part_sim)
    full.advance(20)
    part.advance(7)
    resumed_sim = U.OracleSim(part_sim.state_numpy(), g["params"])
    resumed_sim.step_index = part_sim.step_index
    resumed_sim._last = None
    resumed = mk(resumed_sim)
    resumed_sim._last = part_sim._last.copy()                      # threshold cooldown stamps travel with the simulator
    resumed.load_state_dict(pickle.loads(pickle.dumps(part.state_dict())))
    resumed.advance(13)
    for m in (full, resumed):
        if hasattr(m, "materialize_logs"):
            m.materialize_logs()
    key = lambda w: (w.created, w.plant, w.component_id, w.action, w.work_order_id, w.executed_at, w.success)   # noqa: E731
    assert sorted(key(w) for w in resumed.created_log) == sorted(key(w) for w in full.created_log) and len(full.created_log) > 10
    np.testing.assert_array_equal(resumed_sim.state_numpy(), full_sim.state_numpy())
    with pytest.raises(ValueError):
        other = "ColumnarAutoMaintenance" if cls == "BatchedAutoMaintenance" else "BatchedAutoMaintenance"
        getattr(M, other)(U.OracleSim(st, g["params"]), M.ThresholdTable(cfg)).load_state_dict(part.state_dict())


def test_bookkeeping_reset_of_some_plants_agrees_across_classes():
    """Episode reset in the middle of a run: the reset plants lose their pending orders, dedupe stamps and numbering (they
    start again at WO-000001) in all three bookkeeping classes alike; the other plants are untouched."""
    M = _maint()
    from nuclear_sim_b200 import field_index
    g = np.load(os.path.join(U.GOLDEN, "maint_oil_top_off.npz"), allow_pickle=False)
    cfg = json.loads(str(g["log"]))["maintenance_system"]
    ix = field_index()
    n = 16
    st = np.tile(g["state0"], (n, 1))
    st[:, ix["fw.pump[0].lub.oil_level"]] = 58.0 + np.linspace(-0.2, 0.6, n)
    st[::2, ix["fw.pump[1].lub.oil_contamination_level"]] = 15.21
    logs = []
    for cls in ("BatchedAutoMaintenance", "ColumnarAutoMaintenance", "NativeAutoMaintenance"):
        sim = U.OracleSim(st, g["params"])
        m = getattr(M, cls)(sim, M.ThresholdTable(cfg), aggressive=True)
        m.advance(5)
        for p in (3, 6, 9):
            sim.reset_plant(p)
            sim._last[:, p] = -np.inf
        m.reset([3, 6, 9])
        m.advance(9)
        if hasattr(m, "materialize_logs"):
            m.materialize_logs()
        logs.append(sorted((w.created, w.plant, w.component_id, w.action, w.work_order_id,
                            -1.0 if w.executed_at is None else w.executed_at, bool(w.success)) for w in m.created_log))
    assert logs[0] == logs[1] == logs[2] and len(logs[0]) > 20
    again = [k for k in logs[0] if k[1] in (3, 6, 9) and k[4] == "WO-000001"]
    assert len(again) >= 2          # numbering restarted for reset plants (their first order before and after the reset)


@pytest.mark.parametrize("cls", ["ColumnarAutoMaintenance", "NativeAutoMaintenance"])
def test_interleaved_parts_equal_one_batch(cls):
    """advance_interleaved on two halves of a batch (own simulator and books each, resumed in turn launch by launch) creates
    and executes the same work orders at the same times and ends in the same states as advance() on the whole batch."""
    M = _maint()
    from nuclear_sim_b200 import field_index
    g = np.load(os.path.join(U.GOLDEN, "maint_oil_top_off.npz"), allow_pickle=False)
    cfg = json.loads(str(g["log"]))["maintenance_system"]
    ix = field_index()
    n = 32
    st = np.tile(g["state0"], (n, 1))
    rng = np.random.RandomState(3)
    st[:, ix["fw.pump[0].lub.oil_level"]] = 58.0 + rng.uniform(-0.2, 0.9, n)
    st[:, ix["fw.pump[1].lub.oil_contamination_level"]] = 15.2 - rng.uniform(-0.02, 0.03, n)
    whole_sim = U.OracleSim(st, g["params"])
    whole = getattr(M, cls)(whole_sim, M.ThresholdTable(cfg), aggressive=True)
    whole.advance(21)
    parts_sim = [U.OracleSim(st[:16], g["params"]), U.OracleSim(st[16:], g["params"])]
    parts = [getattr(M, cls)(q, M.ThresholdTable(cfg), aggressive=True) for q in parts_sim]
    M.advance_interleaved(parts, 21)
    key = lambda w, off: (w.created, w.plant + off, w.component_id, w.action, w.work_order_id,  # noqa: E731
                          -1.0 if w.executed_at is None else w.executed_at, bool(w.success))
    for m in [whole] + parts:
        m.materialize_logs()
    a = sorted(key(w, 0) for w in whole.created_log)
    b = sorted([key(w, 0) for w in parts[0].created_log] + [key(w, 16) for w in parts[1].created_log])
    assert a == b and len(a) > 30
    np.testing.assert_array_equal(np.concatenate([q.state_numpy() for q in parts_sim]), whole_sim.state_numpy())


def test_native_work_order_table_c_api():
    """nps_wo_* directly through the C ABI (no simulator, no device): grouping with the rule table, the three issuance
    rules, due selection with and without the one-per-plant quirk, capacity report, bad arguments, export / import."""
    import ctypes
    from nuclear_sim_b200 import _clib
    L = _clib.lib()
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)   # noqa: E731
    i64 = lambda *v: np.array(v, dtype=np.int64)      # noqa: E731
    # 3 rows: rows 0, 1 -> component 0, row 2 -> component 1; row 0 promotes to action 5 above 10.0
    row_comp, fallback, row_action, row_prio, row_sub = i64(0, 0, 1), i64(1, 2, 3), i64(1, 2, 3), i64(1, 3, 0), i64(0, 0, 2)
    rule_thr, rule_act = np.array([[10.0], [np.inf], [np.inf]]), i64(5, 0, 0).reshape(3, 1)
    delay = np.array([60.0, 0.0, 0.0, 0.0, 0.0])

    def make(head):
        h = ctypes.c_void_p()
        assert L.nps_wo_create(4, 2, 3, 1, P(row_comp), P(rule_thr), P(rule_act), P(fallback), P(row_action), P(row_prio), P(row_sub),
                               P(delay), 24.0, head, ctypes.byref(h)) == 0
        return h
    h = make(1)
    plant, row, value = i64(0, 0, 1, 3), i64(0, 1, 0, 2), np.array([11.0, 1.0, 9.0, 7.0])
    g = [np.zeros(4, dtype=np.int64) for _ in range(7)]
    nm = ctypes.c_int64(0)
    ng = L.nps_wo_group(h, 4, P(plant), P(row), P(value), *[P(a) for a in g], ctypes.byref(nm))
    assert ng == 3 and nm.value == 1
    assert g[0][:3].tolist() == [0, 2, 3] and g[1][:3].tolist() == [2, 1, 1] and g[2][:3].tolist() == [0, 1, 3]
    assert g[4][:3].tolist() == [5, 1, 3]                    # promoted (11 > 10), fallback (9 < 10), row 2's own action
    assert g[6][:3].tolist() == [0, 0, 2]                    # sub-component only when the row's own action was selected
    assert L.nps_wo_group(h, 1, P(i64(9)), P(i64(0)), P(np.array([1.0])), *[P(a) for a in g], ctypes.byref(nm)) == -1
    ok = np.ones(8, dtype=np.uint8)
    og, os_ = np.zeros(4, dtype=np.int64), np.zeros(4, dtype=np.int64)
    issue = lambda t: L.nps_wo_issue(h, t, 3, P(g[2]), P(g[3]), P(g[4]), P(g[5]), P(g[6]), P(ok), 8, P(og), P(os_))   # noqa: E731
    assert issue(5.0) == 3 and os_[:3].tolist() == [1, 1, 1]
    assert issue(10.0) == 0                                   # active order / stamp younger than the window
    cols = [np.zeros(8, dtype=np.int64) for _ in range(6)] + [np.zeros(8), np.zeros(8)]
    due = lambda t, cap: L.nps_wo_due(h, t, cap, *[P(a) for a in cols])   # noqa: E731
    assert due(5.0, 8) == 2 and cols[0][:2].tolist() == [0, 1]            # plant 3's order has priority 0: 60 minutes delay
    assert L.nps_wo_complete(h) == 0 and L.nps_wo_n_pending(h) == 1
    assert issue(12.0) == 0 and issue(30.0) == 2              # stamps: 12 - 5 < 24 blocks, 30 - 5 >= 24 does not; plant 3 still active
    assert due(65.0, 1) == 3                                  # capacity too small: the needed size, nothing handed out
    assert due(65.0, 8) == 3 and sorted(cols[0][:3].tolist()) == [0, 1, 3]
    n_p, n_s = ctypes.c_int64(0), ctypes.c_int64(0)
    L.nps_wo_sizes(h, ctypes.byref(n_p), ctypes.byref(n_s))
    pc, pt = np.zeros((6, n_p.value), dtype=np.int64), np.zeros((2, n_p.value))
    sk, st_, nc = np.zeros((2, n_s.value), dtype=np.int64), np.zeros(n_s.value), np.zeros(4, dtype=np.int64)
    assert L.nps_wo_export(h, P(pc), P(pt), P(sk), P(st_), P(nc)) == 0 and nc.tolist() == [2, 2, 0, 1]
    h2 = make(0)
    assert L.nps_wo_import(h2, pc.shape[1], P(pc), P(pt), sk.shape[1], P(sk), P(st_), P(nc)) == 0
    assert L.nps_wo_n_pending(h2) == 3
    cols2 = [np.zeros(8, dtype=np.int64) for _ in range(6)] + [np.zeros(8), np.zeros(8)]
    assert L.nps_wo_due(h2, 65.0, 8, *[P(a) for a in cols2]) == 3 and cols2[0][:3].tolist() == sorted(cols2[0][:3].tolist())
    L.nps_wo_destroy(h)
    L.nps_wo_destroy(h2)


def test_single_violation_fast_path_equals_orchestrate():
    """BatchedAutoMaintenance precomputes the decision for one-violation events; it must agree with orchestrate()
    for every threshold row of the reference configuration, below and above every rule threshold."""
    M = _maint()
    tab = M.ThresholdTable(_template_maintenance_config())
    for r in tab.rows:
        rules, fallback = M.single_violation_rules(r.component_id, r.parameter, r.action)
        probes = [0.0, 1e9] + [t * f for t, _ in rules for f in (0.999, 1.001)]
        for val in probes:
            want = M.orchestrate(r.component_id, [{"parameter": r.parameter, "value": val, "action": r.action}], r.action)
            got = fallback
            for thr, promoted in rules:
                if val > thr:
                    got = promoted
                    break
            assert got == want, (r.component_id, r.parameter, val)


def test_batched_timing_sweep_and_optimizer():
    """TimingOptimizer as one batched sweep (CPU stand-in engine): the committed initial oil level reproduces the
    reference's trigger time (140 min, maint_oil_top_off fixture), trigger time grows with the initial level, and the
    optimiser lands a 2.0 h target within the timestep."""
    from nuclear_sim_b200 import optimize as O
    g = np.load(os.path.join(U.GOLDEN, "maint_oil_top_off.npz"), allow_pickle=False)
    log = json.loads(str(g["log"]))
    cfg = log["maintenance_system"]
    fac = lambda st, p, dev: U.OracleSim(st, p)
    fld = "fw.pump[0].lub.oil_level"
    from nuclear_sim_b200 import field_index
    level0 = float(g["state0"][field_index()[fld]])
    vals = np.array([58.5, 59.0, level0, 61.0])
    hrs = O.trigger_time_sweep(g["state0"], g["params"], cfg, fld, vals, "oil_top_off", 5.0, component_id="FWP-1", engine_factory=fac)
    assert hrs[2] == log["created"][0]["t"] / 60.0
    assert np.all(np.diff(hrs) > 0)
    v, t, sweeps = O.optimize_for_target_timing(g["state0"], g["params"], cfg, fld, 58.1, 62.0, "oil_top_off", 2.0,
                                                tolerance_hours=5.0 / 60.0, component_id="FWP-1", n_candidates=24,
                                                engine_factory=fac)
    assert abs(t - 2.0) <= 5.0 / 60.0 and sweeps <= 2 and 58.1 < v < level0


def test_batch_optimize_timing_several_actions_in_one_batch():
    """ICOptimizer.batch_optimize_timing as batched sweeps (CPU stand-in engine): two actions with different fields and
    targets share one batch; each group answers exactly what a dedicated single-action sweep answers."""
    from nuclear_sim_b200 import optimize as O
    g = np.load(os.path.join(U.GOLDEN, "maint_oil_top_off.npz"), allow_pickle=False)
    cfg = json.loads(str(g["log"]))["maintenance_system"]
    fac = lambda st, p, dev: U.OracleSim(st, p)
    f1, f2 = "fw.pump[0].lub.oil_level", "fw.pump[1].lub.oil_level"
    v1, v2 = np.array([58.5, 59.5, 60.5]), np.array([59.0, 60.0])
    both = O.trigger_time_sweep_multi(g["state0"], g["params"], cfg, [(f1, v1, "oil_top_off", "FWP-1"),
                                                                       (f2, v2, "oil_top_off", "FWP-2")], 5.0, engine_factory=fac)
    solo1 = O.trigger_time_sweep(g["state0"], g["params"], cfg, f1, v1, "oil_top_off", 5.0, component_id="FWP-1", engine_factory=fac)
    solo2 = O.trigger_time_sweep(g["state0"], g["params"], cfg, f2, v2, "oil_top_off", 5.0, component_id="FWP-2", engine_factory=fac)
    np.testing.assert_array_equal(both[0], solo1)
    np.testing.assert_array_equal(both[1], solo2)
    res = O.batch_optimize_timing(g["state0"], g["params"], cfg,
                                  {"oil_top_off": (f1, 58.1, 62.0, 2.0, "FWP-1")}, tolerance_hours=5.0 / 60.0,
                                  n_candidates=24, engine_factory=fac)
    v, t, sweeps = res["oil_top_off"]
    assert abs(t - 2.0) <= 5.0 / 60.0 and sweeps <= 2
