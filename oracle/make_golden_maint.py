"""TEST INFRASTRUCTURE — maintenance fixtures from the LIVE reference (/root/reference).

  python oracle/make_golden_maint.py effects     -> tests/golden/maint_effects.npz
  python oracle/make_golden_maint.py scenarios   -> tests/golden/maint_<scenario>.npz
  python oracle/make_golden_maint.py sweep       -> tests/golden/maint_effects_sweep.npz

effects:   before/after PlantState vectors around AutoMaintenanceSystem._perform_maintenance_action
           (systems/maintenance/auto_maintenance.py:582-673) for every (component, action) pair the engine restates,
           called exactly as _execute_work_order calls it, on plants with degraded initial conditions.
sweep:     EVERY action name of the reference catalogue (MaintenanceActionType, systems/maintenance/maintenance_actions.py:19-198,
           127 names) plus the uncatalogued names the template and the components use, against every component class a
           work order can address (pump, pump system, steam generator, SG system, HP / LP stage, turbine, condenser,
           turbine bearing lubrication, steam jet ejector):
           success flag and the sparse state change of each call, so that "which branch does this name take on this
           class" is pinned for the whole catalogue and not only for the pairs the engine restates.
scenarios: full NuclearPlantSimulator.step with state management and the automatic maintenance system, set up the
           way MaintenanceScenarioRunner does (data_gen/runners/maintenance_scenario_runner.py:205-330): per step the
           state after the step, the threshold events StateManager emitted, the work orders created and executed.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
sys.path.insert(0, _REPO)
from oracle import refplant as R  # noqa: E402

GOLDEN = os.path.join(_REPO, "tests", "golden")

# (component id, action, extracted_component_id or None)
EFFECT_CASES = [
    ("FWP-1", "oil_top_off", None), ("FWP-1", "oil_change", None), ("FWP-2", "bearing_replacement", None),
    ("FWP-1", "bearing_replacement", "pump_bearings"), ("FWP-3", "bearing_replacement", "thrust_bearing"),
    ("FWP-4", "bearing_replacement", "motor_bearings"), ("FWP-1", "seal_replacement", None),
    ("FWP-2", "system_cleaning", None), ("FWP-1", "bearing_inspection", None), ("FWP-1", "impeller_inspection", None),
    ("FWP-1", "impeller_replacement", None), ("FWP-1", "lubrication_system_check", None),
    ("FWP-1", "motor_inspection", None), ("FWP-1", "oil_analysis", None), ("FWP-1", "vibration_analysis", None),
    ("FWP-1", "cavitation_analysis", None), ("FWP-1", "npsh_analysis", None), ("FWP-1", "lubrication_inspection", None),
    ("FWP-3", "component_overhaul", None),
    ("SG-0", "tsp_chemical_cleaning", None), ("SG-1", "tsp_mechanical_cleaning", None), ("SG-0", "scale_removal", None),
    ("SG-2", "moisture_separator_maintenance", None), ("SG-1", "secondary_side_cleaning", None),
    ("SG-2", "routine_maintenance", None), ("SG-0", "tube_bundle_inspection", None), ("SG-0", "eddy_current_testing", None),
    ("SG-1", "tube_interior_scale_cleaning", None), ("SG-2", "tube_bundle_overhaul", None),
    ("HP-3", "efficiency_analysis", None), ("LP-2", "cleaning", None), ("HP-1", "blade_replacement", None),
    ("LP-6", "overhaul", None), ("SECONDARY-COMP-001-TURB", "efficiency_analysis", None),
    ("SECONDARY-COMP-001-COND", "condenser_tube_cleaning", None), ("SECONDARY-COMP-001-COND", "condenser_tube_plugging", None),
    ("SECONDARY-COMP-001-COND", "condenser_chemical_cleaning", None), ("SECONDARY-COMP-001-COND", "vacuum_system_test", None),
    ("SECONDARY-COMP-001-COND", "vacuum_leak_detection", None), ("SECONDARY-COMP-001-COND", "condenser_performance_test", None),
    # system-level targets: EnhancedTurbinePhysics, EnhancedSteamGeneratorPhysics, FeedwaterPumpSystem (no method at all)
    ("SECONDARY-COMP-001-TURB", "turbine_performance_test", None), ("SECONDARY-COMP-001-TURB", "turbine_system_optimization", None),
    ("SECONDARY-COMP-001-TURB", "turbine_protection_test", None), ("SECONDARY-COMP-001-TURB", "thermal_stress_analysis", None),
    ("SECONDARY-COMP-001-TURB", "vibration_analysis", None), ("SECONDARY-COMP-001-TURB", "routine_maintenance", None),
    ("SECONDARY-COMP-001-SG", "system_coordination_maintenance", None), ("SECONDARY-COMP-001-SG", "system_steam_quality_maintenance", None),
    ("SECONDARY-COMP-001-SG", "load_balancing_maintenance", None), ("SECONDARY-COMP-001-SG", "routine_maintenance", None),
    ("SECONDARY-COMP-001-SG", "tsp_chemical_cleaning", None),
    ("FEE-001", "oil_change", None), ("FEE-001", "routine_maintenance", None),
    # turbine bearing lubrication system and steam jet ejectors (degraded by hand just before the call: EFFECT_TWEAKS)
    ("TB-LUB-001", "turbine_oil_change", None), ("TB-LUB-001", "turbine_oil_top_off", None),
    ("TB-LUB-001", "oil_filter_replacement", None), ("TB-LUB-001", "oil_cooler_cleaning", None),
    ("TB-LUB-001", "lubrication_system_test", None), ("TB-LUB-001", "routine_maintenance", None),
    ("TB-LUB-001", "oil_change", None),
    ("SJE-001", "vacuum_ejector_cleaning", None), ("SJE-002", "vacuum_ejector_mechanical_cleaning", None),
    ("SJE-001", "vacuum_ejector_nozzle_replacement", None), ("SJE-002", "vacuum_ejector_inspection", None),
    ("SJE-001", "routine_maintenance", None), ("SJE-002", "oil_change", None),
    # the findings-dependent branches of the pump inspections and the lubrication check (worn pump, low oil: EFFECT_TWEAKS)
    ("FWP-4", "bearing_inspection", None), ("FWP-4", "impeller_inspection", None), ("FWP-4", "lubrication_system_check", None),
]


def _worn_pump(inst):
    L = inst.lubrication_system
    for k in L.component_wear:
        L.component_wear[k] = 12.5
    L.oil_level, L.oil_contamination_level = 81.0, 13.0


def _degrade_turbine_lub(inst):
    inst.oil_level, inst.oil_contamination_level, inst.oil_acidity_number = 62.0, 11.5, 0.42
    inst.oil_moisture_content, inst.oil_temperature, inst.lubrication_effectiveness = 0.06, 63.0, 0.71
    inst.oil_cooling_effectiveness = 0.68
    inst.component_wear["oil_coolers"] = 7.25


def _degrade_ejector(inst):
    inst.nozzle_fouling_factor, inst.diffuser_fouling_factor, inst.nozzle_erosion_factor = 0.62, 0.48, 0.86
    inst.overall_performance_factor = inst.nozzle_fouling_factor * inst.diffuser_fouling_factor * inst.nozzle_erosion_factor


EFFECT_TWEAKS = {"TB-LUB-001": _degrade_turbine_lub, "SJE-001": _degrade_ejector, "SJE-002": _degrade_ejector,
                 "FWP-4": _worn_pump}


def runner_style_plant(action, dt=5.0, noise=False):
    """Plant + maintenance monitoring configured as MaintenanceScenarioRunner does."""
    cfg = R.compose_config(action, duration_hours=24.0)
    rp = R.make_reference_plant(cfg, dt=dt, heat_source="constant", noise_enabled=noise, noise_std_percent=0.1,
                                enable_state_management=True)
    sim = rp.sim
    with R.quiet():
        sim.state_manager.config = cfg
        sim.maintenance_system.setup_monitoring_from_state_manager(sim.state_manager, aggressive_mode=True)
    return rp, cfg


def effects():
    L = R._layout()
    names = np.array(L.field_names())
    before, after, comp, act, arg, ok, params = [], [], [], [], [], [], None
    plants = {}
    for ic in ("oil_change", "tsp_chemical_cleaning"):
        plants[ic] = runner_style_plant(ic)[0]
    rng = np.random.RandomState(11)
    for i, (cid, action, sub) in enumerate(EFFECT_CASES):
        rp = plants["oil_change"] if (cid.startswith(("FWP", "TB-", "SJE")) or "COND" in cid or cid[:2] in ("HP", "LP") or "TURB" in cid or cid == "FEE-001") \
            else plants["tsp_chemical_cleaning"]
        sim = rp.sim
        for _ in range(2):
            z = np.array([0.0, rng.standard_normal(), rng.random_sample(), rng.random_sample(), rng.random_sample()])
            rp.step(8, 1.0, z)
        ms = sim.maintenance_system
        inst = sim.state_manager.get_registered_instance_info()[cid]["instance"]
        wo = types.SimpleNamespace(metadata={"extracted_component_id": sub} if sub else {})
        if (cid, action) == ("SECONDARY-COMP-001-TURB", "turbine_protection_test"):     # a latched trip for the test to reset
            pr = sim.secondary_physics.turbine.protection_system
            pr.trip_active, pr.trip_reasons = True, ["Low Vacuum"]
            pr.trip_timers["vibration"] = 1.5
        if cid in EFFECT_TWEAKS:
            EFFECT_TWEAKS[cid](inst)
        b = R.extract_state(sim)
        with R.quiet():
            res = ms._perform_maintenance_action(inst, action, wo)
        a = R.extract_state(sim)
        if params is None:
            params = R.extract_params(sim)
        before.append(b); after.append(a); comp.append(cid); act.append(action); arg.append(sub or ""); ok.append(bool(res.success))
        ch = np.nonzero(~((a == b) | (np.isnan(a) & np.isnan(b))))[0]
        print(f"[effects] {cid:24s} {action:32s} {sub or '':16s} success={res.success!s:5s} changed={len(ch):3d} "
              + ", ".join(names[j].split('.', 2)[-1] for j in ch[:6]))
    sn, pn = np.array(L.field_names("PlantState")), np.array(L.field_names("PlantParams"))
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, "maint_effects.npz"), before=np.array(before), after=np.array(after),
                        component=np.array(comp), action=np.array(act), arg=np.array(arg), success=np.array(ok),
                        params=params, state_names=sn, param_names=pn)


SWEEP_TARGETS = ["FWP-2", "FEE-001", "SG-1", "SECONDARY-COMP-001-SG", "HP-2", "LP-3", "SECONDARY-COMP-001-TURB",
                 "SECONDARY-COMP-001-COND", "TB-LUB-001", "SJE-001", "SJE-002"]
SWEEP_EXTRA_ACTIONS = ["tube_bundle_overhaul", "cleaning", "overhaul", "system_cleaning", "primary_scale_cleaning",
                       "system_coordination_maintenance", "system_steam_quality_maintenance", "load_balancing_maintenance",
                       "vacuum_ejector_mechanical_cleaning", "no_such_action"]


def sweep():
    R.setup_paths()
    from systems.maintenance.maintenance_actions import MaintenanceActionType
    L = R._layout()
    actions = [t.value for t in MaintenanceActionType] + SWEEP_EXTRA_ACTIONS
    rng = np.random.RandomState(23)
    comp, act, ok, start, idx, val, bases, seg_target = [], [], [], [0], [], [], [], []
    params = None
    for ti, cid in enumerate(SWEEP_TARGETS):
        # a fresh degraded plant per target class: the calls of one class act on one evolving state
        rp = runner_style_plant("oil_change" if not cid.startswith("SG") and "-SG" not in cid else "tsp_chemical_cleaning")[0]
        sim = rp.sim
        for _ in range(3):
            z = np.array([0.0, rng.standard_normal(), rng.random_sample(), rng.random_sample(), rng.random_sample()])
            rp.step(8, 1.0, z)
        if params is None:
            params = R.extract_params(sim)
        ms = sim.maintenance_system
        inst = sim.state_manager.get_registered_instance_info()[cid]["instance"]
        # two passes: catalogue order, then (after three more plant steps, so that timers and wear have moved again)
        # reverse order - a call that only resets what an earlier call already reset shows its effect in the other pass
        for order in (actions, actions[::-1]):
            cur = R.extract_state(sim)
            bases.append(cur.copy()); seg_target.append(cid)
            n_ok = n_changed = 0
            for action in order:
                wo = types.SimpleNamespace(metadata={})
                with R.quiet():
                    res = ms._perform_maintenance_action(inst, action, wo)
                a = R.extract_state(sim)
                ch = np.nonzero(~((a == cur) | (np.isnan(a) & np.isnan(cur))))[0]
                comp.append(len(bases) - 1); act.append(action); ok.append(bool(res.success))
                idx.extend(ch.tolist()); val.extend(a[ch].tolist()); start.append(len(idx))
                n_ok += bool(res.success); n_changed += len(ch) > 0
                cur = a
            print(f"[sweep] {cid:26s} {len(order)} actions: {n_ok} succeed, {n_changed} change the state")
            for _ in range(3):
                z = np.array([0.0, rng.standard_normal(), rng.random_sample(), rng.random_sample(), rng.random_sample()])
                rp.step(8, 1.0, z)
    np.savez_compressed(os.path.join(GOLDEN, "maint_effects_sweep.npz"), targets=np.array(seg_target),
                        base=np.array(bases), target=np.array(comp, np.int32), action=np.array(act),
                        success=np.array(ok), start=np.array(start, np.int64), index=np.array(idx, np.int32),
                        value=np.array(val, np.float64), params=params,
                        state_names=np.array(L.field_names("PlantState")))


def orchestrator_sweep(n_per_component=100, seed=77):
    """tests/golden/orchestrator_sweep.json.gz: decisions of the reference's MaintenanceOrchestrator (decision_only, called the
    way StateManager calls it: state_manager.py:1540-1572) for random sets of 1-5 simultaneous violations of one component,
    drawn from that component's own threshold rows of the maintenance template, with values on both sides of every
    promotion / comprehensive trigger."""
    R.setup_paths()
    from systems.maintenance.maintenance_orchestrator import MaintenanceOrchestrator
    with open(os.path.join(_REPO, "nuclear-sim_b200", "data", "maintenance_system_template.json")) as fh:
        cfg = json.load(fh)
    from nuclear_sim_b200.maintenance import ThresholdTable
    rows = ThresholdTable(cfg).rows
    by_comp = {}
    for r in rows:
        by_comp.setdefault(r.component_id, []).append(r)
    orch = MaintenanceOrchestrator()
    rng = np.random.RandomState(seed)
    cases = []
    for cid in list(by_comp) + ["TB-LUB-001", "XX-UNKNOWN-9"]:
        pool = by_comp.get(cid) or by_comp["FWP-1"]
        for _ in range(n_per_component):
            k = int(rng.randint(1, 6))
            pick = [pool[j] for j in rng.choice(len(pool), size=min(k, len(pool)), replace=False)]
            viol = []
            for r in pick:
                thr = float(r.threshold)
                scale = rng.choice([0.2, 0.9, 1.01, 1.2, 2.0, 5.0, 20.0])
                val = float(thr * scale if thr != 0 else scale)
                viol.append({"parameter": r.parameter, "value": val, "action": r.action, "threshold": thr,
                             "comparison": r.comparison, "priority": r.priority, "component_id": r.sub_component})
            requested = viol[0]["action"] if rng.rand() < 0.9 else None
            with R.quiet():
                d = orch.orchestrate_maintenance(component=None, component_id=cid, violations=[dict(v) for v in viol],
                                                 requested_action=requested, decision_only=True)
            cases.append({"component": cid, "violations": viol, "requested": requested, "selected": d.get("selected_action"),
                          "decision": d.get("orchestration_decision")})
    import gzip
    with gzip.open(os.path.join(GOLDEN, "orchestrator_sweep.json.gz"), "wt", compresslevel=9) as fh:
        json.dump({"cases": cases}, fh, separators=(",", ":"))
    from collections import Counter
    print("[orchestrator]", len(cases), "cases;", Counter(c["decision"] for c in cases))


def scenario(name, action, T, dt=5.0, tweak=None):
    """Full step with maintenance; records states, events, work orders."""
    rp, cfg = runner_style_plant(action, dt=dt)
    sim = rp.sim
    sm, ms = sim.state_manager, sim.maintenance_system
    if tweak:
        tweak(sim)
    L = R._layout()
    events = []      # dict(step, t, component, violations[(param, value, action, priority, component_id)], action)
    created = []     # dict(step, t, wo, component, action, priority, planned_start, extracted)
    executed = []    # dict(step, t, wo, component, action, success)
    cur = {"step": -1}
    orig_emit = sm._emit_batched_threshold_violation

    def emit(cid, violations, optimal, ts):
        events.append({"step": cur["step"], "t": float(ts), "component": cid, "action": optimal,
                       "violations": [[v["parameter"], float(v["value"]), v["action"], v["priority"], v.get("component_id")]
                                      for v in violations]})
        return orig_emit(cid, violations, optimal, ts)
    sm._emit_batched_threshold_violation = emit
    orig_create = ms._create_automatic_work_order

    def create(*a, **k):
        wo = orig_create(*a, **k)
        if wo is not None:
            created.append({"step": cur["step"], "t": float(wo.created_date), "wo": wo.work_order_id, "component": wo.component_id,
                            "action": wo.maintenance_actions[0].action_type, "priority": wo.priority.name,
                            "planned_start": float(wo.planned_start_date),
                            "extracted": getattr(wo, "metadata", {}).get("extracted_component_id") if hasattr(wo, "metadata") else None})
        return wo
    ms._create_automatic_work_order = create
    # _execute_work_order raises inside StateManager.record_maintenance_result at HEAD (state_manager.py:1657 reads a
    # non-existent self.current_time), AFTER the component was maintained; the exception aborts the rest of
    # AutoMaintenanceSystem.update and is swallowed by sim.step (sim.py:210-216).  So the execution is recorded at
    # the point the component is actually touched.
    orig_perf = ms._perform_maintenance_action

    def perform(component, action_type, work_order=None):
        res = orig_perf(component, action_type, work_order)
        executed.append({"step": cur["step"], "t": float(ms.last_check_time), "wo": getattr(work_order, "work_order_id", None),
                         "component": getattr(work_order, "component_id", None), "action": action_type,
                         "success": bool(res.success)})
        return res
    ms._perform_maintenance_action = perform

    state0 = R.extract_state(sim)
    params = R.extract_params(sim)
    NS = L.N_STATE
    states = np.zeros((T, NS))
    noise = np.zeros((T, 5))
    rng = np.random.RandomState(2024)
    for t in range(T):
        cur["step"] = t
        z = np.array([0.0, rng.standard_normal(), rng.random_sample(), rng.random_sample(), rng.random_sample()])
        noise[t] = z
        rp.step(8, 1.0, z)
        states[t] = R.extract_state(sim)
    mcfg = cfg.get("maintenance_system", {})
    sn, pn = np.array(L.field_names("PlantState")), np.array(L.field_names("PlantParams"))
    np.savez_compressed(os.path.join(GOLDEN, f"maint_{name}.npz"), state0=state0, params=params, states=states, noise=noise,
                        log=np.array(json.dumps({"events": events, "created": created, "executed": executed,
                                                 "maintenance_system": mcfg})),
                        state_names=sn, param_names=pn)
    print(f"[scenario] {name}: T={T} events={len(events)} created={[(c['step'], c['component'], c['action']) for c in created]} "
          f"executed={[(e['step'], e['component'], e['action'], e['success']) for e in executed]}")


def scenarios():
    scenario("oil_top_off", "oil_top_off", 60)
    scenario("tsp_chemical_cleaning", "tsp_chemical_cleaning", 60)
    scenario("oil_change", "oil_change", 60)
    scenario("scale_removal", "scale_removal", 60)


if __name__ == "__main__":
    if not R.reference_available():
        sys.exit("reference not found")
    what = sys.argv[1:] or ["effects", "scenarios"]
    if "effects" in what:
        effects()
    if "scenarios" in what:
        scenarios()
    if "sweep" in what:
        sweep()
    if "orchestrator" in what:
        orchestrator_sweep()


def state_log_csv(T=12):
    """tests/golden/ref_state_log_oil_top_off.csv: the reference's OWN wide state log (StateManager.export_to_csv,
    state_manager.py:296-386: `time` + 788 columns) for the first T steps of the maint_oil_top_off scenario — the same
    plant, the same noise stream, so row t belongs to states[t] of tests/golden/maint_oil_top_off.npz.  The GPU export
    test compares the device ring buffer's CSV with this file."""
    rp, cfg = runner_style_plant("oil_top_off", dt=5.0)
    rng = np.random.RandomState(2024)
    for t in range(T):
        z = np.array([0.0, rng.standard_normal(), rng.random_sample(), rng.random_sample(), rng.random_sample()])
        rp.step(8, 1.0, z)
    out = os.path.join(GOLDEN, "ref_state_log_oil_top_off.csv")
    with R.quiet():
        rp.sim.state_manager.export_to_csv(out)
    import csv
    rows = list(csv.reader(open(out)))
    print(f"[state-log] {out}: {len(rows) - 1} rows x {len(rows[0])} columns")


if __name__ == "__main__" and "state_log" in sys.argv[1:]:
    state_log_csv()
