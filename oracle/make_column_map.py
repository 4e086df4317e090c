"""TEST/DEV INFRASTRUCTURE — derive the reference-column -> PlantState-field map from the LIVE reference.

The reference logs 789 columns per step (StateManager.collect_states, simulator/state/state_manager.py:152-211) named
``category.subcategory[_<id>].variable``; threshold monitoring (state_manager.py:1371-1410) and the CSV exports
(state_manager.py:296-407) are keyed by those names.  The batched engine carries the same quantities as PlantState
fields under its own names; this script pairs them by VALUE over several steps of a perturbed live plant (a column
maps to a field when their values agree bit for bit at every recorded step and the pairing is unambiguous, name
similarity breaking ties) and records the Python type of each column (the threshold checker only accepts int/float,
state_manager.py:1404-1407).

Writes nuclear-sim_b200/data/reference_columns.json.  Run in the build container only (needs /root/reference).
"""
from __future__ import annotations

import difflib
import json
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
sys.path.insert(0, _REPO)
from oracle import refplant as R  # noqa: E402


def record(action="oil_top_off", steps=7, seed=3):
    R.setup_paths()
    cfg = R.compose_config(action, duration_hours=1.0)
    rp = R.make_reference_plant(cfg, dt=5.0, heat_source="constant", noise_enabled=True, noise_std_percent=0.5,
                                enable_state_management=True)
    sim = rp.sim
    sm = sim.state_manager
    rows, states = [], []
    orig = sm._check_maintenance_thresholds

    def chk(ts, row):
        rows.append(dict(row))
        states.append(R.extract_state(sim))
        return orig(ts, row)
    sm._check_maintenance_thresholds = chk
    rng = np.random.RandomState(seed)
    for t in range(steps):
        z = np.array([rng.standard_normal(), rng.standard_normal(), rng.random_sample(), rng.random_sample(),
                      rng.random_sample()])
        sim.primary_physics.heat_source.set_power_setpoint(92.0 + 1.5 * t)
        rp.step(8, 1.0, z)
    return rows, np.array(states)


def preferred_prefix(col: str):
    """PlantState prefix of the component a reference column belongs to (state_manager.py:502-545 naming)."""
    import re
    parts = col.split(".")
    if parts[0] == "primary":
        return "pri."
    if len(parts) < 3:
        return None
    sub = parts[1]
    m = re.match(r"feedwater_FWP-(\d+)$", sub)
    if m:
        return f"fw.pump[{int(m.group(1)) - 1}]."
    m = re.match(r"steam_generator_SG-(\d+)$", sub)
    if m:
        return f"sgs.sg[{int(m.group(1))}]."
    m = re.match(r"turbine_HP-(\d+)$", sub)
    if m:
        return f"turb.stage[{int(m.group(1)) - 1}]."
    m = re.match(r"turbine_LP-(\d+)$", sub)
    if m:
        return f"turb.stage[{int(m.group(1)) + 7}]."
    if sub == "turbine_TB-LUB-001":
        return "turb.lub."
    if sub.startswith("turbine"):
        # the turbine system's dict is updated by every stage in turn (stage_system.py:1029-1030): LP-6 wins
        return "turb.stage[13]." if parts[2] in STAGE_KEYS else "turb."
    if sub.startswith("steam_generator"):
        return "sgs."
    if sub.startswith("feedwater"):
        return "fw."
    if sub.startswith("condenser"):
        return "cond."
    if sub == "water_chemistry":
        return "wc_main."
    if sub.startswith("ph_control"):
        return "ph."
    return "sec."


def report_field(col: str):
    """ReportState field (csrc/plant/state.h) that carries a per-unit report column, by construction of its name."""
    import re
    parts = col.split(".")
    if len(parts) < 3:
        return None
    m = re.match(r"feedwater_FWP-(\d+)$", parts[1])
    if m:
        return f"rep.fwp_{parts[2]}[{int(m.group(1)) - 1}]"
    m = re.match(r"steam_generator_SG-(\d+)$", parts[1])
    if m:
        return f"rep.sg_{parts[2]}[{int(m.group(1))}]"
    return None


STAGE_KEYS = ("inlet_pressure", "outlet_pressure", "inlet_temperature", "outlet_temperature", "power_output", "efficiency",
              "extraction_flow", "loading_factor", "blade_condition", "deposit_thickness", "operating_hours")

# columns that are pure functions of carried fields (evaluated by the flag / export kernels)
DERIVED = {
    # FeedwaterPumpLubricationSystem.get_state_dict: feedwater/pump_lubrication.py:1592-1596
    "sum_wear_level": "pump_sum_wear",
}


def main():
    L = R._layout()
    names = L.field_names()
    rows, states = [], []
    for action, seed in (("oil_top_off", 3), ("tsp_chemical_cleaning", 5), ("scale_removal", 9)):
        r, s = record(action, seed=seed)
        rows += r
        states.append(s)
    states = np.concatenate(states)
    cols = list(rows[0].keys())
    out = {}
    n_direct = 0
    for c in cols:
        vals = [r.get(c) for r in rows]
        v0 = vals[0]
        pytype = type(v0).__name__
        entry = {"type": pytype, "numeric": isinstance(v0, (int, float)), "field": None}
        if c.rsplit(".", 1)[-1] in DERIVED and "feedwater_FWP-" in c:
            entry["derived"] = DERIVED[c.rsplit(".", 1)[-1]]
            entry["unit"] = preferred_prefix(c)
        try:
            arr = np.array([float(v) for v in vals], dtype=np.float64)
        except (TypeError, ValueError):
            out[c] = entry
            continue
        eq = np.all((states == arr[:, None]) | (np.isnan(states) & np.isnan(arr[:, None])), axis=0)
        cand = [names[i] for i in np.nonzero(eq)[0]]
        if cand:
            pref = preferred_prefix(c)
            inside = [f for f in cand if pref and f.startswith(pref)]
            rep = report_field(c)          # report-only quantity of THIS unit (ReportState arrays are indexed by unit)
            if rep in cand:
                inside = [rep]
            if not inside and not np.ptp(arr) > 0:
                entry["const"] = float(arr[0])     # a constant that happens to equal unrelated fields: not a mapping
                out[c] = entry
                continue
            pool = inside or cand
            tail = c.rsplit(".", 1)[-1]
            best = max(pool, key=lambda f: difflib.SequenceMatcher(None, tail, f.rsplit(".", 1)[-1]).ratio())
            entry["field"] = best
            entry["n_candidates"] = len(pool)
            entry["in_prefix"] = bool(inside)
            entry["varies"] = bool(np.ptp(arr) > 0)
            n_direct += 1
        out[c] = entry
    # second pass: columns that are a constant, a rescaled field, or an aggregate over the repeated units
    import re
    idx = {n: i for i, n in enumerate(names)}
    groups = {}
    for n in names:
        m = re.match(r"^(sgs\.sg|fw\.pump|turb\.bearing|turb\.stage|cond\.ejector)\[(\d+)\]\.(.+)$", n)
        if m:
            groups.setdefault((m.group(1), m.group(3)), []).append(n)
        m = re.match(r"^(.*)\[(\d+)\]$", n)          # plain arrays (sec.prev_sg_levels[i], fw.lc_level_errors[i], ...)
        if m:
            groups.setdefault(("array", m.group(1)), []).append(n)
    scales = [1.0, 1e-6, 1e-3, 1e3, 1e6, 100.0, 0.01, 60.0, 1 / 60.0, 3600.0, 1 / 3600.0]

    def close(a, b):
        return np.all(np.abs(a - b) <= 1e-12 * np.maximum(1.0, np.abs(b)))
    n_second = 0
    for c in cols:
        e = out[c]
        if e.get("field") or e.get("derived") or "const" in e or not e.get("numeric"):
            continue
        try:
            arr = np.array([float(r.get(c)) for r in rows], dtype=np.float64)
        except (TypeError, ValueError):
            continue
        if not np.ptp(arr) > 0:
            e["const"] = float(arr[0]); n_second += 1
            continue
        pref = preferred_prefix(c) or ""
        found = None
        for sc_ in scales[1:]:
            hits = [f for f in names if close(states[:, idx[f]] * sc_, arr)]
            hits = [f for f in hits if f.startswith(pref)] or hits
            if len(hits) == 1:
                found = {"expr": "scaled", "fields": hits, "scale": sc_}
                break
        if not found:
            for (unit, leaf), members in groups.items():
                cols_ = states[:, [idx[f] for f in members]]
                for op, val in (("sum", cols_.sum(1)), ("mean", cols_.sum(1) / len(members)), ("max", cols_.max(1)), ("min", cols_.min(1)),
                                ("mean_abs", np.abs(cols_).sum(1) / len(members)), ("max_abs", np.abs(cols_).max(1))):
                    for sc_ in (1.0, 1e-6, 1e-3, 1e3, 100.0):
                        if np.ptp(val) > 0 and close(val * sc_, arr):
                            found = {"expr": op, "fields": members, "scale": sc_}
                            break
                    if found:
                        break
                if found:
                    break
        if found:
            e.update(found); n_second += 1
    print(f"second pass: {n_second} more columns (constants / rescaled / unit aggregates)")
    path = os.path.join(_REPO, "nuclear-sim_b200", "data", "reference_columns.json")
    with open(path, "w") as fh:
        json.dump({"columns": out, "n_columns": len(cols), "n_direct": n_direct}, fh, indent=0, sort_keys=False)
    print(f"{len(cols)} columns, {n_direct} map to a PlantState field -> {path}")
    dump_components()
    return out


def dump_components():
    """Ordered component table (id, class, equipment type) and the MaintenanceActionType value list."""
    R.setup_paths()
    cfg = R.compose_config("oil_top_off", duration_hours=1.0)
    rp = R.make_reference_plant(cfg, dt=5.0, heat_source="constant", enable_state_management=True)
    sim = rp.sim
    sm, ms = sim.state_manager, sim.maintenance_system
    with R.quiet():
        sm.config = cfg
        ms.setup_monitoring_from_state_manager(sm, aggressive_mode=True)
        from simulator.state.component_metadata import ComponentRegistry
        from systems.maintenance.maintenance_actions import MaintenanceActionType
    reg = sm.get_registered_instance_info()
    order = list(sm.maintenance_thresholds.keys()) + [c for c in reg if c not in sm.maintenance_thresholds]
    comps = []
    for cid in order:
        meta = ComponentRegistry.get_component(cid)
        et = meta["metadata"].equipment_type.value if meta else None
        comps.append({"id": cid, "class_name": reg[cid]["class_name"], "equipment_type": et,
                      "monitored": cid in sm.maintenance_thresholds})
    data = os.path.join(_REPO, "nuclear-sim_b200", "data")
    with open(os.path.join(data, "pwr3000_components.json"), "w") as fh:
        json.dump({"components": comps}, fh, indent=0)
    with open(os.path.join(data, "maintenance_action_types.json"), "w") as fh:
        json.dump({"actions": sorted(a.value for a in MaintenanceActionType)}, fh, indent=0)
    print(f"{len(comps)} components, {len(list(MaintenanceActionType))} action types")


if __name__ == "__main__":
    main()
