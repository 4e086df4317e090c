"""Trajectory export in the reference's schemas.

* wide CSV  — StateManager.export_to_csv / export_by_category / export_by_subcategory
              (nuclear_simulator/simulator/state/state_manager.py:296-386): one row per logged step, first column
              ``time`` (ISO datetime), then ``category.subcategory[_<id>].variable`` columns in the reference's order;
* long CSV  — data/plant_data_logger.py:82-157: ``timestamp,parameter_name,value,unit,quality``.

The reference builds each row from Python ``get_state_dict()`` calls (789 columns).  The batched engine carries the
same quantities as PlantState fields; data/reference_columns.json (written by oracle/make_column_map.py from a live
plant) names the field behind every column.  Columns that are constants in the reference are written as that constant;
the values the reference computes on the fly for the log (flow restrictions, heat-flow bookkeeping, alarm counts ...)
are the ReportState fields the step kernel writes (csrc/plant/report.h).  Anything without a source would be listed in
``ColumnSchema.unavailable``; with the round-1 schema that list is empty (788 of 788 columns).
"""
from __future__ import annotations

import csv
import re
import datetime as _dt
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from ._layout import field_index
from .maintenance import load_reference_columns


# PumpStatus values (systems/primary/coolant/pump_models.py:19-26) in the order of the status code in csrc/plant/state.h
PUMP_STATUS_NAMES = ("running", "stopped", "starting", "stopping", "tripped")

# FeedwaterPumpLubricationSystem.maintenance_action_flags keys: feedwater/pump_lubrication.py:90-104
PUMP_MAINTENANCE_FLAGS = ("oil_change", "oil_top_off", "bearing_replacement", "seal_replacement", "component_overhaul",
                          "system_cleaning", "bearing_inspection", "impeller_inspection", "impeller_replacement",
                          "lubrication_system_check", "motor_inspection", "oil_analysis", "vibration_analysis")


class ColumnSchema:
    """Reference column order + how to obtain each column from a PlantState vector."""

    def __init__(self, columns: Optional[Dict[str, dict]] = None):
        cols = columns if columns is not None else load_reference_columns()
        ix = field_index()
        self.names: List[str] = []
        self.kind: List[Tuple[str, object, str]] = []     # (how, payload, python type name)
        self.unavailable: List[str] = []
        for name, e in cols.items():
            if name == "time":
                continue
            leaf = name.rsplit(".", 1)[-1]
            if leaf.endswith("_occurred") and "feedwater_FWP-" in name:
                # one-shot flags set by perform_maintenance and cleared by the next get_state_dict
                # (feedwater/pump_lubrication.py:642-643, 1636-1641): true only in the row of the step that executed it
                comp = name.split(".")[1].replace("feedwater_", "")
                self.names.append(name); self.kind.append(("event", (comp, leaf[:-len("_occurred")]), e["type"]))
                continue
            m = re.match(r"secondary\.feedwater_FWP-(\d+)\.status$", name)
            if m:    # PumpStatus.value string (feedwater/pump_system.py:1068) from the status code (csrc/plant/state.h)
                self.names.append(name); self.kind.append(("pump_status", ix[f"fw.pump[{int(m.group(1)) - 1}].status"], "str"))
                continue
            if leaf == "heat_flow_energy_balance_ok" and "rep.hf_energy_balance_percent" in ix:
                # HeatFlowTracker.validate_energy_balance: |percent error| < validation_tolerance * 100
                # (systems/secondary/heat_flow_tracker.py:224,342)
                self.names.append(name); self.kind.append(("balance_ok", ix["rep.hf_energy_balance_percent"], "bool"))
                continue
            if e.get("field") in ix:
                self.names.append(name); self.kind.append(("field", ix[e["field"]], e["type"]))
            elif e.get("derived") == "pump_sum_wear":
                u = e["unit"]
                w = [ix[f"{u}lub.component_wear[{c}]"] for c in range(5)]
                self.names.append(name); self.kind.append(("sum_wear", w, e["type"]))
            elif "const" in e:
                self.names.append(name); self.kind.append(("const", e["const"], e["type"]))
            elif e.get("expr") and all(f in ix for f in e.get("fields", [])):
                # a rescaled field or an aggregate over repeated units (found by value on live plants, see make_column_map.py)
                self.names.append(name); self.kind.append((e["expr"], ([ix[f] for f in e["fields"]], float(e.get("scale", 1.0))), e["type"]))
            else:
                self.unavailable.append(name)

    def select(self, prefix: Optional[str] = None) -> List[int]:
        return [i for i, n in enumerate(self.names) if prefix is None or n.startswith(prefix)]

    def logged_fields(self, which: Sequence[int]) -> List[int]:
        """PlantState field indices needed to produce the selected columns (for the device ring buffer)."""
        need = []
        for i in which:
            how, payload, _ = self.kind[i]
            if how in ("field", "pump_status", "balance_ok"):
                need.append(payload)
            elif how == "sum_wear":
                need += list(payload)
            elif how not in ("const", "event"):
                need += list(payload[0])
        return sorted(set(need))

    def row(self, state: np.ndarray, which: Sequence[int], events=frozenset()) -> list:
        """events: {(component_id, action)} executed in the step this row closes."""
        out = []
        for i in which:
            how, payload, typ = self.kind[i]
            if how == "event":
                comp, action = payload
                out.append(any(c == comp and (action == "maintenance_action" or a == action) and a in PUMP_MAINTENANCE_FLAGS
                               for c, a in events))
                continue
            if how == "pump_status":
                out.append(PUMP_STATUS_NAMES[int(state[payload])])
                continue
            if how == "balance_ok":
                out.append(bool(abs(float(state[payload])) < 0.01 * 100))
                continue
            if how == "field":
                v = float(state[payload])
            elif how == "sum_wear":
                w = state[payload]
                v = float(w[0] + max(w[1], w[2], w[3]) + w[4])
            elif how == "const":
                v = payload
            else:
                x, scale = state[payload[0]], payload[1]
                v = {"scaled": lambda: x[0], "sum": lambda: x.sum(), "mean": lambda: x.sum() / len(x), "max": lambda: x.max(),
                     "min": lambda: x.min(), "mean_abs": lambda: np.abs(x).sum() / len(x), "max_abs": lambda: np.abs(x).max()}[how]()
                v = float(v) * scale
            if typ in ("bool", "bool_"):
                out.append(bool(v))
            elif typ == "int":
                out.append(int(v))
            else:
                out.append(v)
        return out


class TrajectoryStore:
    """Host-side row store with StateManager's export surface for ONE plant's rows (full PlantState per step)."""

    def __init__(self, start_datetime: Optional[_dt.datetime] = None, max_rows: int = 100000):
        self.start_datetime = start_datetime or _dt.datetime(2024, 1, 1)
        self.times: List[_dt.datetime] = []
        self.rows: List[np.ndarray] = []
        self.events: List[frozenset] = []
        self.max_rows = int(max_rows)
        self.schema = ColumnSchema()

    def add_row(self, when: _dt.datetime, state: np.ndarray, events=frozenset()) -> None:
        self.times.append(when)
        self.rows.append(np.array(state, dtype=np.float64, copy=True))
        self.events.append(frozenset(events))
        if len(self.rows) > self.max_rows:       # StateManager trims the oldest rows when over capacity
            del self.rows[0]; del self.times[0]; del self.events[0]

    def clear(self) -> None:
        self.times.clear(); self.rows.clear(); self.events.clear()

    def _write(self, filename: str, which: Sequence[int], time_range=None) -> int:
        n = 0
        with open(filename, "w", newline="") as fh:
            w = csv.writer(fh)
            w.writerow(["time"] + [self.schema.names[i] for i in which])
            for t, s, ev in zip(self.times, self.rows, self.events):
                if time_range is not None and not (time_range[0] <= t <= time_range[1]):
                    continue
                w.writerow([t.isoformat()] + self.schema.row(s, which, ev))
                n += 1
        return n

    def export_to_csv(self, filename: str, time_range=None, variables: Optional[Iterable[str]] = None) -> int:
        which = self.schema.select() if variables is None else [self.schema.names.index(v) for v in variables]
        return self._write(filename, which, time_range)

    def export_by_category(self, category: str, filename: str, time_range=None) -> int:
        return self._write(filename, self.schema.select(f"{category}."), time_range)

    def export_by_subcategory(self, category: str, subcategory: str, filename: str, time_range=None) -> int:
        return self._write(filename, self.schema.select(f"{category}.{subcategory}."), time_range)


def export_ring_to_csv(filename: str, ring: np.ndarray, logged_field_ids: Sequence[int], plant: int,
                       start_datetime: _dt.datetime, dt_minutes: float, first_row_step: int = 1,
                       schema: Optional[ColumnSchema] = None, which: Optional[Sequence[int]] = None,
                       events: Optional[Sequence[frozenset]] = None) -> int:
    """One plant's trajectory out of the device ring buffer (BatchedNuclearPlantSimulator.drain_log():
    [rows, n_logged, n_plants]) as the reference's wide CSV (StateManager.export_to_csv, state_manager.py:296-386).
    `logged_field_ids` are the PlantState indices of the ring's field axis; columns whose fields were not logged are
    left out unless they are constants.  Row r is stamped start + (first_row_step + r) * dt."""
    schema = schema or ColumnSchema()
    have = {int(f): j for j, f in enumerate(logged_field_ids)}
    which = list(schema.select() if which is None else which)
    which = [i for i in which if set(schema.logged_fields([i])) <= set(have)]
    n_state = len(field_index())
    n = 0
    with open(filename, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["time"] + [schema.names[i] for i in which])
        for r in range(ring.shape[0]):
            state = np.full(n_state, np.nan)
            for f, j in have.items():
                state[f] = ring[r, j, plant]
            t = start_datetime + _dt.timedelta(minutes=(first_row_step + r) * dt_minutes)
            w.writerow([t.isoformat()] + schema.row(state, which, events[r] if events else frozenset()))
            n += 1
    return n


# PlantDataLogger.extract_all_parameters (data/plant_data_logger.py:91-136): 22 parameters per step, in this order
def plant_data_logger_parameters(state: np.ndarray, simulation_time: float) -> List[Tuple[str, object, str]]:
    ix = field_index()
    g = lambda f: float(state[ix[f]])
    out = [("neutron_flux", g("pri.neutron_flux"), "neutrons/cm\u00b2/s"), ("reactivity", g("pri.reactivity"), "\u0394k/k")]
    out += [(f"delayed_neutron_precursors_group_{i + 1}", g(f"pri.precursors[{i}]"), "relative") for i in range(6)]
    out += [("fuel_temperature", g("pri.fuel_temperature"), "\u00b0C"), ("coolant_temperature", g("pri.coolant_temperature"), "\u00b0C"),
            ("coolant_pressure", g("pri.coolant_pressure"), "MPa"), ("coolant_flow_rate", g("pri.coolant_flow_rate"), "kg/s"),
            ("steam_temperature", g("pri.steam_temperature"), "\u00b0C"), ("steam_pressure", g("pri.steam_pressure"), "MPa"),
            ("steam_flow_rate", g("pri.steam_flow_rate"), "kg/s"), ("feedwater_flow_rate", g("pri.feedwater_flow_rate"), "kg/s"),
            ("control_rod_position", g("pri.control_rod_position"), "%"), ("steam_valve_position", g("pri.steam_valve_position"), "%"),
            ("power_level", g("pri.power_level"), "%"), ("scram_status", int(g("pri.scram_status")), "boolean"),
            ("thermal_power", g("pri.neutron_flux") / 1e12 * 3000, "MW"), ("simulation_time", float(simulation_time), "s")]
    return out


def export_long_format(filename: str, timestamps: Sequence[str], states: Sequence[np.ndarray], sim_times: Sequence[float],
                       quality: str = "GOOD") -> int:
    """PlantDataLogger long format: timestamp,parameter_name,value,unit,quality (data/plant_data_logger.py:82-157)."""
    n = 0
    with open(filename, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["timestamp", "parameter_name", "value", "unit", "quality"])
        for ts, s, t in zip(timestamps, states, sim_times):
            for name, value, unit in plant_data_logger_parameters(s, t):
                w.writerow([ts, name, value, unit, quality])
                n += 1
    return n
