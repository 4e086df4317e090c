"""Host side of threshold monitoring, work-order issuance and maintenance execution for the batched engine.

The reference does this per plant in Python objects:
  * StateManager._check_maintenance_thresholds / _find_parameter_in_row_data / _emit_batched_threshold_violation
    (simulator/state/state_manager.py:1307-1410, 1574-1625)            -> flag kernel + ``ThresholdTable`` (binding)
  * MaintenanceOrchestrator._make_maintenance_decision
    (systems/maintenance/maintenance_orchestrator.py:192-380, 469-624)  -> ``orchestrate``
  * AutoMaintenanceSystem.update / _handle_state_manager_threshold / _create_automatic_work_order /
    _execute_scheduled_work_orders / _execute_work_order
    (systems/maintenance/auto_maintenance.py:200-580)                   -> ``BatchedAutoMaintenance``
  * StateManager.record_maintenance_result / _reset_threshold_cooldowns_for_maintenance
    (simulator/state/state_manager.py:1639-1830)                        -> cooldown stamps reset on the device
  * component.perform_maintenance(...)                                  -> nps_apply_maintenance (csrc/plant/maintenance.h)

Events are sparse (a few plants per check), so the bookkeeping is plain Python dicts keyed by plant; everything that
touches all plants every step (threshold compare, cooldown stamps, effects) stays on the device.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence


from ._layout import field_index

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

# ---------------------------------------------------------------------------------------------------------------
# component table of the PWR3000 plant: ids, classes and equipment types as registered by the reference
# (data/pwr3000_components.json is written by oracle/make_column_map.py from a live plant; order = the order
# StateManager.maintenance_thresholds lists them, which is the order events are emitted in)
# ---------------------------------------------------------------------------------------------------------------
ACTION_NAMES = (
    "oil_change", "oil_top_off", "bearing_replacement", "seal_replacement", "component_overhaul", "system_cleaning",
    "bearing_inspection", "impeller_inspection", "impeller_replacement", "lubrication_system_check", "motor_inspection",
    "oil_analysis", "vibration_analysis", "tsp_chemical_cleaning", "tsp_mechanical_cleaning", "tube_bundle_inspection",
    "moisture_separator_maintenance", "scale_removal", "eddy_current_testing", "secondary_side_cleaning",
    "routine_maintenance", "tube_interior_scale_cleaning", "primary_scale_cleaning", "cleaning", "blade_replacement",
    "overhaul", "condenser_tube_cleaning", "condenser_tube_plugging", "condenser_chemical_cleaning", "vacuum_system_test",
    "vacuum_leak_detection", "turbine_performance_test", "turbine_system_optimization", "turbine_protection_test",
    "thermal_stress_analysis", "system_coordination_maintenance", "system_steam_quality_maintenance",
    "load_balancing_maintenance", "water_chemistry_adjustment", "tsp_inspection", "tsp_flow_test",
    "tube_interior_inspection", "tube_interior_eddy_current_testing", "primary_chemistry_optimization",
    "condenser_water_treatment", "turbine_oil_change", "turbine_oil_top_off", "oil_filter_replacement",
    "oil_cooler_cleaning", "lubrication_system_test", "vacuum_ejector_cleaning", "vacuum_ejector_nozzle_replacement",
    "vacuum_ejector_inspection", "vacuum_ejector_mechanical_cleaning", "other")
ACTION_CODE = {n: i for i, n in enumerate(ACTION_NAMES)}
BEARING_ARG = {None: 0, "": 0, "all": 0, "motor_bearings": 1, "pump_bearings": 2, "thrust_bearing": 3}


def target_code(component_id: str) -> int:
    """MaintTarget of csrc/plant/maintenance.h for a reference component id."""
    if component_id.startswith("FWP-"):
        return int(component_id[4:]) - 1
    if component_id == "FEE-001":
        return 4
    if component_id.startswith("SG-"):
        return 5 + int(component_id[3:])
    if component_id == "SECONDARY-COMP-001-SG":
        return 8
    if component_id.startswith("HP-"):
        return 9 + int(component_id[3:]) - 1
    if component_id.startswith("LP-"):
        return 9 + 8 + int(component_id[3:]) - 1
    if component_id == "SECONDARY-COMP-001-TURB":
        return 23
    if component_id == "SECONDARY-COMP-001-COND":
        return 24
    if component_id == "TB-LUB-001":
        return 25
    if component_id.startswith("SJE-"):
        return 26 + int(component_id[4:]) - 1
    raise KeyError(component_id)


def action_code(action: str) -> int:
    return ACTION_CODE.get(action, ACTION_CODE["other"])


def load_components() -> List[dict]:
    with open(os.path.join(_DATA, "pwr3000_components.json")) as fh:
        return json.load(fh)["components"]


def load_reference_columns() -> Dict[str, dict]:
    with open(os.path.join(_DATA, "reference_columns.json")) as fh:
        return json.load(fh)["columns"]


# ---------------------------------------------------------------------------------------------------------------
# threshold table
# ---------------------------------------------------------------------------------------------------------------
DERIVED_CODES = {"pump_sum_wear": 0}   # csrc: field = -(2 + 4 * code + unit) ... see ThresholdTable.device_rows


@dataclass
class ThresholdRow:
    component_id: str
    parameter: str
    comparison: str
    threshold: float
    cooldown_hours: float
    action: str
    priority: str
    sub_component: Optional[str]      # threshold_config['component_id'] (bearing selector)
    column: Optional[str] = None      # reference column the checker resolves, None = inert (no column matches)
    field: Optional[str] = None       # PlantState field carrying that column
    derived: Optional[str] = None     # or a derived quantity
    unit: Optional[str] = None        # PlantState prefix of the unit a derived quantity is evaluated on
    note: str = ""


def _subsystem_to_equipment(name: str) -> str:
    # StateManager._create_maintenance_config_from_comprehensive: state_manager.py:1107-1112
    return {"feedwater": "pump", "turbine": "turbine_stage", "steam_generator": "steam_generator",
            "condenser": "condenser"}.get(name, name)


def candidate_columns(component_id: str, parameter: str) -> List[str]:
    """The seven names StateManager._find_parameter_in_row_data tries, in order (state_manager.py:1385-1402)."""
    p = {"impeller_inspection_wear": "impeller_wear"}.get(parameter, parameter)
    c = component_id
    return [f"{c}.{p}", f"secondary.feedwater_{c}.{p}", f"secondary.feedwater.{p}", f"secondary.{c}.{p}",
            f"secondary.steam_generator_{c}.{p}", f"secondary.turbine_{c}.{p}", f"secondary.condenser_{c}.{p}"]


class ThresholdTable:
    """Threshold rows of one plant design, bound to PlantState fields the way the reference binds them to columns."""

    def __init__(self, maintenance_system_config: dict, components: Optional[List[dict]] = None,
                 columns: Optional[Dict[str, dict]] = None):
        components = components if components is not None else load_components()
        columns = columns if columns is not None else load_reference_columns()
        cfgs = {}
        for subsystem, data in (maintenance_system_config.get("component_configs") or {}).items():
            cfgs[_subsystem_to_equipment(subsystem)] = data.get("thresholds", {}) or {}
        self.mode = maintenance_system_config.get("maintenance_mode", "realistic")
        self.rows: List[ThresholdRow] = []
        fields = field_index()
        for comp in components:
            cid = comp["id"]
            # AutoMaintenanceSystem.setup_monitoring_from_state_manager filters: auto_maintenance.py:135-144
            if cid.endswith("-LUB") or "lubrication" in cid.lower() or any(s in cid for s in ("-CTRL", "-PROT", "-DIAG", "-MON")):
                continue
            for pname, tc in cfgs.get(comp["equipment_type"], {}).items():
                row = ThresholdRow(cid, pname, tc.get("comparison", "greater_than"), tc.get("threshold"),
                                   float(tc.get("cooldown_hours", 24.0)), tc.get("action"), tc.get("priority", "MEDIUM"),
                                   tc.get("component_id"))
                # "turbine_TB-LUB" / lowercase-"t" skip: state_manager.py:1330 (sic)
                if "turbine_TB-LUB" in cid or "t" in cid:
                    row.note = "skipped by the checker (component id contains 't')"
                elif row.threshold is None:
                    row.note = "no threshold value"
                else:
                    for name in candidate_columns(cid, pname):
                        if name in columns:
                            col = columns[name]
                            if not col.get("numeric", False):
                                # isinstance(value, (int, float)) fails -> the loop moves on to the next pattern
                                continue
                            row.column = name
                            if col.get("field") in fields:
                                row.field = col["field"]
                            elif col.get("derived"):
                                row.derived, row.unit = col["derived"], col.get("unit")
                            else:
                                row.note = "column exists in the reference but is not carried in PlantState"
                            break
                    else:
                        row.note = "inert: no logged column matches (same in the reference)"
                self.rows.append(row)

    def __len__(self):
        return len(self.rows)

    def bound(self) -> List[int]:
        return [i for i, r in enumerate(self.rows) if r.field or r.derived]

    def unsupported(self) -> List[ThresholdRow]:
        return [r for r in self.rows if r.column and not (r.field or r.derived)]

    def device_rows(self):
        """(field index or derived code, comparator, value, cooldown_hours) per row; inert rows get field -1."""
        fields = field_index()
        out = []
        for r in self.rows:
            if r.field:
                f = fields[r.field]
            elif r.derived:
                unit = int(r.unit.split("[")[1].split("]")[0])
                f = -(2 + 4 * DERIVED_CODES[r.derived] + unit)
            else:
                f = -1
            out.append((f, r.comparison, float(r.threshold) if r.threshold is not None else 0.0, r.cooldown_hours))
        return out


# ---------------------------------------------------------------------------------------------------------------
# orchestrator decision (decision_only mode): maintenance_orchestrator.py:81-380
# ---------------------------------------------------------------------------------------------------------------
_HIERARCHY = {   # maintenance_orchestrator.py:469-624 (only the parts _make_maintenance_decision reads)
    "feedwater_pump": {
        "comprehensive": [
            ("component_overhaul", ["bearing_replacement", "seal_replacement", "oil_change", "motor_inspection",
                                    "impeller_replacement", "bearing_inspection", "oil_analysis", "vibration_analysis"],
             {"multiple_major_actions": 2, "total_violations": 4, "bearing_wear_threshold": 15.0,
              "system_health_factor_threshold": 0.75}),
            ("comprehensive_system_inspection", ["bearing_inspection", "motor_inspection", "impeller_inspection",
                                                 "vibration_analysis", "oil_analysis", "lubrication_inspection"],
             {"multiple_major_actions": 3, "total_violations": 5})],
        "coordinated": {"bearing_replacement": ["oil_change", "oil_analysis", "vibration_analysis"],
                        "impeller_replacement": ["cavitation_analysis", "npsh_analysis", "bearing_inspection"],
                        "seal_replacement": ["oil_analysis", "lubrication_inspection"],
                        "motor_inspection": ["bearing_inspection", "vibration_analysis"]},
        "promotion": {"oil_change": ("bearing_replacement", ["bearing_wear > 5.0", "motor_temperature > 80.0", "vibration_increase > 2.0"]),
                      "bearing_inspection": ("bearing_replacement", ["bearing_wear > 10.0", "vibration_increase > 5.0"]),
                      "oil_top_off": ("oil_change", ["oil_contamination_level > 12.0", "oil_acidity_number > 1.4"]),
                      "lubrication_system_check": ("oil_change", ["oil_contamination_level > 12.0"]),
                      "oil_analysis": ("oil_change", ["oil_contamination_level > 15.0", "oil_acidity_number > 1.4"])}},
    "turbine_stage": {
        "comprehensive": [("turbine_overhaul", ["blade_replacement", "bearing_replacement", "rotor_balancing",
                                                "vibration_analysis", "turbine_oil_change"],
                           {"multiple_major_actions": 2, "vibration_threshold": 10.0, "efficiency_degradation_threshold": 0.15})],
        "coordinated": {"blade_replacement": ["rotor_balancing", "vibration_analysis", "performance_test"],
                        "bearing_replacement": ["turbine_oil_change", "vibration_analysis"],
                        "rotor_balancing": ["vibration_analysis", "critical_speed_test"]},
        "promotion": {"turbine_oil_change": ("bearing_replacement", ["bearing_wear > 8.0", "vibration_increase > 3.0"])}},
    "steam_generator": {
        "comprehensive": [
            ("tube_bundle_overhaul", ["tube_cleaning", "tube_inspection", "tsp_chemical_cleaning", "tsp_mechanical_cleaning",
                                      "eddy_current_testing", "scale_removal", "tube_interior_inspection",
                                      "tube_interior_scale_cleaning", "tube_interior_eddy_current_testing", "tsp_inspection"],
             {"multiple_major_actions": 3, "fouling_threshold": 20.0, "tube_plugging_percentage_threshold": 5.0,
              "heat_transfer_degradation_threshold": 0.25}),
            ("comprehensive_steam_generator_inspection", ["tube_bundle_inspection", "tsp_inspection", "tube_interior_inspection",
                                                          "tube_sheet_inspection", "moisture_separator_maintenance",
                                                          "tsp_flow_test", "eddy_current_testing"],
             {"multiple_major_actions": 4, "total_violations": 6})],
        "coordinated": {"tsp_chemical_cleaning": ["tsp_inspection", "tsp_flow_test", "water_chemistry_adjustment"],
                        "tsp_mechanical_cleaning": ["tsp_inspection", "tube_bundle_inspection"],
                        "tube_interior_scale_cleaning": ["tube_interior_inspection", "primary_chemistry_optimization",
                                                         "tube_interior_eddy_current_testing"],
                        "tube_cleaning": ["tube_inspection", "water_chemistry_adjustment"],
                        "scale_removal": ["water_chemistry_adjustment", "tube_inspection"],
                        "eddy_current_testing": ["tube_bundle_inspection", "tube_interior_inspection"],
                        "primary_chemistry_optimization": ["tube_interior_inspection", "water_chemistry_adjustment"]},
        "promotion": {"tsp_inspection": ("tsp_chemical_cleaning", ["fouling_fraction > 0.15", "heat_transfer_degradation > 0.10"]),
                      "tube_interior_inspection": ("tube_interior_scale_cleaning", ["scale_thickness > 1.0", "thermal_resistance > 0.0005"]),
                      "tsp_flow_test": ("tsp_mechanical_cleaning", ["pressure_drop_ratio > 3.0", "flow_maldistribution > 0.25"]),
                      "tube_cleaning": ("tube_bundle_overhaul", ["fouling_factor > 0.3", "heat_transfer_degradation > 0.2"]),
                      "water_chemistry_adjustment": ("primary_chemistry_optimization", ["scale_formation_rate > 0.01", "chemistry_imbalance > 0.1"])}},
    "condenser": {
        "comprehensive": [("condenser_overhaul", ["condenser_tube_cleaning", "condenser_tube_inspection",
                                                  "vacuum_system_check", "condenser_performance_test"],
                           {"multiple_major_actions": 2, "vacuum_degradation_threshold": 5.0,
                            "heat_transfer_degradation_threshold": 0.15})],
        "coordinated": {"condenser_tube_cleaning": ["condenser_tube_inspection", "vacuum_system_check"],
                        "vacuum_ejector_cleaning": ["vacuum_system_test", "condenser_performance_test"]},
        "promotion": {}},
}


def infer_component_type(component_id: str) -> str:
    """MaintenanceOrchestrator._infer_component_type_from_id: maintenance_orchestrator.py:173-186."""
    c = component_id.lower()
    if "fwp" in c or "feedwater" in c:
        return "feedwater_pump"
    if "tb" in c or "turbine" in c:
        return "turbine_stage"
    if "sg" in c or "steam_generator" in c:
        return "steam_generator"
    if "cd" in c or "condenser" in c:
        return "condenser"
    return "unknown"


def orchestrate(component_id: str, violations: Sequence[dict], requested_action: Optional[str]) -> str:
    """selected_action of _make_maintenance_decision (maintenance_orchestrator.py:188-255)."""
    h = _HIERARCHY.get(infer_component_type(component_id), {})
    v_actions = [v.get("action") for v in violations if v.get("action")]
    if not requested_action:
        if not v_actions:
            return "routine_maintenance"
        prio = {"component_overhaul": 10, "bearing_replacement": 9, "impeller_replacement": 8, "seal_replacement": 7,
                "motor_inspection": 6, "oil_change": 5, "bearing_inspection": 4, "oil_analysis": 3,
                "lubrication_system_check": 2, "oil_top_off": 2, "routine_maintenance": 1}
        requested_action = max(v_actions, key=lambda a: prio.get(a, 0))
    # _check_comprehensive_promotion / _meets_comprehensive_criteria: :257-279, 332-361
    for name, encompasses, trig in h.get("comprehensive", []):
        n_enc = sum(1 for a in v_actions if a in encompasses)
        if "multiple_major_actions" in trig and n_enc >= trig["multiple_major_actions"]:
            return name
        if "total_violations" in trig and len(violations) >= trig["total_violations"]:
            return name
        for key, thr in trig.items():
            if key.endswith("_threshold"):
                pn = key.replace("_threshold", "")
                for v in violations:
                    if v.get("parameter") == pn and v.get("value", 0) > thr:
                        return name
    # _check_action_coordination: :281-303 (the base action itself is selected)
    for base, coordinated in h.get("coordinated", {}).items():
        if requested_action == base and any(a in v_actions for a in coordinated):
            return base
    # _check_action_promotion / _check_promotion_conditions: :305-330, 363-381
    rule = h.get("promotion", {}).get(requested_action)
    if rule:
        promote_to, conditions = rule
        for cond in conditions:
            if ">" in cond:
                pn, thr = cond.split(">")
                pn, thr = pn.strip(), float(thr.strip())
                for v in violations:
                    if v.get("parameter") == pn and v.get("value", 0) > thr:
                        return promote_to
    return requested_action


def single_violation_rules(component_id: str, parameter: str, action: Optional[str]):
    """For the common event with exactly ONE violation the decision of ``orchestrate`` depends only on the value:
    returns [(threshold, selected_action), ...] to test in order (value > threshold -> selected_action), and the
    fallback action.  Derived from the same tables, so ``orchestrate`` stays the single statement of the rules
    (tests/test_maintenance_host.py checks the two against each other)."""
    h = _HIERARCHY.get(infer_component_type(component_id), {})
    if not action:
        return [], "routine_maintenance"
    rules = []
    for name, encompasses, trig in h.get("comprehensive", []):
        n_enc = 1 if action in encompasses else 0
        if "multiple_major_actions" in trig and n_enc >= trig["multiple_major_actions"]:
            return [], name
        if "total_violations" in trig and 1 >= trig["total_violations"]:
            return [], name
        for key, thr in trig.items():
            if key.endswith("_threshold") and key.replace("_threshold", "") == parameter:
                rules.append((thr, name))
    rule = h.get("promotion", {}).get(action)   # coordination needs a second action in the batch: never with one violation
    if rule:
        promote_to, conditions = rule
        for cond in conditions:
            if ">" in cond:
                pn, thr = cond.split(">")
                if pn.strip() == parameter:
                    rules.append((float(thr.strip()), promote_to))
    return rules, action


# ---------------------------------------------------------------------------------------------------------------
# work orders
# ---------------------------------------------------------------------------------------------------------------
_KNOWN_ACTIONS_CACHE: Optional[set] = None
_PRIORITY_RANK = {"LOW": 1, "MEDIUM": 2, "HIGH": 3, "CRITICAL": 4, "EMERGENCY": 5}

# which threshold parameters a performed action re-arms: state_manager.py:1797-1808
_COOLDOWN_RESET = {
    "scale_removal": ["tube_wall_temperature", "fouling_factor", "efficiency", "thermal_resistance"],
    "oil_change": ["oil_level", "oil_contamination", "oil_temperature", "oil_acidity"],
    "oil_top_off": ["oil_level"],
    "bearing_replacement": ["bearing_wear", "bearing_temperature", "vibration_level"],
    "bearing_inspection": ["bearing_wear", "bearing_temperature"],
    "vibration_analysis": ["vibration_level"],
    "chemical_cleaning": ["fouling_factor", "efficiency", "tube_wall_temperature"],
    "tube_cleaning": ["fouling_factor", "tube_wall_temperature"],
    "performance_optimization": ["efficiency"],
}


def known_action_names() -> set:
    """Values of the reference's MaintenanceActionType enum (systems/maintenance/maintenance_actions.py:19-198); a work
    order is only created for a known action (auto_maintenance.py:338-346).  data/maintenance_action_types.json is
    written by oracle/make_column_map.py from the live enum."""
    global _KNOWN_ACTIONS_CACHE
    if _KNOWN_ACTIONS_CACHE is None:
        with open(os.path.join(_DATA, "maintenance_action_types.json")) as fh:
            _KNOWN_ACTIONS_CACHE = set(json.load(fh)["actions"])
    return _KNOWN_ACTIONS_CACHE


@dataclass
class WorkOrder:
    work_order_id: str
    plant: int
    component_id: str
    action: str
    priority: str
    created: float            # minutes
    planned_start: float      # minutes
    sub_component: Optional[str] = None
    status: str = "SCHEDULED"
    executed_at: Optional[float] = None
    success: Optional[bool] = None


@dataclass
class _PlantBook:
    orders: List[WorkOrder] = field(default_factory=list)
    recent_triggers: Dict[str, float] = field(default_factory=dict)
    n_created: int = 0


class BatchedAutoMaintenance:
    """AutoMaintenanceSystem for N plants that share one clock (they are stepped in lockstep).

    ``aggressive`` mirrors setup_monitoring_from_state_manager(aggressive_mode=True) as used by
    MaintenanceScenarioRunner (maintenance_scenario_runner.py:296-299): every priority executes with zero delay.
    """

    def __init__(self, sim, table: ThresholdTable, aggressive: bool = True, head_quirks: bool = True):
        """head_quirks=True reproduces the reference at HEAD: StateManager.record_maintenance_result raises
        AttributeError (state_manager.py:1657 reads a non-existent ``self.current_time``) right after the component
        was maintained; the exception aborts the rest of AutoMaintenanceSystem.update and is swallowed by
        NuclearPlantSimulator.step (sim.py:210-216).  Net effect: at most ONE work order per plant executes per
        15-minute update, no maintenance history is recorded, violations are not cleared and threshold cooldowns
        are never reset.  head_quirks=False runs the code path the reference intends (all due orders execute,
        cooldowns of the addressed parameters are reset)."""
        self.sim = sim
        self.table = table
        self.head_quirks = head_quirks
        self.check_interval_hours = 0.25
        self.last_check_time = 0.0
        self.work_order_cooldown_hours = 24.0
        if aggressive:
            self.delays = {"EMERGENCY": 0.0, "CRITICAL": 0.0, "HIGH": 0.0, "MEDIUM": 0.0, "LOW": 0.0}
        else:   # auto_maintenance.py:452-466 (hours): CRITICAL = half the HIGH delay
            self.delays = {"EMERGENCY": 0.0, "CRITICAL": 0.5, "HIGH": 1.0, "MEDIUM": 4.0, "LOW": 24.0}
        self.books: Dict[int, _PlantBook] = {}
        self._pending: Dict[int, List[WorkOrder]] = {}     # scheduled, not yet executed, per plant in creation order
        self.created_log: List[WorkOrder] = []
        self.executed_log: List[WorkOrder] = []
        self.event_log: List[dict] = []
        self._rows_by_component: Dict[str, List[int]] = {}
        for i, r in enumerate(table.rows):
            self._rows_by_component.setdefault(r.component_id, []).append(i)
        self._single = [single_violation_rules(r.component_id, r.parameter, r.action) for r in table.rows]
        sim.set_thresholds(table.device_rows())

    def state_dict(self) -> dict:
        """Everything a resumed run needs to continue with the same work orders (env.save_checkpoint)."""
        return {"kind": "object", "last_check_time": self.last_check_time, "books": self.books, "pending": self._pending,
                "created_log": self.created_log, "executed_log": self.executed_log, "event_log": self.event_log}

    def load_state_dict(self, d: dict) -> None:
        if d.get("kind") != "object":
            raise ValueError("checkpoint was written by a different bookkeeping class")
        self.last_check_time = d["last_check_time"]
        self.books, self._pending = d["books"], d["pending"]
        self.created_log, self.executed_log, self.event_log = d["created_log"], d["executed_log"], d["event_log"]

    def reset(self, plants) -> None:
        """Forget the work-order books of plants that start a new episode (their clocks restart with the reset): pending
        orders, dedupe stamps and counters.  The simulator's own reset() clears their threshold cooldown stamps."""
        for p in (int(q) for q in plants):
            self.books.pop(p, None)
            self._pending.pop(p, None)

    # -- AutoMaintenanceSystem.update: auto_maintenance.py:200-236 ------------------------------------------------
    def update(self, t_minutes: float) -> List[WorkOrder]:
        """Call after the physics step and BEFORE check(): executes the work orders that are due (sim.py:209-213)."""
        if self.last_check_time > 0.0 and t_minutes - self.last_check_time < self.check_interval_hours * 60:
            return []
        self.last_check_time = t_minutes
        due: List[WorkOrder] = []
        for plant in sorted(self._pending):
            for wo in self._pending[plant]:       # creation order (WorkOrderManager.work_orders dict order)
                if t_minutes >= wo.planned_start:
                    due.append(wo)
                    if self.head_quirks:
                        break
        if not due:
            return []
        status = self.sim.apply_maintenance([(wo.plant, target_code(wo.component_id), action_code(wo.action),
                                              BEARING_ARG.get(wo.sub_component, 0) if wo.action == "bearing_replacement" else 0)
                                             for wo in due])
        for wo, st in zip(due, status):
            if st == 2:
                raise NotImplementedError(f"perform_maintenance on {wo.component_id} is not restated on the device")
            wo.status, wo.executed_at, wo.success = "COMPLETED", t_minutes, bool(st == 1)
            self.executed_log.append(wo)
            pend = self._pending[wo.plant]
            pend.remove(wo)
            if not pend:
                del self._pending[wo.plant]
            if wo.success and not self.head_quirks:   # record_maintenance_result -> _reset_threshold_cooldowns_for_maintenance
                addressed = _COOLDOWN_RESET.get(wo.action, [])
                rows = [i for i in self._rows_by_component.get(wo.component_id, []) if self.table.rows[i].parameter in addressed]
                if rows:
                    self.sim.reset_cooldowns(wo.plant, rows)
        return due

    # -- StateManager._check_maintenance_thresholds + AutoMaintenanceSystem._handle_state_manager_threshold --------
    def check(self, t_minutes: float) -> List[WorkOrder]:
        """Flag kernel + host drain: emits the reference's batched events and creates work orders."""
        self.sim.check_thresholds()
        fired = self.sim.drain_events()
        if not fired:
            return []
        try:
            values = self.sim.read_threshold_values(sorted({p for p, _ in fired}), self.table, events=fired)
        except TypeError:      # engines without the sparse form (tests' CPU stand-in)
            values = self.sim.read_threshold_values(sorted({p for p, _ in fired}), self.table)
        return self._process(t_minutes, fired, values)

    def handle_step_events(self, events) -> List[WorkOrder]:
        """Violations recorded INSIDE fused launches (sim.drain_step_events(): plant, row, step, value, time_minutes,
        sorted by step): the same bookkeeping as check(), step by step, with each step's own time stamp."""
        created: List[WorkOrder] = []
        i, n = 0, len(events)
        while i < n:
            j = i
            while j < n and events["step"][j] == events["step"][i]:
                j += 1
            fired = [(int(p), int(r)) for p, r in zip(events["plant"][i:j], events["row"][i:j])]
            values = {f: float(v) for f, v in zip(fired, events["value"][i:j])}
            created += self._process(float(events["time_minutes"][i]), fired, values)
            i = j
        return created

    def gate_open(self, t_minutes: float) -> bool:
        """Would update(t_minutes) pass its check-interval gate (auto_maintenance.py:213-217)?"""
        return not (self.last_check_time > 0.0 and t_minutes - self.last_check_time < self.check_interval_hours * 60)

    def advance(self, n_steps: int, actions=None, magnitudes=None, noise=None, power_setpoint=None,
                t0_minutes: Optional[float] = None, max_k: int = 128, timers: Optional[dict] = None) -> None:
        """n_steps reference steps with monitoring and automatic maintenance, in as few launches as exactness allows.

        Thresholds are evaluated on the device after every substep, so events carry their own step.  Work orders only
        execute when update() passes its 15-minute gate, and the step at which that happens is known in advance; a
        launch therefore runs up to and including the next gate step, skips the in-launch check of that last substep,
        and the host then does what the reference does at that step, in its order (sim.py:209-223): update() - execute
        due orders on the device - then the threshold check of the gate step (flag kernel)."""
        for _ in self.advance_launches(n_steps, actions, magnitudes, noise, power_setpoint, t0_minutes, max_k, timers):
            pass

    def advance_launches(self, n_steps: int, actions=None, magnitudes=None, noise=None, power_setpoint=None,
                         t0_minutes: Optional[float] = None, max_k: int = 128, timers: Optional[dict] = None):
        """advance() as a generator: yields right after each launch has been QUEUED (nothing waited for yet) and does
        that launch's host work - event drain, bookkeeping, maintenance, gate-step check - when resumed.  A driver that
        holds several independent plant batches (advance_interleaved) resumes them in turn, so one batch's host work
        runs while another batch's launch occupies the GPU."""
        sim = self.sim
        if getattr(sim, "_mon", None) is None:
            sim.enable_monitor()
        dt = sim.dt
        t0 = float(sim.current_time_minutes()) if t0_minutes is None else float(t0_minutes)
        sl = lambda x, a, b: None if x is None else (x[a:b] if getattr(x, "ndim", 0) and x.shape[0] == n_steps and x.ndim >= 2 else x)  # noqa: E731
        done = 0
        while done < n_steps:
            k = 1
            while done + k < n_steps and k < max_k and not self.gate_open(t0 + (done + k) * dt):
                k += 1
            t_end = t0 + (done + k) * dt
            on_gate = self.gate_open(t_end)
            if timers is not None:
                import time as _time
                c0 = _time.perf_counter()
            sim.step(actions=sl(actions, done, done + k), magnitudes=sl(magnitudes, done, done + k),
                     noise=sl(noise, done, done + k), power_setpoint=sl(power_setpoint, done, done + k), K=k,
                     skip_last_check=on_gate)
            yield done + k
            if timers is not None:      # attribute the wall clock: step kernel (synchronised) vs everything after it
                getattr(sim, "synchronize", lambda: None)()
                c1 = _time.perf_counter()
            if timers is not None:
                c_d0 = _time.perf_counter()
            ev_in_launch = sim.drain_step_events()
            if timers is not None:
                timers["event_drain"] = timers.get("event_drain", 0.0) + (_time.perf_counter() - c_d0)
            self.handle_step_events(ev_in_launch)
            if on_gate:
                self.update(t_end)
                self.check(t_end)
            if timers is not None:
                getattr(sim, "synchronize", lambda: None)()
                c2 = _time.perf_counter()
                timers["step_kernel"] = timers.get("step_kernel", 0.0) + (c1 - c0)
                timers["bookkeeping"] = timers.get("bookkeeping", 0.0) + (c2 - c1)
                timers["launches"] = timers.get("launches", 0) + 1
            done += k

    def _process(self, t_minutes: float, fired, values) -> List[WorkOrder]:
        by_plant: Dict[int, Dict[str, List[int]]] = {}
        for plant, t in fired:
            by_plant.setdefault(plant, {}).setdefault(self.table.rows[t].component_id, []).append(t)
        created = []
        for plant in sorted(by_plant):
            # component order = table order (dict order of maintenance_thresholds), parameters in config order
            comps = sorted(by_plant[plant], key=lambda c: self._rows_by_component[c][0])
            for cid in comps:
                violations = []
                for t in sorted(by_plant[plant][cid]):
                    r = self.table.rows[t]
                    violations.append({"parameter": r.parameter, "value": values[(plant, t)], "threshold": r.threshold,
                                       "comparison": r.comparison, "action": r.action, "priority": r.priority,
                                       "component_id": r.sub_component})
                if len(violations) == 1:
                    rules, action = self._single[by_plant[plant][cid][0]]
                    for thr, promoted in rules:
                        if violations[0]["value"] > thr:
                            action = promoted
                            break
                    priority = violations[0]["priority"]
                else:
                    action = orchestrate(cid, violations, violations[0]["action"])
                    priority = max((v["priority"] for v in violations), key=lambda p: _PRIORITY_RANK.get(p, 2))
                self.event_log.append({"t": t_minutes, "plant": plant, "component": cid, "action": action,
                                       "violations": violations})
                sub = None
                for v in violations:   # auto_maintenance.py:259-267
                    if v.get("action") == action and v.get("component_id"):
                        sub = v["component_id"]
                        break
                wo = self._create_work_order(plant, cid, action, priority, t_minutes, sub)
                if wo:
                    created.append(wo)
        return created

    # -- _create_automatic_work_order: auto_maintenance.py:332-450 --------------------------------------------------
    def _create_work_order(self, plant, cid, action, priority, t_minutes, sub) -> Optional[WorkOrder]:
        if not action or action not in known_action_names():
            return None
        book = self.books.setdefault(plant, _PlantBook())
        key = f"{cid}:{action}"
        if key in book.recent_triggers:
            # minutes compared against an hours constant (sic): auto_maintenance.py:351-358
            if t_minutes - book.recent_triggers[key] < self.work_order_cooldown_hours:
                return None
        for wo in self._pending.get(plant, ()):
            if wo.component_id == cid and wo.action == action:      # an active order for the same component and action
                return None
        book.n_created += 1
        prio = priority.upper() if priority.upper() in self.delays else "MEDIUM"
        wo = WorkOrder(f"WO-{book.n_created:06d}", plant, cid, action, prio, t_minutes,
                       t_minutes + self.delays[prio] * 60.0, sub)
        book.orders.append(wo)
        self._pending.setdefault(plant, []).append(wo)
        book.recent_triggers[key] = t_minutes
        self.created_log.append(wo)
        return wo


# ---------------------------------------------------------------------------------------------------------------
# the same bookkeeping, columnar: for batches where tens of thousands of plants can fire in one check
# ---------------------------------------------------------------------------------------------------------------
class ColumnarAutoMaintenance(BatchedAutoMaintenance):
    """BatchedAutoMaintenance with array bookkeeping instead of per-order Python objects.

    Same rules, same order, same results (tests/test_maintenance_host.py runs both on the same scenarios); the per-event
    work is numpy on whole event batches: decisions for the common one-violation event are a table lookup
    (single_violation_rules), dedupe stamps live in one sorted key array, pending orders and the logs are columns.
    Events with several violations of one component in one step (rare) take the orchestrate() path one by one.
    With it the host side of BASELINE config #5 (131 072 plants per GPU, every pump crossing a threshold) costs less
    than the step kernel (profiles/r02_cfg5_*.json); the object form above needs ~14 us per work order."""

    _PRIOS = ("LOW", "MEDIUM", "HIGH", "CRITICAL", "EMERGENCY")

    def __init__(self, sim, table: ThresholdTable, aggressive: bool = True, head_quirks: bool = True):
        super().__init__(sim, table, aggressive, head_quirks)
        import numpy as np
        self.np = np
        rows = table.rows
        self.comp_ids: List[str] = []
        comp_of = {}
        for r in rows:
            if r.component_id not in comp_of:
                comp_of[r.component_id] = len(self.comp_ids)
                self.comp_ids.append(r.component_id)
        self.actions: List[str] = []
        act_of: Dict[str, int] = {}

        def aid(name):
            if name not in act_of:
                act_of[name] = len(self.actions)
                self.actions.append(name)
            return act_of[name]
        R = len(rows)
        self.row_comp = np.array([comp_of[r.component_id] for r in rows], dtype=np.int64)
        assert (np.diff(self.row_comp) >= 0).all(), "threshold rows of one component must be contiguous in the table"
        max_rules = max([len(s[0]) for s in self._single] + [1])
        self.rule_thr = np.full((R, max_rules), np.inf)
        self.rule_act = np.zeros((R, max_rules), dtype=np.int64)
        self.fallback = np.zeros(R, dtype=np.int64)
        self.row_action = np.zeros(R, dtype=np.int64)
        for i, (rules, fb) in enumerate(self._single):
            self.fallback[i] = aid(fb)
            self.row_action[i] = aid(rows[i].action) if rows[i].action else -1
            for j, (thr, act) in enumerate(rules):
                self.rule_thr[i, j], self.rule_act[i, j] = thr, aid(act)
        self._aid = aid
        self.row_prio = np.array([self._PRIOS.index(r.priority.upper()) if r.priority.upper() in self.delays else 1 for r in rows],
                                 dtype=np.int64)
        self.subs: List[Optional[str]] = [None]          # sub-component selectors (threshold_config['component_id']) by index
        for r in rows:
            if r.sub_component and r.sub_component not in self.subs:
                self.subs.append(r.sub_component)
        self.row_sub = np.array([self.subs.index(r.sub_component) if r.sub_component else 0 for r in rows], dtype=np.int64)
        self.sub_arg = np.array([BEARING_ARG.get(q, 0) for q in self.subs], dtype=np.int64)
        self.prio_delay = np.array([self.delays[p] * 60.0 for p in self._PRIOS])
        self._refresh_action_tables()
        self.comp_target = np.array([target_code(c) for c in self.comp_ids], dtype=np.int64)
        self.n_comp = len(self.comp_ids)
        # dedupe stamps: sorted keys (plant, component, action) -> time of the last work order created
        self.rk = np.zeros(0, dtype=np.int64)
        self.rt = np.zeros(0)
        self.n_created = np.zeros(sim.n_plants, dtype=np.int64)
        z = lambda dt=np.int64: np.zeros(0, dtype=dt)   # noqa: E731
        self.pend = {"plant": z(), "comp": z(), "act": z(), "prio": z(), "sub": z(), "seq": z(), "created": z(float), "planned": z(float)}
        self.created_cols: List[dict] = []
        self.executed_cols: List[dict] = []
        self.event_cols: List[dict] = []
        # rows whose cooldown a performed action re-arms (record_maintenance_result path, head_quirks=False)
        self._reset_rows: Dict[tuple, List[int]] = {}
        # wall-clock split of the bookkeeping: pure host work (numpy) vs calls that wait for the device
        # (event drains, the maintenance kernel with its copies, the gate-step flag kernel)
        self.seconds_host = 0.0
        self.seconds_device_calls = 0.0

    def _refresh_action_tables(self):
        np = self.np
        known = known_action_names()
        self.act_known = np.array([a in known for a in self.actions], dtype=bool)
        self.act_code = np.array([action_code(a) for a in self.actions], dtype=np.int64)
        self.act_is_bearing = np.array([a == "bearing_replacement" for a in self.actions], dtype=bool)

    def state_dict(self) -> dict:
        return {"kind": "columnar", "last_check_time": self.last_check_time, "actions": list(self.actions), "rk": self.rk, "rt": self.rt,
                "n_created": self.n_created, "pend": self.pend, "created_cols": self.created_cols,
                "executed_cols": self.executed_cols, "event_cols": self.event_cols}

    def load_state_dict(self, d: dict) -> None:
        if d.get("kind") != "columnar":
            raise ValueError("checkpoint was written by a different bookkeeping class")
        for name in d["actions"]:            # action ids are positions in this list: restore it before the columns
            self._aid(name)
        assert self.actions[:len(d["actions"])] == list(d["actions"])
        self._refresh_action_tables()
        self.last_check_time = d["last_check_time"]
        self.rk, self.rt, self.n_created, self.pend = d["rk"], d["rt"], d["n_created"], d["pend"]
        self.created_cols, self.executed_cols, self.event_cols = d["created_cols"], d["executed_cols"], d["event_cols"]

    def _key(self, plant, comp, act):
        return (plant * self.n_comp + comp) * 4096 + act

    def reset(self, plants) -> None:
        np = self.np
        plants = np.asarray(list(plants), dtype=np.int64)
        keep = ~np.isin(self.pend["plant"], plants)
        self.pend = {k: v[keep] for k, v in self.pend.items()}
        kp = self.rk // (self.n_comp * 4096)
        keep = ~np.isin(kp, plants)
        self.rk, self.rt = self.rk[keep], self.rt[keep]
        self.n_created[plants] = 0

    # -- events -> work orders ---------------------------------------------------------------------------------------
    def handle_step_events(self, events):
        np = self.np
        n = len(events)
        if n == 0:
            return 0
        steps = events["step"]
        cuts = np.flatnonzero(np.diff(steps)) + 1
        made = 0
        for a, b in zip(np.concatenate([[0], cuts]), np.concatenate([cuts, [n]])):
            made += self._process_arrays(float(events["time_minutes"][a]), events["plant"][a:b].astype(np.int64),
                                         events["row"][a:b].astype(np.int64), events["value"][a:b])
        return made

    def check(self, t_minutes: float):
        import time as _time
        if not hasattr(self.sim, "check_thresholds_events"):
            return super().check(t_minutes)
        if getattr(self.sim, "_mon", None) is None:
            self.sim.enable_monitor()
        c0 = _time.perf_counter()
        self.sim.check_thresholds_events()
        ev = self.sim.drain_step_events()
        self.seconds_device_calls += _time.perf_counter() - c0
        if len(ev) == 0:
            return 0
        return self._process_arrays(t_minutes, ev["plant"].astype(self.np.int64), ev["row"].astype(self.np.int64), ev["value"])

    def _process(self, t_minutes, fired, values):      # object-form entry point of the base class
        np = self.np
        if not fired:
            return []
        pl = np.array([f[0] for f in fired], dtype=np.int64)
        rw = np.array([f[1] for f in fired], dtype=np.int64)
        return self._process_arrays(t_minutes, pl, rw, np.array([values[f] for f in fired]))

    def _process_arrays(self, t, plant, row, value) -> int:
        """One step's violations (sorted by (plant, row)) -> events, decisions, work orders."""
        import time as _time
        c0 = _time.perf_counter()
        try:
            return self._process_arrays_impl(t, plant, row, value)
        finally:
            self.seconds_host += _time.perf_counter() - c0

    def _process_arrays_impl(self, t, plant, row, value) -> int:
        np = self.np
        comp = self.row_comp[row]
        gkey = plant * self.n_comp + comp
        # components of a plant in table order, rows in config order.  A component's rows are contiguous in the table, so
        # events that arrive sorted by (plant, row) - drain_step_events() delivers them that way - are already in order
        skey = plant * 4096 + row
        if len(skey) > 1 and not bool((skey[1:] >= skey[:-1]).all()):
            order = np.argsort(skey, kind="stable")
            plant, row, value, comp, gkey = plant[order], row[order], value[order], comp[order], gkey[order]
        first = np.concatenate([[True], gkey[1:] != gkey[:-1]])
        starts = np.flatnonzero(first)
        counts = np.diff(np.concatenate([starts, [len(gkey)]]))
        g_plant, g_comp, g_row, g_val = plant[starts], comp[starts], row[starts], value[starts]
        # one-violation events: table lookup (first matching rule wins)
        act = self.fallback[g_row].copy()
        for j in range(self.rule_thr.shape[1] - 1, -1, -1):
            m = g_val > self.rule_thr[g_row, j]
            act[m] = self.rule_act[g_row[m], j]
        prio = self.row_prio[g_row].copy()
        sub = np.where(self.row_action[g_row] == act, self.row_sub[g_row], 0)
        multi = np.flatnonzero(counts > 1)
        for gi in multi:                                 # several violations of one component in one step: orchestrate()
            lo, hi = starts[gi], starts[gi] + counts[gi]
            viol = []
            for t_row, v in zip(row[lo:hi], value[lo:hi]):
                r = self.table.rows[int(t_row)]
                viol.append({"parameter": r.parameter, "value": float(v), "threshold": r.threshold, "comparison": r.comparison,
                             "action": r.action, "priority": r.priority, "component_id": r.sub_component})
            name = orchestrate(self.comp_ids[int(g_comp[gi])], viol, viol[0]["action"])
            if name not in self.actions:
                self._aid(name)
                self._refresh_action_tables()
            act[gi] = self.actions.index(name)
            pr = max((v["priority"] for v in viol), key=lambda q: _PRIORITY_RANK.get(q, 2))
            prio[gi] = self._PRIOS.index(pr.upper()) if pr.upper() in self.delays else 1
            sub[gi] = 0
            for v in viol:
                if v.get("action") == name and v.get("component_id"):
                    sub[gi] = self.subs.index(v["component_id"])
                    break
        self.event_cols.append({"t": t, "plant": g_plant, "comp": g_comp, "act": act.copy(), "starts": starts, "n": counts,
                                "row_all": row, "value_all": value})
        # _create_automatic_work_order: known action, dedupe stamp (minutes vs hours, sic), no active order for (component, action)
        ok = self.act_known[act] & (act >= 0)
        key = self._key(g_plant, g_comp, act)
        if len(self.rk):
            # a stamp only ever blocks while t - stamp < 24 (minutes against the hours constant, sic) and the clock only
            # moves forward, so older stamps are dropped instead of carried: the key array stays a few checks long
            live = (t - self.rt) < self.work_order_cooldown_hours
            if not live.all():
                self.rk, self.rt = self.rk[live], self.rt[live]
        if len(self.rk):
            pos = np.minimum(np.searchsorted(self.rk, key), len(self.rk) - 1)
            ok &= ~(self.rk[pos] == key)
        if len(self.pend["plant"]):      # an active order for the same (component, action): sorted lookup (np.isin is 15x slower here)
            pk = np.sort(self._key(self.pend["plant"], self.pend["comp"], self.pend["act"]))
            pos = np.minimum(np.searchsorted(pk, key), len(pk) - 1)
            ok &= ~(pk[pos] == key)
        idx = np.flatnonzero(ok)
        if len(idx) == 0:
            return 0
        c_plant, c_comp, c_act, c_prio, c_sub, c_key = g_plant[idx], g_comp[idx], act[idx], prio[idx], sub[idx], key[idx]
        # WO-%06d numbering per plant, in creation order (plants ascending, components in table order)
        firstp = np.concatenate([[True], c_plant[1:] != c_plant[:-1]])
        run_start = np.maximum.accumulate(np.where(firstp, np.arange(len(c_plant)), 0))
        rank = np.arange(len(c_plant)) - run_start
        seq = self.n_created[c_plant] + rank + 1
        up, cnt = np.unique(c_plant, return_counts=True)
        self.n_created[up] += cnt
        created = np.full(len(idx), t)
        planned = t + self.prio_delay[c_prio]
        new = {"plant": c_plant, "comp": c_comp, "act": c_act, "prio": c_prio, "sub": c_sub, "seq": seq, "created": created, "planned": planned}
        self.pend = {k: np.concatenate([self.pend[k], new[k]]) for k in self.pend}
        self.created_cols.append(new)
        # dedupe stamps of the new orders (their keys are not in the array: a live stamp would have blocked them)
        allk = np.concatenate([self.rk, c_key])
        allt = np.concatenate([self.rt, created])
        o = np.argsort(allk, kind="stable")
        self.rk, self.rt = allk[o], allt[o]
        return len(idx)

    # -- due work orders -> device -------------------------------------------------------------------------------------
    def update(self, t_minutes: float):
        import time as _time
        c0, d0 = _time.perf_counter(), self.seconds_device_calls
        try:
            return self._update_impl(t_minutes)
        finally:
            self.seconds_host += (_time.perf_counter() - c0) - (self.seconds_device_calls - d0)

    def _update_impl(self, t_minutes: float):
        np = self.np
        if not self.gate_open(t_minutes):
            return 0
        self.last_check_time = t_minutes
        P = self.pend
        if len(P["plant"]) == 0:
            return 0
        due = np.flatnonzero(t_minutes >= P["planned"])
        if len(due) == 0:
            return 0
        if self.head_quirks:      # at most one order per plant per update: the first due one in creation order
            o = np.lexsort((due, P["plant"][due]))      # pending rows are in creation order; group by plant
            pl = P["plant"][due][o]
            firstp = np.concatenate([[True], pl[1:] != pl[:-1]])
            due = np.sort(due[o][firstp])
        else:
            due = due[np.lexsort((due, P["plant"][due]))]
        d_act = P["act"][due]
        req = np.stack([P["plant"][due], self.comp_target[P["comp"][due]], self.act_code[d_act],
                        np.where(self.act_is_bearing[d_act], self.sub_arg[P["sub"][due]], 0)], axis=1).astype(np.int32)
        import time as _time
        c0 = _time.perf_counter()
        status = np.asarray(self.sim.apply_maintenance(req))
        self.seconds_device_calls += _time.perf_counter() - c0
        if (status == 2).any():
            bad = int(np.flatnonzero(status == 2)[0])
            raise NotImplementedError(f"perform_maintenance on {self.comp_ids[int(P['comp'][due][bad])]} is not restated on the device")
        done = {k: v[due] for k, v in P.items()}
        done["executed_at"] = np.full(len(due), t_minutes)
        done["success"] = status == 1
        self.executed_cols.append(done)
        keep = np.ones(len(P["plant"]), dtype=bool)
        keep[due] = False
        self.pend = {k: v[keep] for k, v in P.items()}
        if not self.head_quirks:
            for i in np.flatnonzero(done["success"]):
                name, cid = self.actions[int(done["act"][i])], self.comp_ids[int(done["comp"][i])]
                rows = self._reset_rows.get((cid, name))
                if rows is None:
                    addressed = _COOLDOWN_RESET.get(name, [])
                    rows = [r for r in self._rows_by_component.get(cid, []) if self.table.rows[r].parameter in addressed]
                    self._reset_rows[(cid, name)] = rows
                if rows:
                    self.sim.reset_cooldowns(int(done["plant"][i]), rows)
        return len(due)

    # -- object views of the logs (tests, exports) -------------------------------------------------------------------
    def _orders(self, cols, executed=False) -> List[WorkOrder]:
        out = []
        for c in cols:
            for i in range(len(c["plant"])):
                wo = WorkOrder(f"WO-{int(c['seq'][i]):06d}", int(c["plant"][i]), self.comp_ids[int(c["comp"][i])],
                               self.actions[int(c["act"][i])], self._PRIOS[int(c["prio"][i])], float(c["created"][i]),
                               float(c["planned"][i]), self.subs[int(c["sub"][i])])
                if executed:
                    wo.status, wo.executed_at, wo.success = "COMPLETED", float(c["executed_at"][i]), bool(c["success"][i])
                out.append(wo)
        return out

    @property
    def n_work_orders_created(self) -> int:
        return int(sum(len(c["plant"]) for c in self.created_cols))

    @property
    def n_work_orders_executed(self) -> int:
        return int(sum(len(c["plant"]) for c in self.executed_cols))

    def counts_by_action(self) -> Dict[str, int]:
        np = self.np
        out: Dict[str, int] = {}
        for c in self.created_cols:
            ids, cnt = np.unique(c["act"], return_counts=True)
            for a, k in zip(ids, cnt):
                out[self.actions[int(a)]] = out.get(self.actions[int(a)], 0) + int(k)
        return out

    def materialize_logs(self) -> None:
        """Fill created_log / executed_log / event_log (lists of objects, as BatchedAutoMaintenance keeps them)."""
        # (plant, number, creation time): numbering restarts when a plant's episode is reset
        ex = {(w.plant, w.work_order_id, w.created): w for w in self._orders(self.executed_cols, executed=True)}
        self.created_log = [ex.get((w.plant, w.work_order_id, w.created), w) for w in self._orders(self.created_cols)]
        self.executed_log = list(ex.values())
        self.event_log = []
        for c in self.event_cols:
            for i in range(len(c["plant"])):
                lo, hi = int(c["starts"][i]), int(c["starts"][i]) + int(c["n"][i])
                viol = []
                for t_row, v in zip(c["row_all"][lo:hi], c["value_all"][lo:hi]):
                    r = self.table.rows[int(t_row)]
                    viol.append({"parameter": r.parameter, "value": float(v), "threshold": r.threshold, "comparison": r.comparison,
                                 "action": r.action, "priority": r.priority, "component_id": r.sub_component})
                self.event_log.append({"t": c["t"], "plant": int(c["plant"][i]), "component": self.comp_ids[int(c["comp"][i])],
                                       "action": self.actions[int(c["act"][i])], "violations": viol})


class NativeAutoMaintenance(ColumnarAutoMaintenance):
    """ColumnarAutoMaintenance with the per-step bookkeeping in the library's native work-order table
    (csrc/nps_workorders.cpp, nps_wo_* in include/nps_b200.h): grouping of a step's violations, the single-violation
    decision, dedupe stamps, the pending table, numbering and the due-order selection run in C++ over plain arrays;
    numpy only carries the log columns.  Same rules, same order, same results as the two classes above
    (tests/test_maintenance_host.py runs all three on the same scenarios); events with several violations still go
    through orchestrate() one by one."""

    def __init__(self, sim, table: ThresholdTable, aggressive: bool = True, head_quirks: bool = True):
        super().__init__(sim, table, aggressive, head_quirks)
        import ctypes
        from . import _clib
        np = self.np
        self._L = _clib.lib()
        self._ct = ctypes
        rule_thr = np.ascontiguousarray(self.rule_thr, dtype=np.float64)
        rule_act = np.ascontiguousarray(self.rule_act, dtype=np.int64)
        h = ctypes.c_void_p()
        keep = [np.ascontiguousarray(a, dtype=np.int64) for a in (self.row_comp, self.fallback, self.row_action, self.row_prio, self.row_sub)]
        delay = np.ascontiguousarray(self.prio_delay, dtype=np.float64)
        rc = self._L.nps_wo_create(int(sim.n_plants), int(self.n_comp), int(len(self.row_comp)), int(rule_thr.shape[1]),
                                   self._p(keep[0]), self._p(rule_thr), self._p(rule_act), self._p(keep[1]), self._p(keep[2]),
                                   self._p(keep[3]), self._p(keep[4]), self._p(delay), float(self.work_order_cooldown_hours),
                                   int(bool(head_quirks)), ctypes.byref(h))
        if rc != 0:
            raise ValueError("nps_wo_create: bad arguments")
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._L.nps_wo_destroy(h)
            except Exception:
                pass

    @staticmethod
    def _p(a):
        return a.__array_interface__["data"][0]      # the raw address (argtypes are c_void_p); ndarray.ctypes costs ~7 us a call

    # -- events -> work orders ---------------------------------------------------------------------------------------
    def _process_arrays_impl(self, t, plant, row, value) -> int:
        np = self.np
        plant = np.ascontiguousarray(plant, dtype=np.int64)
        row = np.ascontiguousarray(row, dtype=np.int64)
        value = np.ascontiguousarray(value, dtype=np.float64)
        n = len(plant)
        if n == 0:
            return 0
        skey = plant * 4096 + row
        if n > 1 and not bool((skey[1:] >= skey[:-1]).all()):
            order = np.argsort(skey, kind="stable")
            plant, row, value = plant[order], row[order], value[order]
        g = [np.empty(n, dtype=np.int64) for _ in range(7)]       # start, count, plant, comp, act, prio, sub
        n_multi = self._ct.c_int64(0)
        ng = int(self._L.nps_wo_group(self._h, n, self._p(plant), self._p(row), self._p(value), *[self._p(a) for a in g],
                                      self._ct.byref(n_multi)))
        if ng < 0:
            raise ValueError("nps_wo_group: event outside the threshold table / plant range")
        starts, counts, g_plant, g_comp, act, prio, sub = [a[:ng] for a in g]
        if n_multi.value:
            for gi in np.flatnonzero(counts > 1):                # several violations of one component in one step
                lo, hi = int(starts[gi]), int(starts[gi] + counts[gi])
                viol = []
                for t_row, v in zip(row[lo:hi], value[lo:hi]):
                    r = self.table.rows[int(t_row)]
                    viol.append({"parameter": r.parameter, "value": float(v), "threshold": r.threshold, "comparison": r.comparison,
                                 "action": r.action, "priority": r.priority, "component_id": r.sub_component})
                name = orchestrate(self.comp_ids[int(g_comp[gi])], viol, viol[0]["action"])
                if name not in self.actions:
                    self._aid(name)
                    self._refresh_action_tables()
                act[gi] = self.actions.index(name)
                pr = max((v["priority"] for v in viol), key=lambda q: _PRIORITY_RANK.get(q, 2))
                prio[gi] = self._PRIOS.index(pr.upper()) if pr.upper() in self.delays else 1
                sub[gi] = 0
                for v in viol:
                    if v.get("action") == name and v.get("component_id"):
                        sub[gi] = self.subs.index(v["component_id"])
                        break
        self.event_cols.append({"t": t, "plant": g_plant, "comp": g_comp, "act": act.copy(), "starts": starts, "n": counts,
                                "row_all": row, "value_all": value})
        ok = np.ascontiguousarray(self.act_known, dtype=np.uint8)
        out_g, out_seq = np.empty(ng, dtype=np.int64), np.empty(ng, dtype=np.int64)
        made = int(self._L.nps_wo_issue(self._h, float(t), ng, self._p(g_plant), self._p(g_comp), self._p(act), self._p(prio),
                                        self._p(sub), self._p(ok), len(ok), self._p(out_g), self._p(out_seq)))
        if made < 0:
            raise ValueError("nps_wo_issue: bad arguments")
        if made == 0:
            return 0
        idx = out_g[:made]
        c_prio = prio[idx]
        self.created_cols.append({"plant": g_plant[idx], "comp": g_comp[idx], "act": act[idx], "prio": c_prio, "sub": sub[idx],
                                  "seq": out_seq[:made].copy(), "created": np.full(made, float(t)),
                                  "planned": float(t) + self.prio_delay[c_prio]})
        return made

    # -- due work orders -> device -------------------------------------------------------------------------------------
    def _update_impl(self, t_minutes: float):
        np = self.np
        if not self.gate_open(t_minutes):
            return 0
        self.last_check_time = t_minutes
        cap = int(self._L.nps_wo_n_pending(self._h))
        if cap == 0:
            return 0
        ci = [np.empty(cap, dtype=np.int64) for _ in range(6)]    # plant, comp, act, prio, sub, seq
        cf = [np.empty(cap, dtype=np.float64) for _ in range(2)]  # created, planned
        nd = int(self._L.nps_wo_due(self._h, float(t_minutes), cap, *[self._p(a) for a in ci], *[self._p(a) for a in cf]))
        if nd <= 0:
            return 0
        d_plant, d_comp, d_act, d_prio, d_sub, d_seq = [a[:nd] for a in ci]
        req = np.stack([d_plant, self.comp_target[d_comp], self.act_code[d_act],
                        np.where(self.act_is_bearing[d_act], self.sub_arg[d_sub], 0)], axis=1).astype(np.int32)
        import time as _time
        c0 = _time.perf_counter()
        status = np.asarray(self.sim.apply_maintenance(req))
        self.seconds_device_calls += _time.perf_counter() - c0
        if (status == 2).any():
            bad = int(np.flatnonzero(status == 2)[0])
            raise NotImplementedError(f"perform_maintenance on {self.comp_ids[int(d_comp[bad])]} is not restated on the device")
        self._L.nps_wo_complete(self._h)
        done = {"plant": d_plant, "comp": d_comp, "act": d_act, "prio": d_prio, "sub": d_sub, "seq": d_seq,
                "created": cf[0][:nd], "planned": cf[1][:nd], "executed_at": np.full(nd, t_minutes), "success": status == 1}
        self.executed_cols.append(done)
        if not self.head_quirks:
            for i in np.flatnonzero(done["success"]):
                name, cid = self.actions[int(done["act"][i])], self.comp_ids[int(done["comp"][i])]
                rows = self._reset_rows.get((cid, name))
                if rows is None:
                    addressed = _COOLDOWN_RESET.get(name, [])
                    rows = [r for r in self._rows_by_component.get(cid, []) if self.table.rows[r].parameter in addressed]
                    self._reset_rows[(cid, name)] = rows
                if rows:
                    self.sim.reset_cooldowns(int(done["plant"][i]), rows)
        return nd

    def reset(self, plants) -> None:
        np = self.np
        plants = np.ascontiguousarray(list(plants), dtype=np.int64)
        self._L.nps_wo_reset_plants(self._h, self._p(plants), len(plants))
        self.n_created[plants] = 0

    # -- checkpoints -------------------------------------------------------------------------------------------------
    def state_dict(self) -> dict:
        np, ct = self.np, self._ct
        n_p, n_s = ct.c_int64(0), ct.c_int64(0)
        self._L.nps_wo_sizes(self._h, ct.byref(n_p), ct.byref(n_s))
        pc, pt = np.zeros((6, n_p.value), dtype=np.int64), np.zeros((2, n_p.value), dtype=np.float64)
        sk, stt = np.zeros((2, n_s.value), dtype=np.int64), np.zeros(n_s.value, dtype=np.float64)
        ncr = np.zeros(self.sim.n_plants, dtype=np.int64)
        self._L.nps_wo_export(self._h, self._p(pc), self._p(pt), self._p(sk), self._p(stt), self._p(ncr))
        return {"kind": "native", "last_check_time": self.last_check_time, "actions": list(self.actions), "pend_cols": pc,
                "pend_times": pt, "stamp_keys": sk, "stamp_times": stt, "n_created": ncr, "created_cols": self.created_cols,
                "executed_cols": self.executed_cols, "event_cols": self.event_cols}

    def load_state_dict(self, d: dict) -> None:
        if d.get("kind") != "native":
            raise ValueError("checkpoint was written by a different bookkeeping class")
        np = self.np
        for name in d["actions"]:
            self._aid(name)
        assert self.actions[:len(d["actions"])] == list(d["actions"])
        self._refresh_action_tables()
        self.last_check_time = d["last_check_time"]
        pc = np.ascontiguousarray(d["pend_cols"], dtype=np.int64)
        pt = np.ascontiguousarray(d["pend_times"], dtype=np.float64)
        sk = np.ascontiguousarray(d["stamp_keys"], dtype=np.int64)
        stt = np.ascontiguousarray(d["stamp_times"], dtype=np.float64)
        ncr = np.ascontiguousarray(d["n_created"], dtype=np.int64)
        rc = self._L.nps_wo_import(self._h, pc.shape[1], self._p(pc), self._p(pt), sk.shape[1], self._p(sk), self._p(stt), self._p(ncr))
        if rc != 0:
            raise ValueError("nps_wo_import: bad arguments")
        self.created_cols, self.executed_cols, self.event_cols = d["created_cols"], d["executed_cols"], d["event_cols"]


def advance_interleaved(maints: Sequence["BatchedAutoMaintenance"], n_steps: int, inputs: Optional[Sequence[dict]] = None,
                        streams: Optional[Sequence] = None, max_k: int = 128) -> None:
    """advance() for several INDEPENDENT plant batches on one GPU, software-pipelined from one host thread.

    Plants do not interact, so a batch can be cut into parts with their own simulator and bookkeeping object (the same
    global plant ids, initial conditions and inputs).  Each part's launches go to its own CUDA stream; the parts are
    resumed in turn (advance_launches), so while the host drains and processes the events of part A the GPU runs the
    launch of part B.  Results are those of advance() on every part - and therefore those of one big batch.
    inputs: per part, the keyword arguments of advance() (actions, magnitudes, noise, power_setpoint, t0_minutes)."""
    import contextlib
    n = len(maints)
    inputs = list(inputs) if inputs is not None else [{} for _ in range(n)]
    if streams is None:
        try:
            import torch
            streams = [torch.cuda.Stream(device=m.sim.device) for m in maints] if all(hasattr(m.sim, "slab") for m in maints) else None
        except Exception:
            streams = None

    def ctx(i):
        if streams is None:
            return contextlib.nullcontext()
        import torch
        return torch.cuda.stream(streams[i])
    if streams is not None:
        import torch
        for i, m in enumerate(maints):                 # what was queued on the current stream before (state writes) comes first
            streams[i].wait_stream(torch.cuda.current_stream(m.sim.device))
    gens = []
    for i, m in enumerate(maints):
        with ctx(i):
            gens.append(m.advance_launches(n_steps, max_k=max_k, **inputs[i]))
    live = list(range(n))
    while live:
        for i in list(live):
            with ctx(i):
                try:
                    next(gens[i])
                except StopIteration:
                    live.remove(i)
    if streams is not None:
        import torch
        for i, m in enumerate(maints):
            torch.cuda.current_stream(m.sim.device).wait_stream(streams[i])


def advance_threaded(maints: Sequence["BatchedAutoMaintenance"], n_steps: int, inputs: Optional[Sequence[dict]] = None,
                     max_k: int = 128) -> None:
    """advance() for several INDEPENDENT plant batches on one GPU, one host thread and one CUDA stream per batch.

    Same contract as advance_interleaved; here the overlap comes from the threads: a thread that waits for its stream
    (event drain, maintenance status) or runs inside the native work-order table has released the interpreter lock, so
    another batch's host work proceeds and its launch is queued as soon as it is ready."""
    import threading
    import torch
    n = len(maints)
    inputs = list(inputs) if inputs is not None else [{} for _ in range(n)]
    streams = [torch.cuda.Stream(device=m.sim.device) for m in maints]
    for i, m in enumerate(maints):
        streams[i].wait_stream(torch.cuda.current_stream(m.sim.device))
    errors: List[BaseException] = []

    def work(i):
        try:
            torch.cuda.set_device(maints[i].sim.device)
            with torch.cuda.stream(streams[i]):
                maints[i].advance(n_steps, max_k=max_k, **inputs[i])
        except BaseException as e:      # noqa: BLE001  (re-raised in the caller's thread)
            errors.append(e)
    threads = [threading.Thread(target=work, args=(i,)) for i in range(n)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    for i, m in enumerate(maints):
        torch.cuda.current_stream(m.sim.device).wait_stream(streams[i])
