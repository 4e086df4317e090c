"""``NuclearPlantSimulator`` — the reference's scalar simulator surface on top of the batched engine (N = 1 view).

Mirrors nuclear_simulator/simulator/core/sim.py:27-258 (constructor arguments, ``step()`` return dict, ``reset()``,
``get_observation()``, ``.dt``, ``.state``, ``.primary_physics.heat_source``, ``.secondary_physics``, ``.state_manager``,
``.maintenance_system``) for the attributes the reference's callers actually read
(data_gen/runners/maintenance_scenario_runner.py:232-238, 383-405, 651-671, 949-994, 1084-1096), so a
runner written against the reference can drive a plant that lives in the batched engine's SoA slab.

Initial conditions: the reference turns nested config into a plant through ~5 000 lines of dataclass /
``_apply_initial_conditions`` code (initialisation, not the step path).  This adapter starts from a PlantState /
PlantParams vector pair: a committed snapshot (``snapshot=...``), or the vectors a reference-side binding extracts from
a plant built by the reference's own constructors (INTEGRATION.md shows that binding).

Random draws: the reference consumes ``ConstantHeatSource.rng.normal(0, sigma)`` (RandomState(noise_seed), one per
step) and global ``np.random`` in the pH controller.  The host draws the same legacy streams
(``RandomState(seed).standard_normal()`` reproduces ``rng.normal(0, s)/s`` bit for bit) and ships them to the device.
"""
from __future__ import annotations

import datetime as _dt
import warnings
from types import SimpleNamespace
from typing import Any, Dict, List, Optional

import numpy as np

from ._layout import field_index
from .export import TrajectoryStore
from .maintenance import BatchedAutoMaintenance, ThresholdTable
from .snapshots import load_snapshot

NO_ACTION = 8


class _ScalarState:
    """``sim.state.<attr>`` -> python float of plant `p` (ReactorState attribute names, primary/__init__.py:48-106)."""

    def __init__(self, owner, prefixes=("pri.",)):
        object.__setattr__(self, "_o", owner)
        object.__setattr__(self, "_pre", prefixes)

    def _idx(self, name):
        ix = field_index()
        for pre in self._pre:
            if pre + name in ix:
                return ix[pre + name]
        if name == "delayed_neutron_precursors":
            return [ix[f"pri.precursors[{i}]"] for i in range(6)]
        raise AttributeError(name)

    def __getattr__(self, name):
        i = self._idx(name)
        row = self._o._row()
        if isinstance(i, list):
            return np.array([row[j] for j in i])
        v = float(row[i])
        return bool(v) if name in ("scram_status", "feedwater_pump_status", "feedwater_system_available") else v

    def __setattr__(self, name, value):
        i = self._idx(name)
        self._o._write_fields({i: float(value)})


class _HeatSource:
    """ConstantHeatSource / ReactorHeatSource surface (heat_sources/constant_heat_source.py:44-102)."""

    def __init__(self, owner, rated_power_mw):
        self._o = owner
        self.rated_power_mw = float(rated_power_mw)

    def set_power_setpoint(self, power_percent: float) -> None:
        sp = max(0.0, min(150.0, float(power_percent)))     # constant_heat_source.py:94-102 clamps to [0, 150]
        ix = field_index()
        self._o._write_fields({ix["pri.hs_setpoint_percent"]: sp,
                               ix["pri.hs_current_power_mw"]: (sp / 100.0) * self.rated_power_mw})

    @property
    def power_setpoint_percent(self):
        return float(self._o._row()[field_index()["pri.hs_setpoint_percent"]])

    def get_thermal_power_mw(self):
        return float(self._o._row()[field_index()["pri.thermal_power_mw"]])


class _StateManagerFacade:
    """The StateManager members the runners touch (state_manager.py:98-125, 296-386, 1627-1700, 1846-1882)."""

    def __init__(self, owner, max_rows):
        self._o = owner
        self.config = None
        self.start_datetime = _dt.datetime(2024, 1, 1)      # the reference picks a random 2024 date (state_manager.py:89-93)
        self.current_datetime = self.start_datetime
        self.store = TrajectoryStore(self.start_datetime, max_rows)
        self.maintenance_history: List[dict] = []

    def advance_time(self, dt_minutes):
        self.current_datetime += _dt.timedelta(minutes=dt_minutes)
        return self.current_datetime

    def get_elapsed_time(self):
        return self.current_datetime - self.start_datetime

    def get_current_threshold_violations(self) -> dict:
        m = self._o._maint
        out: Dict[str, dict] = {}
        if m is not None:
            for e in m.event_log:   # last batched event per component, as StateManager stores it (:1607-1612)
                out.setdefault(e["component"], {})["batched_violations"] = {
                    "component_id": e["component"], "parameter": "multiple_violations", "value": len(e["violations"]),
                    "action": e["action"], "violations": e["violations"], "timestamp": e["t"]}
        return out

    def get_maintenance_history(self, component_id=None) -> list:
        return [r for r in self.maintenance_history if component_id is None or r["component_id"] == component_id]

    def get_component_state_snapshot(self, component_id: str) -> dict:
        """Latest logged values of one component, keyed by the reference's variable names."""
        sc = self.store.schema
        row = self._o._row()
        out = {}
        for i, name in enumerate(sc.names):
            parts = name.split(".")
            if len(parts) >= 3 and (parts[1].endswith("_" + component_id) or parts[1] == component_id):
                out[parts[-1]] = sc.row(row, [i])[0]
        return out

    def verify_maintenance_action(self, component_id, action_type, expected_changes=None, tolerance=0.1) -> bool:
        return any(r["component_id"] == component_id and r["action_type"] == action_type and r["success"]
                   for r in reversed(self.maintenance_history))

    def export_to_csv(self, filename, time_range=None, variables=None):
        return self.store.export_to_csv(filename, time_range, variables)

    def export_by_category(self, category, filename, time_range=None):
        return self.store.export_by_category(category, filename, time_range)

    def export_by_subcategory(self, category, subcategory, filename, time_range=None):
        return self.store.export_by_subcategory(category, subcategory, filename, time_range)

    def get_data_info(self) -> dict:
        return {"rows": len(self.store.rows), "columns": len(self.store.schema.names) + 1,
                "unavailable_columns": len(self.store.schema.unavailable)}

    def clear_data(self):
        self.store.clear()


_PRIORITY_VALUE = {"LOW": 1, "MEDIUM": 2, "HIGH": 3, "CRITICAL": 4, "EMERGENCY": 5}


class _WorkOrderView:
    """Reference-shaped view of a work order (systems/maintenance/work_orders.py:67-135) for code written against
    WorkOrder attributes / to_dict() — e.g. MaintenanceScenarioRunner._check_maintenance_triggers (:696-819)."""

    def __init__(self, wo):
        self._wo = wo
        self.work_order_id, self.component_id = wo.work_order_id, wo.component_id
        a = wo.action
        kind = ("emergency" if wo.priority == "EMERGENCY" else "inspection" if ("inspection" in a or "analysis" in a)
                else "cleaning" if ("cleaning" in a or "flush" in a) else "corrective")     # auto_maintenance.py:389-397
        self.work_order_type = SimpleNamespace(value=kind)
        self.priority = SimpleNamespace(value=_PRIORITY_VALUE.get(wo.priority, 2), name=wo.priority)
        self.title = f"Auto: {a.replace('_', ' ').title()} - {wo.component_id}"
        self.description = "Automatic maintenance triggered by: State manager threshold"
        self.created_date, self.planned_start_date = wo.created, wo.planned_start
        self.planned_duration = 0.0
        self.auto_generated, self.trigger_id = True, f"event_{wo.created}"

    @property
    def status(self):
        return SimpleNamespace(value=self._wo.status.lower())

    def to_dict(self) -> dict:
        wo = self._wo
        return {"work_order_id": wo.work_order_id, "component_id": wo.component_id, "work_order_type": self.work_order_type.value,
                "priority": self.priority.value, "status": wo.status.lower(), "title": self.title,
                "description": self.description, "created_date": wo.created, "planned_start_date": wo.planned_start,
                "actual_start_date": wo.executed_at, "actual_completion_date": wo.executed_at,
                "auto_generated": True, "trigger_id": self.trigger_id,
                "maintenance_actions": [{"action_type": wo.action, "success": wo.success}]}


class _MaintenanceFacade:
    """AutoMaintenanceSystem members the runners touch (auto_maintenance.py:66-96, 98-197, 675-760)."""

    def __init__(self, owner):
        self._o = owner
        self.check_interval_hours = 0.25
        self.auto_execute_maintenance = True
        self.current_update_work_orders: list = []

    def setup_monitoring_from_state_manager(self, state_manager, aggressive_mode: bool = False):
        cfg = (state_manager.config or {}).get("maintenance_system") if isinstance(state_manager.config, dict) else None
        if not cfg:
            return
        self._o._maint = BatchedAutoMaintenance(self._o._engine, ThresholdTable(cfg), aggressive=aggressive_mode)

    def get_system_status(self) -> dict:
        m = self._o._maint
        return {"auto_execute_enabled": self.auto_execute_maintenance, "check_interval_hours": self.check_interval_hours,
                "last_check_time": m.last_check_time if m else 0.0,
                "work_orders_created": len(m.created_log) if m else 0,
                # at HEAD the counter sits after the statement that raises (auto_maintenance.py:567-580): it stays 0
                "work_orders_executed": (0 if (m is None or m.head_quirks) else len(m.executed_log)),
                "state_manager_stats": {"maintenance_thresholds": len({r.component_id for r in m.table.rows}) if m else 0}}

    def get_recent_work_orders(self, limit: int = 10) -> list:
        """auto_maintenance.py:707-722: dicts of the most recent work orders, newest first."""
        m = self._o._maint
        if m is None:
            return []
        out = [_WorkOrderView(wo).to_dict() for wo in m.created_log[-limit:]]
        out.sort(key=lambda d: d["created_date"], reverse=True)
        return out[:limit]


class NuclearPlantSimulator:
    """Scalar facade over plant `plant_index` of a batched engine (engine=None builds a 1-plant CUDA engine)."""

    def __init__(self, dt: float = 1.0, heat_source=None, enable_secondary: bool = True,
                 enable_state_management: bool = True, max_state_rows: int = 100000,
                 secondary_config=None, secondary_config_file: str = None, *,
                 initial_state: Optional[np.ndarray] = None, params: Optional[np.ndarray] = None,
                 snapshot: Optional[str] = None, engine=None, plant_index: int = 0,
                 ph_seed: int = 1000, device: str = "cuda:0"):
        if secondary_config_file is not None:
            raise NotImplementedError("secondary_config_file: build the plant with the reference's config code and pass "
                                      "initial_state/params (INTEGRATION.md)")
        ixp = field_index("PlantParams")
        if initial_state is None or params is None:
            s0, p0 = load_snapshot(snapshot or "pwr3000_oil_top_off_dt5")
            initial_state = s0 if initial_state is None else initial_state
            params = p0 if params is None else params
        params = np.array(params, dtype=np.float64, copy=True)
        params[ixp["dt"]] = float(dt)
        params[ixp["enable_secondary"]] = float(bool(enable_secondary))
        self._heat_rng = None
        if heat_source is not None:     # ConstantHeatSource(...) / ReactorHeatSource(...) from the reference, or a look-alike
            is_const = type(heat_source).__name__ == "ConstantHeatSource" or hasattr(heat_source, "noise_enabled")
            params[ixp["heat_source_type"]] = 0.0 if is_const else 1.0
            params[ixp["rated_power_mw"]] = float(getattr(heat_source, "rated_power_mw", 3000.0))
            params[ixp["noise_enabled"]] = float(bool(getattr(heat_source, "noise_enabled", False)))
            params[ixp["noise_std_percent"]] = float(getattr(heat_source, "noise_std_percent", 0.0))
            params[ixp["noise_filter_time_constant"]] = float(getattr(heat_source, "noise_filter_time_constant", 30.0))
            if params[ixp["noise_enabled"]]:
                # the reference keeps its generator in heat_source.rng (constant_heat_source.py:60); drawing
                # standard_normal() from that same object continues the very stream rng.normal(0, sigma) would consume
                rng = getattr(heat_source, "rng", None)
                self._heat_rng = rng if hasattr(rng, "standard_normal") else np.random.RandomState(getattr(heat_source, "noise_seed", None))
        self._ph_rng = np.random.RandomState(ph_seed)
        self.dt = float(dt)
        self.enable_secondary = bool(enable_secondary)
        self.enable_state_management = bool(enable_state_management)
        self._p = int(plant_index)
        if engine is None:
            from .batched import BatchedNuclearPlantSimulator
            engine = BatchedNuclearPlantSimulator(1, np.asarray(initial_state, dtype=np.float64), params, device=device)
        self._engine = engine
        self._maint: Optional[BatchedAutoMaintenance] = None
        self._elapsed_minutes = 0.0
        self._cache = None
        self.load_demand = 100.0
        self.cooling_water_temp = 25.0
        self.state = _ScalarState(self)
        self.primary_physics = SimpleNamespace(state=self.state, heat_source=_HeatSource(self, params[ixp["rated_power_mw"]]),
                                               rated_power_mw=float(params[ixp["rated_power_mw"]]))
        self.secondary_physics = _ScalarState(self, prefixes=("sec.", "fw.", "sgs.", "turb.", "cond.")) if enable_secondary else None
        if enable_state_management:
            self.state_manager = _StateManagerFacade(self, max_state_rows)
            self.maintenance_system = _MaintenanceFacade(self)
            if isinstance(secondary_config, dict):
                self.state_manager.config = secondary_config
                mode = (secondary_config.get("maintenance_system") or {}).get("maintenance_mode")
                self.maintenance_system.setup_monitoring_from_state_manager(
                    self.state_manager, aggressive_mode=mode in ("aggressive", "ultra_aggressive"))
        else:
            self.state_manager = None
            self.maintenance_system = None

    # -- plumbing -------------------------------------------------------------------------------------------------
    def _row(self) -> np.ndarray:
        if self._cache is None:
            self._cache = self._engine.state_numpy()[self._p]
        return self._cache

    def _write_fields(self, values: Dict[int, float]) -> None:
        self._engine.write_fields(self._p, values)
        self._cache = None

    # -- NuclearPlantSimulator.step: sim.py:130-258 ------------------------------------------------------------------
    def step(self, action=None, magnitude: float = 1.0, load_demand: float = None, cooling_water_temp: float = None) -> Dict:
        ix = field_index()
        if load_demand is not None:
            self.load_demand = load_demand
        if cooling_water_temp is not None:
            self.cooling_water_temp = cooling_water_temp
            self._write_fields({ix["sim.cooling_water_temp"]: float(cooling_water_temp)})
        a = NO_ACTION if action is None else int(getattr(action, "value", action))
        z = np.array([self._heat_rng.standard_normal() if self._heat_rng is not None else 0.0,
                      self._ph_rng.standard_normal(), self._ph_rng.random_sample(), self._ph_rng.random_sample(),
                      self._ph_rng.random_sample()])
        obs, reward, done = self._engine.step_plant(self._p, a, float(magnitude), z)
        self._cache = None
        row = self._row()
        info: Dict[str, Any] = {}
        if self.enable_state_management:
            now = self.state_manager.advance_time(self.dt)
            self._elapsed_minutes = self.state_manager.get_elapsed_time().total_seconds() / 60.0
            info["datetime"] = now.isoformat()
        else:
            self._elapsed_minutes += self.dt
            info["datetime"] = None
        info.update({"time": self._elapsed_minutes, "thermal_power": float(row[ix["pri.thermal_power_mw"]]),
                     "scram_activated": bool(row[ix["pri.scram_activated"]]),
                     "reactivity": float(row[ix["pri.total_reactivity_pcm"]])})
        executed_now = frozenset()
        if self.enable_state_management and self._maint is not None:
            try:    # sim.py:209-223: execute due work orders, then collect states (threshold check)
                executed = self._maint.update(self._elapsed_minutes)
                for wo in executed:
                    if not self._maint.head_quirks:
                        self.state_manager.maintenance_history.append(
                            {"component_id": wo.component_id, "action_type": wo.action, "success": wo.success,
                             "effectiveness": 1.0 if wo.success else 0.0, "timestamp": self._elapsed_minutes})
                if executed:
                    executed_now = frozenset((wo.component_id, wo.action) for wo in executed)
                    info["maintenance_work_orders"] = [vars(wo) for wo in executed]
                    self._cache = None
                    row = self._row()
                self.maintenance_system.current_update_work_orders = [
                    _WorkOrderView(wo) for wo in self._maint.check(self._elapsed_minutes)]
            except NotImplementedError:
                raise
            except Exception as e:     # the reference swallows maintenance failures into warnings (sim.py:214-216)
                warnings.warn(f"Maintenance system update failed: {e}")
        if self.enable_state_management:
            self.state_manager.store.add_row(self.state_manager.current_datetime, row, executed_now)
        if self.enable_secondary:
            g = lambda f, d: float(row[ix[f]]) if np.isfinite(row[ix[f]]) else d
            info.update({"electrical_power": g("sec.electrical_power_output", 0.0),
                         "thermal_efficiency": max(0.0, min(g("sec.thermal_efficiency", 0.0), 0.35)),
                         "steam_flow": g("sec.total_steam_flow", 1665.0), "steam_pressure": g("sec.sg_avg_pressure", 6.895),
                         "condenser_pressure": g("sec.condenser_pressure", 0.007),
                         "condenser_heat_rejection": g("sec.total_system_heat_rejection", 0.0)})
        if not self.enable_secondary:
            obs = obs[:12]          # the reference's observation has the 12 primary entries only (sim.py:292-333)
        return {"observation": obs, "reward": float(reward), "done": bool(done), "info": info}

    def get_observation(self) -> np.ndarray:
        obs = self._engine.observe_plant(self._p)
        return obs if self.enable_secondary else obs[:12]

    def calculate_reward(self, secondary_result: dict = None) -> float:
        """sim.py:500-544 on the current state.  Called by hand it is the primary-side reward unless the caller passes a
        step's secondary result, exactly as in the reference; step() returns the full reward computed on the device."""
        s = self.state
        base = -abs(s.power_level - 100) / 100
        if s.fuel_temperature > 800:
            base += -(s.fuel_temperature - 800) / 100
        else:
            base += 0
        base += -(s.coolant_pressure - 16) if s.coolant_pressure > 16 else 0
        base += -100 if s.scram_status else 0
        if secondary_result is None:
            return base
        sec = (secondary_result["thermal_efficiency"] - 0.30) * 10
        sec += -abs(secondary_result["electrical_power_mw"] - self.load_demand / 100.0 * 1100.0) / 100
        p_sg = secondary_result["sg_avg_pressure"]
        sec += -abs(p_sg - 6.895) * 5 if (p_sg < 5.0 or p_sg > 8.0) else 0
        p_c = secondary_result["condenser_pressure"]
        sec += -(p_c - 0.007) * 100 if p_c > 0.01 else 0
        return base + sec * 0.5

    def reset(self, start_at_steady_state: bool = True) -> np.ndarray:
        self._engine.reset_plant(self._p)
        self._cache = None
        self._elapsed_minutes = 0.0
        self.time = 0.0          # sim.py:573 (the reference, too, only has this attribute after a reset)
        if self.state_manager is not None:
            self.state_manager.clear_data()
            self.state_manager.current_datetime = self.state_manager.start_datetime
        self.load_demand = 100.0
        self.cooling_water_temp = 25.0
        return self.get_observation()
