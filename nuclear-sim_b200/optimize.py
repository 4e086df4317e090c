"""Batched trigger-time sweeps — the reference's ``TimingOptimizer`` (data_gen/optimization/timing_optimizer.py:24-381)
and ``ICOptimizer.batch_optimize_timing`` (data_gen/optimization/ic_optimizer.py:162-207) re-expressed for the batched engine.

The reference searches for an initial-condition value that makes a maintenance action trigger at a target time by
BISECTION, building and running one fresh ``NuclearPlantSimulator`` per probe (timing_optimizer.py:288).  With N plants
in one batch every candidate value is simply one plant: one sweep answers the whole question, and a second, narrower
sweep refines it to the timestep.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from ._layout import field_index
from .maintenance import BatchedAutoMaintenance, ThresholdTable


def _default_engine(states, params, device):
    from .batched import BatchedNuclearPlantSimulator
    return BatchedNuclearPlantSimulator(states.shape[0], states, params, device=device)


def trigger_time_sweep(base_state: np.ndarray, params: np.ndarray, maintenance_config: dict, field: str, values,
                       target_action: str, horizon_hours: float, component_id: Optional[str] = None, device: str = "cuda:0",
                       engine_factory: Callable = _default_engine) -> np.ndarray:
    """Hours until the first work order for `target_action` (on `component_id` if given) is created, for each candidate
    value of PlantState `field`; NaN where nothing triggers within the horizon.  Every candidate is one plant of ONE batch
    (the reference's _test_trigger_timing, timing_optimizer.py:262-330, once per candidate)."""
    values = np.asarray(values, dtype=np.float64)
    ix = field_index()
    dt = float(params[field_index("PlantParams")["dt"]])
    states = np.tile(np.asarray(base_state, dtype=np.float64), (len(values), 1))
    states[:, ix[field]] = values
    sim = engine_factory(states, np.asarray(params, dtype=np.float64), device)
    maint = BatchedAutoMaintenance(sim, ThresholdTable(maintenance_config), aggressive=True)
    out = np.full(len(values), np.nan)
    steps = int(round(horizon_hours * 60.0 / dt))
    seen = 0
    for t in range(steps):
        sim.step(K=1) if hasattr(sim, "slab") else sim.step()
        now = (t + 1) * dt
        maint.update(now)
        maint.check(now)
        for wo in maint.created_log[seen:]:
            if wo.action == target_action and (component_id is None or wo.component_id == component_id) and np.isnan(out[wo.plant]):
                out[wo.plant] = wo.created / 60.0
        seen = len(maint.created_log)
        if not np.isnan(out).any():
            break
    return out


def optimize_for_target_timing(base_state, params, maintenance_config, field: str, lo: float, hi: float, target_action: str,
                               target_trigger_hours: float, tolerance_hours: float = 0.1, component_id: Optional[str] = None,
                               n_candidates: int = 128, max_sweeps: int = 3, device: str = "cuda:0",
                               engine_factory: Callable = _default_engine) -> Tuple[float, Optional[float], int]:
    """(best value, achieved trigger hours, sweeps used): TimingOptimizer.optimize_for_target_timing
    (timing_optimizer.py:38-140) as at most `max_sweeps` batched sweeps over [lo, hi]."""
    best_v, best_t = float("nan"), None
    for sweep in range(1, max_sweeps + 1):
        cand = np.linspace(lo, hi, n_candidates)
        hours = trigger_time_sweep(base_state, params, maintenance_config, field, cand, target_action,
                                   horizon_hours=target_trigger_hours * 2.0, component_id=component_id, device=device,
                                   engine_factory=engine_factory)
        ok = ~np.isnan(hours)
        if not ok.any():
            return best_v, best_t, sweep
        err = np.where(ok, np.abs(hours - target_trigger_hours), np.inf)
        j = int(np.argmin(err))
        if best_t is None or err[j] < abs(best_t - target_trigger_hours):
            best_v, best_t = float(cand[j]), float(hours[j])
        if err[j] <= tolerance_hours:
            return best_v, best_t, sweep
        step = (hi - lo) / (n_candidates - 1)
        lo, hi = cand[j] - step, cand[j] + step      # zoom in around the best candidate
    return best_v, best_t, max_sweeps


def trigger_time_sweep_multi(base_state: np.ndarray, params: np.ndarray, maintenance_config: dict,
                             probes: Sequence[Tuple[str, Sequence[float], str, Optional[str]]], horizon_hours: float,
                             device: str = "cuda:0", engine_factory: Callable = _default_engine) -> List[np.ndarray]:
    """Several sweeps in ONE batch.  probes[i] = (PlantState field, candidate values, target action, component id or
    None); the plants of probe i are one contiguous group of the batch.  Returns, per probe, the hours until the first
    work order for its action (NaN: nothing within the horizon)."""
    ix = field_index()
    dt = float(params[field_index("PlantParams")["dt"]])
    groups, rows = [], []
    for field, values, action, comp in probes:
        v = np.asarray(values, dtype=np.float64)
        st = np.tile(np.asarray(base_state, dtype=np.float64), (len(v), 1))
        st[:, ix[field]] = v
        groups.append((len(rows), len(rows) + len(v), action, comp))
        rows.extend(st)
    states = np.asarray(rows)
    owner = np.empty(len(states), dtype=np.int64)
    for g, (lo, hi, _, _) in enumerate(groups):
        owner[lo:hi] = g
    sim = engine_factory(states, np.asarray(params, dtype=np.float64), device)
    maint = BatchedAutoMaintenance(sim, ThresholdTable(maintenance_config), aggressive=True)
    out = np.full(len(states), np.nan)
    seen = 0
    for t in range(int(round(horizon_hours * 60.0 / dt))):
        sim.step(K=1) if hasattr(sim, "slab") else sim.step()
        now = (t + 1) * dt
        maint.update(now)
        maint.check(now)
        for wo in maint.created_log[seen:]:
            _, _, action, comp = groups[owner[wo.plant]]
            if wo.action == action and (comp is None or wo.component_id == comp) and np.isnan(out[wo.plant]):
                out[wo.plant] = wo.created / 60.0
        seen = len(maint.created_log)
        if not np.isnan(out).any():
            break
    return [out[lo:hi].copy() for lo, hi, _, _ in groups]


def batch_optimize_timing(base_state, params, maintenance_config,
                          targets: Dict[str, Tuple[str, float, float, float, Optional[str]]], tolerance_hours: float = 0.1,
                          n_candidates: int = 64, max_sweeps: int = 3, device: str = "cuda:0",
                          engine_factory: Callable = _default_engine) -> Dict[str, Tuple[float, Optional[float], int]]:
    """ICOptimizer.batch_optimize_timing (ic_optimizer.py:162-207): several actions, each with its own target trigger
    time.  targets[action] = (PlantState field, lo, hi, target hours, component id or None).  The reference optimises
    the actions one after the other, one simulator per bisection probe; here every candidate of EVERY action is one plant
    of one batch, and each refinement round is one more batch over the actions that have not converged yet.
    Returns action -> (best value, achieved trigger hours or None, sweeps used)."""
    box = {a: [lo, hi] for a, (_, lo, hi, _, _) in targets.items()}
    best: Dict[str, Tuple[float, Optional[float], int]] = {a: (float("nan"), None, 0) for a in targets}
    todo = list(targets)
    for sweep in range(1, max_sweeps + 1):
        if not todo:
            break
        cands = {a: np.linspace(box[a][0], box[a][1], n_candidates) for a in todo}
        horizon = 2.0 * max(targets[a][3] for a in todo)
        hours = trigger_time_sweep_multi(base_state, params, maintenance_config,
                                         [(targets[a][0], cands[a], a, targets[a][4]) for a in todo], horizon,
                                         device=device, engine_factory=engine_factory)
        nxt = []
        for a, h in zip(todo, hours):
            want = targets[a][3]
            ok = ~np.isnan(h)
            if not ok.any():
                best[a] = (best[a][0], best[a][1], sweep)
                continue
            err = np.where(ok, np.abs(h - want), np.inf)
            j = int(np.argmin(err))
            if best[a][1] is None or err[j] < abs(best[a][1] - want):
                best[a] = (float(cands[a][j]), float(h[j]), sweep)
            else:
                best[a] = (best[a][0], best[a][1], sweep)
            if err[j] > tolerance_hours:
                step = (box[a][1] - box[a][0]) / (n_candidates - 1)
                box[a] = [cands[a][j] - step, cands[a][j] + step]
                nxt.append(a)
        todo = nxt
    return best


# ---------------------------------------------------------------------------------------------------------------
# the reference's class, same interface: TimingOptimizer(verbose).optimize_for_target_timing(base_config, ...)
# ---------------------------------------------------------------------------------------------------------------
def load_ic_field_map() -> Dict[str, dict]:
    """data/ic_field_map.json (oracle/make_ic_field_map.py): IC key of the reference's config -> PlantState fields."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "ic_field_map.json")
    with open(path) as fh:
        return json.load(fh)["map"]


class TimingOptimizer:
    """Drop-in for data_gen/optimization/timing_optimizer.py:22-381: same constructor argument, same
    ``optimize_for_target_timing(base_config, target_action, target_trigger_hours, tolerance_hours, max_iterations)``
    -> ``(optimized_config, achieved_trigger_hours, iterations_used)``, same ``param_path`` handling
    (``secondary_system.<subsystem>.initial_conditions.<key>``; list values probed through their first element and set
    element-wise), same bounds / direction heuristics, same bisection decisions.

    What changes is how a bisection is evaluated.  The reference builds and runs one fresh NuclearPlantSimulator per
    probe (timing_optimizer.py:262-330) — ten sequential simulations per parameter.  The test values of a bisection are
    the nodes of a binary tree that is known in advance (each node's value is (lo + hi) / 2 of its parent's interval),
    so ALL nodes down to `max_iterations` levels are evaluated as plants of ONE batch and the reference's decision path
    is then walked through the finished tree: identical values, identical decisions, identical return tuple, one
    launch sequence instead of ten simulator constructions (tests/test_timing_optimizer_class.py replays the
    reference's own class against it).

    base_state / params are the plant the config describes (built by the reference's constructors);
    an IC key is applied to the state vector through data/ic_field_map.json (which fields the reference's constructors
    write for that key).  Keys the constructors ignore ("none") cannot change the trigger time and are skipped after
    one shared baseline probe; keys with derived effects ("rebuild") go through `plant_builder(config) -> state` when
    one is given and are reported as unavailable otherwise.
    """

    def __init__(self, verbose: bool = True, base_state: Optional[np.ndarray] = None, params: Optional[np.ndarray] = None,
                 plant_builder: Optional[Callable] = None, ic_field_map: Optional[Dict[str, dict]] = None,
                 component_id: Optional[str] = None, device: str = "cuda:0", engine_factory: Callable = _default_engine,
                 probe: Optional[Callable] = None):
        self.verbose = verbose
        self.base_state, self.params = base_state, params
        self.plant_builder = plant_builder
        self.ic_map = ic_field_map
        self.component_id = component_id
        self.device, self.engine_factory = device, engine_factory
        self._probe = probe                   # probe(configs, target_action, max_hours) -> [hours or None]; tests inject one
        self.n_batches = 0

    # -- the reference's helpers, restated ----------------------------------------------------------------------------
    def _extract_initial_conditions(self, config: Dict, target_action: str) -> Dict[str, float]:
        """timing_optimizer.py:332-362"""
        out: Dict[str, float] = {}
        for subsystem, sub in (config.get("secondary_system") or {}).items():
            if isinstance(sub, dict) and "initial_conditions" in sub:
                for key, value in sub["initial_conditions"].items():
                    path = f"secondary_system.{subsystem}.initial_conditions.{key}"
                    if isinstance(value, list) and value:
                        out[path] = float(value[0])
                    elif isinstance(value, (int, float)):
                        out[path] = float(value)
        return out

    def _set_config_value(self, config: Dict, param_path: str, value: float) -> None:
        """timing_optimizer.py:364-392"""
        parts = param_path.split(".")
        cur = config
        for part in parts[:-1]:
            if part not in cur:
                cur[part] = {}
            cur = cur[part]
        key = parts[-1]
        if isinstance(cur.get(key), list):
            cur[key] = [value] * len(cur[key])
        else:
            cur[key] = value

    def _get_parameter_bounds(self, param_path: str, baseline_value: float) -> Tuple[float, float]:
        """timing_optimizer.py:196-236"""
        p = param_path.lower()
        if "oil_level" in p:
            return (10.0, min(100.0, baseline_value * 1.5))
        if "temperature" in p:
            return (max(0.0, baseline_value * 0.5), baseline_value * 2.0)
        if "vibration" in p:
            return (0.0, baseline_value * 3.0)
        if "contamination" in p:
            return (0.0, baseline_value * 5.0)
        if "fouling" in p:
            return (0.0, baseline_value * 4.0)
        if "efficiency" in p:
            return (max(0.1, baseline_value * 0.5), min(1.0, baseline_value * 1.2))
        return (max(0.0, baseline_value * 0.3), baseline_value * 3.0)

    def _parameter_increases_degradation(self, param_path: str) -> bool:
        """timing_optimizer.py:238-270"""
        p = param_path.lower()
        for k in ("contamination", "fouling", "vibration", "temperature", "wear", "corrosion", "scale", "deposits"):
            if k in p:
                return True
        for k in ("oil_level", "efficiency", "performance", "quality", "pressure", "flow"):
            if k in p:
                return False
        return True

    # -- probes ---------------------------------------------------------------------------------------------------------
    def _test_trigger_timing_batch(self, configs: Sequence[Dict], target_action: str, max_simulation_hours: float):
        """_test_trigger_timing (timing_optimizer.py:272-330) for many configs at once -> [hours or None]."""
        self.n_batches += 1
        if self._probe is not None:
            return list(self._probe(list(configs), target_action, max_simulation_hours))
        if self.base_state is None or self.params is None:
            raise ValueError("TimingOptimizer needs base_state / params (the plant base_config describes) or a probe")
        ic_map = self.ic_map if self.ic_map is not None else load_ic_field_map()
        ix = field_index()
        base_ics = self._extract_initial_conditions(configs[0], target_action) if configs else {}
        states = np.tile(np.asarray(self.base_state, dtype=np.float64), (len(configs), 1))
        for i, cfg in enumerate(configs):
            for path, v in self._extract_initial_conditions(cfg, target_action).items():
                if i > 0 and base_ics.get(path) == v:
                    continue                                       # unchanged relative to the first (baseline) config
                m = ic_map.get(path, {"kind": "none"})
                if m["kind"] == "copy":
                    if i > 0 or self.plant_builder is None:
                        for f in m["fields"]:
                            states[i, ix[f]] = v
                elif m["kind"] == "rebuild" and i > 0:
                    if self.plant_builder is None:
                        raise NotImplementedError(f"{path}: derived initial condition, needs plant_builder(config)")
                    states[i] = self.plant_builder(cfg)
        hours = trigger_time_sweep_states(states, self.params, configs[0].get("maintenance_system", {}), target_action,
                                          max_simulation_hours, self.component_id, self.device, self.engine_factory)
        return [None if np.isnan(h) else float(h) for h in hours]

    def _test_trigger_timing(self, config: Dict, target_action: str, max_simulation_hours: float) -> Optional[float]:
        return self._test_trigger_timing_batch([config], target_action, max_simulation_hours)[0]

    # -- bisection: whole tree in one batch, then the reference's walk --------------------------------------------------
    def _binary_search_parameter(self, base_config: Dict, target_action: str, param_path: str, baseline_value: float,
                                 target_trigger_hours: float, tolerance_hours: float,
                                 max_iterations: int) -> Tuple[float, Optional[float], int]:
        """timing_optimizer.py:121-194"""
        import copy
        lo0, hi0 = self._get_parameter_bounds(param_path, baseline_value)
        # every interval the walk can reach, level by level: node -> (lo, hi); children share the parent's midpoint
        nodes = {(): (lo0, hi0)}
        order = [()]
        for depth in range(max_iterations - 1):
            for path in [q for q in order if len(q) == depth]:
                lo, hi = nodes[path]
                mid = (lo + hi) / 2
                for bit, iv in ((0, (lo, mid)), (1, (mid, hi))):
                    if abs(iv[1] - iv[0]) < baseline_value * 0.001:      # the walk stops before probing this interval
                        continue
                    nodes[path + (bit,)] = iv
                    order.append(path + (bit,))
        values = {q: (nodes[q][0] + nodes[q][1]) / 2 for q in order}
        uniq = sorted(set(values.values()))                              # equal test values are one plant
        configs = []
        for v in uniq:
            c = copy.deepcopy(base_config)
            self._set_config_value(c, param_path, v)
            configs.append(c)
        by_value = dict(zip(uniq, self._test_trigger_timing_batch(configs, target_action, target_trigger_hours * 2)))
        times = {q: by_value[values[q]] for q in order}
        inc = self._parameter_increases_degradation(param_path)
        best_value, best_time, best_error = baseline_value, None, float("inf")
        min_v, max_v = lo0, hi0
        q = ()
        for iteration in range(max_iterations):
            test_value = (min_v + max_v) / 2
            assert test_value == values[q]
            t = times[q]
            go_up: bool                                   # True: min_value = test_value, False: max_value = test_value
            if t:
                err = abs(t - target_trigger_hours)
                if err < best_error:
                    best_error, best_value, best_time = err, test_value, t
                if err <= tolerance_hours:
                    return test_value, t, iteration + 1
                go_up = (not inc) if t < target_trigger_hours else inc
            else:
                go_up = inc
            if go_up:
                min_v = test_value
            else:
                max_v = test_value
            if abs(max_v - min_v) < baseline_value * 0.001:
                break
            q = q + (1 if go_up else 0,)
            if q not in values:
                break
        return best_value, best_time, max_iterations

    def optimize_for_target_timing(self, base_config: Dict, target_action: str, target_trigger_hours: float,
                                   tolerance_hours: float = 0.1, max_iterations: int = 10):
        """timing_optimizer.py:38-119"""
        import copy
        baseline_ics = self._extract_initial_conditions(base_config, target_action)
        if not baseline_ics:
            return base_config, None, 0
        baseline = self._test_trigger_timing(base_config, target_action, target_trigger_hours * 2)
        if self.verbose:
            print(f"   baseline trigger time: {baseline}")
        if baseline and abs(baseline - target_trigger_hours) <= tolerance_hours:
            return base_config, baseline, 0
        best_config = copy.deepcopy(base_config)
        best_time, best_error, used = baseline, float("inf"), 0
        for param_path, baseline_value in baseline_ics.items():
            value, t, its = self._binary_search_parameter(base_config, target_action, param_path, baseline_value,
                                                          target_trigger_hours, tolerance_hours, max_iterations)
            used += its
            if t:
                err = abs(t - target_trigger_hours)
                if err < best_error:
                    best_error, best_time = err, t
                    self._set_config_value(best_config, param_path, value)
                    if self.verbose:
                        print(f"   {param_path}: {value:.3f} -> {t:.3f} h (error {err:.3f} h)")
                    if err <= tolerance_hours:
                        break
        return best_config, best_time, used


def trigger_time_sweep_states(states: np.ndarray, params: np.ndarray, maintenance_config: dict, target_action: str,
                              horizon_hours: float, component_id: Optional[str] = None, device: str = "cuda:0",
                              engine_factory: Callable = _default_engine) -> np.ndarray:
    """Hours until the first work order for `target_action` is created, per row of `states` (NaN: none within the
    horizon).  One batch; launches are cut at the maintenance gate steps (BatchedAutoMaintenance.advance)."""
    dt = float(params[field_index("PlantParams")["dt"]])
    sim = engine_factory(np.asarray(states, dtype=np.float64), np.asarray(params, dtype=np.float64), device)
    maint = BatchedAutoMaintenance(sim, ThresholdTable(maintenance_config), aggressive=True)
    out = np.full(len(states), np.nan)
    steps = int(round(horizon_hours * 60.0 / dt))
    seen, done = 0, 0
    chunk = max(1, int(round(15.0 / dt)))
    while done < steps:
        k = min(chunk, steps - done)
        maint.advance(k)
        done += k
        for wo in maint.created_log[seen:]:
            if wo.action == target_action and (component_id is None or wo.component_id == component_id) and np.isnan(out[wo.plant]):
                out[wo.plant] = wo.created / 60.0
        seen = len(maint.created_log)
        if not np.isnan(out).any():
            break
    return out
