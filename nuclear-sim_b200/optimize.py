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
