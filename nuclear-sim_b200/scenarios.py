"""Synthetic batch workloads for the BASELINE configs (host-side numpy; inputs only, no physics).

Per-plant variation is applied on the flat PlantState vector of a committed snapshot: the fields
below are the ones the reference's randomised initial conditions touch
(data_gen/config_engine/initial_conditions/randomization_utils.py; consumed keys listed in
SURVEY.md §8b), perturbed multiplicatively by U(1-factor, 1+factor) with seed = plant id so results
do not depend on how plants are sharded over GPUs.
"""
from __future__ import annotations

import numpy as np

from ._layout import field_index, field_names

ACT_ROD_INSERT, ACT_ROD_WITHDRAW, ACT_NO_ACTION = 0, 1, 8
ACT_INCREASE_FEEDWATER, ACT_DECREASE_FEEDWATER = 6, 7

_IC_PATTERNS = (
    "lub.oil_level", "lub.oil_contamination_level", "lub.oil_moisture_content", "lub.oil_acidity_number",
    "lub.oil_temperature", "lub.component_wear", "cavitation_damage", "tsp_thickness", "tif_scale_thickness",
    "tif_comp", "water_level", "fl_biofouling_thickness", "fl_scale_thickness", "fl_corrosion_product_thickness",
    "deposit_thickness", "efficiency_degradation", "vs_current_air_leakage",
)


def ic_field_ids():
    names = field_names("PlantState")
    return np.array([i for i, n in enumerate(names) if any(p in n for p in _IC_PATTERNS)], dtype=np.int64)


def randomized_states(base_state: np.ndarray, plant_ids: np.ndarray, factor: float = 0.1) -> np.ndarray:
    """[len(plant_ids), n_state] initial states; plant p is a pure function of (base_state, p, factor)."""
    ids = ic_field_ids()
    out = np.tile(np.asarray(base_state, dtype=np.float64), (len(plant_ids), 1))
    # counter-based stream: one Philox-free, order-independent draw per (plant, field)
    pid = np.asarray(plant_ids, dtype=np.uint64)[:, None]
    fid = ids.astype(np.uint64)[None, :]
    x = (pid * np.uint64(0x9E3779B97F4A7C15) + fid * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x ^= x >> np.uint64(30); x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27); x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    u = (x >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    out[:, ids] *= 1.0 + factor * (2.0 * u - 1.0)
    ix = field_index()
    for k in range(4):   # oil level is a percentage
        f = ix[f"fw.pump[{k}].lub.oil_level"]
        out[:, f] = np.minimum(out[:, f], 100.0)
    return out


def load_following_inputs(plant_ids: np.ndarray, t0: int, k: int):
    """Actions/magnitudes for substeps t0..t0+k-1: even plants follow sin(t/100) rod control
    (data/gen_training_data.py:401-406), every 4th plant runs the 20xINSERT / 30xWITHDRAW power ramp
    (tests/test_scenarios.py:56-74), odd plants interleave the (inert) feedwater actions (sim.py:274-286).
    Returns actions [k, n] int8, magnitudes [k, n] f64."""
    pid = np.asarray(plant_ids, dtype=np.int64)
    n = len(pid)
    t = np.arange(t0, t0 + k, dtype=np.int64)[:, None]
    s = np.sin((t + (pid[None, :] % 97)) / 100.0)
    rod = np.where(s > 0.5, ACT_ROD_WITHDRAW, np.where(s < -0.5, ACT_ROD_INSERT, ACT_NO_ACTION))
    u = (t + pid[None, :]) % 200
    ramp = np.where(u < 20, ACT_ROD_INSERT, np.where(u < 50, ACT_ROD_WITHDRAW, ACT_NO_ACTION))
    fw = np.where((t // 10) % 2 == 0, ACT_INCREASE_FEEDWATER, ACT_DECREASE_FEEDWATER)
    kind = pid[None, :] % 4
    act = np.where(kind == 0, ramp, np.where(kind % 2 == 0, rod, np.where(s * s > 0.25, rod, fw)))
    mag = 0.2 + 0.8 * (((pid * 2654435761) % 1000) / 1000.0)
    return act.astype(np.int8), np.broadcast_to(mag[None, :], (k, n)).copy()


def _mix64(x: np.ndarray) -> np.ndarray:
    x = x.copy()
    x ^= x >> np.uint64(30); x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27); x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return x


def noise_inputs(plant_ids: np.ndarray, t0: int, k: int, seed: int = 1000) -> np.ndarray:
    """[k, 5, n] host-supplied random streams (z_heat, z_ph, u0, u1, u2).  Counter-based: the five numbers of
    (plant id, step) are a pure function of (seed, plant id, step), so a plant sees the same draws whatever the batch,
    the launch grouping or the GPU it is sharded to."""
    pid = np.asarray(plant_ids, dtype=np.uint64)[None, None, :]
    t = np.arange(t0, t0 + k, dtype=np.uint64)[:, None, None]
    c = np.arange(7, dtype=np.uint64)[None, :, None]
    with np.errstate(over="ignore"):
        x = _mix64(_mix64(pid * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed)) + t * np.uint64(0xD1B54A32D192ED03) + c)
    u = ((x >> np.uint64(11)).astype(np.float64) + 0.5) / float(1 << 53)      # (0, 1)
    z = np.sqrt(-2.0 * np.log(u[:, 0:4:2])) * np.cos(2.0 * np.pi * u[:, 1:4:2])   # Box-Muller: (u0,u1) -> z_heat, (u2,u3) -> z_ph
    return np.concatenate([z, u[:, 4:7]], axis=1)
