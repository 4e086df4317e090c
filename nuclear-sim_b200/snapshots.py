"""Plant snapshots: (initial PlantState vector, PlantParams vector) pairs extracted once from the
reference's own initialisation code (oracle/make_golden.py) and committed under data/.

The reference builds a plant from nested YAML/dict config through ~5 000 lines of dataclass and
initial-condition code (systems/secondary/*/config.py, */_apply_initial_conditions); that code is
initialisation, not the step path, and is reused as-is on the host when the reference is installed
(see INTEGRATION.md).  On a machine without the reference, the committed snapshots below are the
starting points; per-plant variation is applied on the flat state vector (scenarios.py).
"""
from __future__ import annotations

import os

import numpy as np

from ._layout import field_names

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def load_snapshot(name: str = "pwr3000_oil_top_off_dt5"):
    path = os.path.join(DATA_DIR, name + ".npz")
    z = np.load(path, allow_pickle=False)
    names = tuple(str(s) for s in z["state_names"])
    if names != field_names("PlantState"):
        raise RuntimeError(f"{path} was generated for a different state.h; regenerate with oracle/make_golden.py")
    pnames = tuple(str(s) for s in z["param_names"])
    if pnames != field_names("PlantParams"):
        raise RuntimeError(f"{path} was generated for a different PlantParams; regenerate with oracle/make_golden.py")
    return z["state"].astype(np.float64), z["params"].astype(np.float64)
