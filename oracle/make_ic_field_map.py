"""TEST / BUILD INFRASTRUCTURE — which PlantState fields each initial-condition key of the reference's config sets.

    python oracle/make_ic_field_map.py     -> nuclear-sim_b200/data/ic_field_map.json

The reference's optimisers (data_gen/optimization/timing_optimizer.py) vary ONE initial-condition value of a config
dict and build a fresh NuclearPlantSimulator per probe.  For the batched optimiser every probe is one plant of a batch,
so it needs to know where an IC key lands in the flat state vector.  This script asks the reference itself: for every
key `secondary_system.<subsystem>.initial_conditions.<key>` of the composed template config it builds the plant twice
(baseline value, perturbed value — list values are set element-wise to the same number, as
TimingOptimizer._set_config_value does, timing_optimizer.py:355-381) and diffs the extracted PlantState vectors:
  copy     every changed field equals the new value           -> the optimiser writes the fields directly
  none     no field changed (the key is not consumed by the constructors: SURVEY 8b)
  rebuild  fields change in some other way (derived values)    -> needs the reference constructor per probe
"""
from __future__ import annotations

import copy
import json
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
sys.path.insert(0, _REPO)
from oracle import refplant as R  # noqa: E402

OUT = os.path.join(_REPO, "nuclear-sim_b200", "data", "ic_field_map.json")


def main():
    from nuclear_sim_b200.optimize import TimingOptimizer
    L = R._layout()
    names = L.field_names("PlantState")
    cfg = R.compose_config("oil_top_off", duration_hours=4.0)
    to = TimingOptimizer(verbose=False)
    ics = to._extract_initial_conditions(cfg, "oil_top_off")
    build = lambda c: R.extract_state(R.make_reference_plant(c, dt=1.0, heat_source="constant").sim)   # noqa: E731
    base = build(cfg)
    out = {}
    for n, (path, v0) in enumerate(ics.items()):
        v1 = v0 * 1.03125 + 0.015625                     # exactly representable perturbation of ordinary values
        c = copy.deepcopy(cfg)
        to._set_config_value(c, path, v1)
        try:
            st = build(c)
        except Exception as e:                           # the constructors reject the value (type / range)
            out[path] = {"kind": "rebuild", "fields": [], "note": f"{type(e).__name__}"}
            continue
        ch = np.nonzero(~((st == base) | (np.isnan(st) & np.isnan(base))))[0]
        if len(ch) == 0:
            kind = "none"
        elif np.all(st[ch] == v1):
            kind = "copy"
        else:
            kind = "rebuild"
        out[path] = {"kind": kind, "fields": [names[j] for j in ch]}
        if n % 20 == 0:
            print(f"[ic-map] {n}/{len(ics)} {path}: {kind} {len(ch)} fields", flush=True)
    kinds = {k: sum(1 for v in out.values() if v["kind"] == k) for k in ("copy", "none", "rebuild")}
    with open(OUT, "w") as fh:
        json.dump({"generated_by": "oracle/make_ic_field_map.py (live reference)", "base_action": "oil_top_off",
                   "counts": kinds, "map": out}, fh, indent=1, sort_keys=True)
    print(f"[ic-map] {len(out)} keys: {kinds} -> {OUT}")


if __name__ == "__main__":
    if not R.reference_available():
        sys.exit("reference not found")
    main()
