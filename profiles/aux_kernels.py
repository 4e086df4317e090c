#!/usr/bin/env python
"""Auxiliary kernels at 65,536 plants: flag kernel (bitmask form), event-list flag kernel, ring-buffer rows, maintenance kernel.
Prints microseconds per call and the fraction of the measured HBM copy peak where a byte count is meaningful."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nuclear_sim_b200 import BatchedNuclearPlantSimulator, load_snapshot  # noqa: E402
from nuclear_sim_b200 import scenarios as sc  # noqa: E402
from nuclear_sim_b200 import maintenance as M  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    n = 65536
    peak = 6458.7
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    s0, params = load_snapshot("pwr3000_oil_top_off_dt5")
    sim = BatchedNuclearPlantSimulator(n, sc.randomized_states(s0, np.arange(n)), params)
    cfg = json.load(open(os.path.join(ROOT, "nuclear-sim_b200", "data", "maintenance_system_template.json")))
    table = M.ThresholdTable(cfg)
    sim.set_thresholds(table.device_rows())
    sim.enable_monitor()
    sim.step(K=2)
    out = {"plants": n, "hbm_peak_gbs": peak}
    us = timed(sim.check_thresholds)
    out["flag_kernel_bitmask_us"] = us
    us = timed(sim.check_thresholds_events)
    sim.drain_step_events()
    live = len(table.bound())
    out["flag_kernel_event_list_us"] = us
    out["flag_kernel_event_list_frac_of_hbm_peak"] = (live + 1) * 8 * n / (us * 1e-6) / 1e9 / peak
    for label, prefix in (("all_633_fields", None), ("one_pump_fields", "secondary.feedwater_FWP-1.")):
        k = sim.set_logged_columns(prefix, ring_rows=4)
        us = timed(sim.log_row)
        out[f"log_row_{label}_us"] = us
        out[f"log_row_{label}_fields"] = k
        out[f"log_row_{label}_frac_of_hbm_peak"] = 2 * k * 8 * n / (us * 1e-6) / 1e9 / peak
    req = np.stack([np.arange(0, n, 12), np.zeros(len(range(0, n, 12)), dtype=np.int64), np.ones(len(range(0, n, 12)), dtype=np.int64),
                    np.zeros(len(range(0, n, 12)), dtype=np.int64)], axis=1).astype(np.int32)      # oil_top_off on FWP-1 of every 12th plant
    us = timed(lambda: sim.apply_maintenance(req), reps=5)
    out["maintenance_kernel_us_incl_copies"] = us
    out["maintenance_requests"] = int(len(req))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
