#!/usr/bin/env python
"""Time the two auxiliary hot-path kernels against the HBM roofline (CUDA events on the launching stream):
  nps_threshold_kernel — the reference's 331-row maintenance threshold table (90 live rows), 65,536 plants
  nps_log_row_kernel   — trajectory ring-buffer row: every exportable reference column (701 -> distinct fields) and a
                         32-field subset
Prints one JSON line per kernel: algorithmic bytes, time, GB/s, fraction of MEASURED_PEAKS.json hbm_gbs."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nuclear_sim_b200 import BatchedNuclearPlantSimulator, load_snapshot, field_names  # noqa: E402
from nuclear_sim_b200 import scenarios as sc  # noqa: E402
from nuclear_sim_b200.maintenance import ThresholdTable  # noqa: E402
from nuclear_sim_b200.export import ColumnSchema  # noqa: E402


def timed(fn, reps=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > L2 (126 MB)
    tot = 0.0
    for _ in range(reps):
        flush.fill_(1)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e-3


def main():
    n = 65536
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    s0, params = load_snapshot("pwr3000_oil_top_off_dt5")
    sim = BatchedNuclearPlantSimulator(n, sc.randomized_states(s0, np.arange(n)), params)
    g = np.load(os.path.join(ROOT, "tests", "golden", "maint_oil_top_off.npz"), allow_pickle=False)
    tab = ThresholdTable(json.loads(str(g["log"]))["maintenance_system"])
    sim.set_thresholds(tab.device_rows())
    live = len(tab.bound())
    words = (len(tab) + 31) // 32
    t = timed(sim.check_thresholds)
    b = n * (live * 8 * 2 + 8 + words * 4) + (n // 32) * 4      # values + cooldown stamps + clock + flag words + ballot words
    print(json.dumps({"kernel": "nps_threshold_kernel", "plants": n, "threshold_rows": len(tab), "live_rows": live,
                      "algorithmic_bytes": b, "ms": t * 1e3, "GB/s": b / t / 1e9, "frac_of_hbm_peak": b / t / 1e9 / peak}))
    schema = ColumnSchema()
    names = field_names()
    for label, ids in (("all exportable columns", schema.logged_fields(schema.select())), ("32 fields", list(range(0, 1280, 40)))):
        sim.set_logged_fields([names[i] for i in ids], ring_rows=4)
        t = timed(sim.log_row)
        b = n * len(ids) * 8 * 2
        print(json.dumps({"kernel": "nps_log_row_kernel", "what": label, "plants": n, "fields": len(ids), "algorithmic_bytes": b,
                          "ms": t * 1e3, "GB/s": b / t / 1e9, "frac_of_hbm_peak": b / t / 1e9 / peak}))


if __name__ == "__main__":
    main()
