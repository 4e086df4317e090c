#!/usr/bin/env python
"""BASELINE config #5 on one GPU: long-horizon maintenance degradation with the full loop
(step kernel -> due work orders applied on the device -> flag kernel -> host drain -> work orders).
dt = 5 min, one launch = 3 fused substeps = the 15-minute maintenance gate (auto_maintenance.py:74,213-217), 24 h.
Prints one JSON line: plant-steps/s for the whole loop, time split device / host, event counts."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nuclear_sim_b200 import BatchedNuclearPlantSimulator, load_snapshot, field_index  # noqa: E402
from nuclear_sim_b200 import scenarios as sc  # noqa: E402
from nuclear_sim_b200.maintenance import BatchedAutoMaintenance, ThresholdTable  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--plants", type=int, default=131072)
    ap.add_argument("--hours", type=float, default=24.0)
    args = ap.parse_args()
    n, k, dt = args.plants, 3, 5.0
    s0, params = load_snapshot("pwr3000_oil_top_off_dt5")
    ix = field_index()
    pid = np.arange(n)
    st = sc.randomized_states(s0, pid)
    rng = np.random.RandomState(7)
    # initial conditions positioned near maintenance thresholds: oil levels just above 58 %,
    # TSP fouling / tube scale near their triggers, contamination near 15.2 ppm
    for p in range(4):
        st[:, ix[f"fw.pump[{p}].lub.oil_level"]] = 58.0 + rng.uniform(0.0, 6.0, n)
        st[:, ix[f"fw.pump[{p}].lub.oil_contamination_level"]] = 15.2 - rng.uniform(0.0, 0.6, n)
    g = np.load(os.path.join(ROOT, "tests", "golden", "maint_oil_top_off.npz"), allow_pickle=False)
    cfg = json.loads(str(g["log"]))["maintenance_system"]
    sim = BatchedNuclearPlantSimulator(n, st, params)
    maint = BatchedAutoMaintenance(sim, ThresholdTable(cfg), aggressive=True)
    launches = int(args.hours * 60 / (k * dt))
    torch.cuda.synchronize()
    t_dev = t_host = 0.0
    t0 = time.perf_counter()
    for i in range(launches):
        a = time.perf_counter()
        sim.step(K=k)
        torch.cuda.synchronize()
        b = time.perf_counter()
        now = (i + 1) * k * dt
        maint.update(now)
        maint.check(now)
        c = time.perf_counter()
        t_dev += b - a
        t_host += c - b
    total = time.perf_counter() - t0
    by_action = {}
    for wo in maint.created_log:
        by_action[wo.action] = by_action.get(wo.action, 0) + 1
    print(json.dumps({"workload": "cfg5: long-horizon maintenance degradation, dt=5 min, 3 substeps per launch (15-min gate)",
                      "plants": n, "simulated_hours": args.hours, "launches": launches, "plant_steps": n * k * launches,
                      "plant_steps_per_s_whole_loop": n * k * launches / total, "seconds_total": total,
                      "seconds_step_kernel": t_dev, "seconds_flag_kernel_drain_workorders_effects": t_host,
                      "events": len(maint.event_log), "work_orders_created": len(maint.created_log),
                      "work_orders_executed": len(maint.executed_log), "by_action": by_action}))


if __name__ == "__main__":
    main()
