#!/usr/bin/env python
"""What each part of the in-launch monitor costs: 65,536 plants, 128 substeps per launch, device-resident inputs.
Prints ms per launch for: no monitor / stamps only / watch list only / thresholds only / per-substep reward+done only / all."""
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
from nuclear_sim_b200.batched import DEFAULT_WATCH  # noqa: E402


def main():
    n, k = 65536, 128
    s0, params = load_snapshot("pwr3000_reactor_dt1")
    pid = np.arange(n)
    acts, mags = sc.load_following_inputs(pid, 0, k)
    a, m = torch.from_numpy(acts).cuda(), torch.from_numpy(mags).cuda()
    z = torch.from_numpy(np.concatenate([sc.noise_inputs(pid, j, 16) for j in range(0, k, 16)])).cuda()
    cfg = json.load(open(os.path.join(ROOT, "nuclear-sim_b200", "data", "maintenance_system_template.json")))
    rows = M.ThresholdTable(cfg).device_rows()
    out = {}
    for name, thr, watch, per in (("none", False, None, False), ("stamps_only", False, [], False), ("watch_only", False, DEFAULT_WATCH, False),
                                  ("thresholds_only", True, [], False), ("per_substep_only", False, [], True),
                                  ("all", True, DEFAULT_WATCH, True)):
        sim = BatchedNuclearPlantSimulator(n, sc.randomized_states(s0, pid), params)
        if thr:
            sim.set_thresholds(rows)
        if watch is not None:
            sim.enable_monitor(watch=watch, per_substep=per, max_k=k)
        for _ in range(2):
            sim.step(actions=a, magnitudes=m, noise=z, K=k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            sim.step(actions=a, magnitudes=m, noise=z, K=k)
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / 3
        del sim
        torch.cuda.empty_cache()
    base = out["none"]
    print(json.dumps({"ms_per_launch": out, "overhead_pct": {q: 100 * (v / base - 1) for q, v in out.items()}}))


if __name__ == "__main__":
    main()
