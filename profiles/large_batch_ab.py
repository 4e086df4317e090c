#!/usr/bin/env python
"""A/B of library builds on the config-#3 batch (65,536 plants, 128 substeps per launch, load-following inputs, monitor off)
and on 4,096 / 16,384 plants: plant-steps/s of the library NPS_B200_LIB names."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nuclear_sim_b200 import BatchedNuclearPlantSimulator, load_snapshot  # noqa: E402
from nuclear_sim_b200 import scenarios as sc  # noqa: E402

out = {"lib": os.path.basename(os.environ.get("NPS_B200_LIB", "libnps_b200.so"))}
for n, k, reps in ((65536, 128, 5), (16384, 128, 3), (4096, 128, 3)):
    s0, params = load_snapshot("pwr3000_reactor_dt1")
    pid = np.arange(n)
    sim = BatchedNuclearPlantSimulator(n, sc.randomized_states(s0, pid), params)
    acts, mags = sc.load_following_inputs(pid, 0, k)
    a, m = torch.from_numpy(acts).cuda(), torch.from_numpy(mags).cuda()
    for _ in range(3):
        sim.step(actions=a, magnitudes=m, K=k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sim.step(actions=a, magnitudes=m, K=k)
    e1.record()
    torch.cuda.synchronize()
    out[str(n)] = n * k * reps / (e0.elapsed_time(e1) * 1e-3)
    del sim
print(json.dumps(out))
