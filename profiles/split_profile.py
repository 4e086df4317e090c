#!/usr/bin/env python
"""Workload for the ncu capture of the two-threads-per-plant kernel: 4,096 and 16,384 plants, 32 substeps per launch,
two launches each (the first warms the caches)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nuclear_sim_b200 import BatchedNuclearPlantSimulator, load_snapshot  # noqa: E402
from nuclear_sim_b200 import scenarios as sc  # noqa: E402

for n in (4096, 16384):
    s0, params = load_snapshot("pwr3000_reactor_dt1")
    pid = np.arange(n)
    sim = BatchedNuclearPlantSimulator(n, sc.randomized_states(s0, pid), params)
    acts, mags = sc.load_following_inputs(pid, 0, 32)
    a, m = torch.from_numpy(acts).cuda(), torch.from_numpy(mags).cuda()
    for _ in range(2):
        sim.step(actions=a, magnitudes=m, K=32)
    torch.cuda.synchronize()
    print("n", n, flush=True)
