#!/usr/bin/env python
"""Small batches: the two-threads-per-plant launch shape (source half / sink half of a plant in two warps of a block)
against one thread per plant, plant-steps/s at several batch sizes, 128 substeps per launch, monitoring off."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nuclear_sim_b200 import BatchedNuclearPlantSimulator, load_snapshot  # noqa: E402
from nuclear_sim_b200 import scenarios as sc  # noqa: E402


def rate(n, shape, k=128):
    s0, params = load_snapshot("pwr3000_reactor_dt1")
    pid = np.arange(n)
    sim = BatchedNuclearPlantSimulator(n, sc.randomized_states(s0, pid), params)
    sim.set_small_batch_shape(shape)
    acts, mags = sc.load_following_inputs(pid, 0, k)
    a, m = torch.from_numpy(acts).cuda(), torch.from_numpy(mags).cuda()
    for _ in range(2):
        sim.step(actions=a, magnitudes=m, K=k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        sim.step(actions=a, magnitudes=m, K=k)
    e1.record()
    torch.cuda.synchronize()
    return n * k * 3 / (e0.elapsed_time(e1) * 1e-3)


def main():
    out = {}
    for n in (1024, 4096, 8192, 16384, 18944):
        out[n] = {"two_threads_per_plant": rate(n, 0), "one_thread_per_plant": rate(n, 1)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
