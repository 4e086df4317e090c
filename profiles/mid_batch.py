#!/usr/bin/env python
"""Mid-size batches: one uncapped-register thread per plant (one-warp blocks) against the 448 x 128-register shape.
NPS_LARGE_BATCH (read at nps_create) moves the switch point; prints plant-steps/s for both shapes at several sizes."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nuclear_sim_b200 import BatchedNuclearPlantSimulator, load_snapshot  # noqa: E402
from nuclear_sim_b200 import scenarios as sc  # noqa: E402


def rate(n, large_batch, k=128):
    os.environ["NPS_LARGE_BATCH"] = str(large_batch)
    s0, params = load_snapshot("pwr3000_reactor_dt1")
    pid = np.arange(n)
    sim = BatchedNuclearPlantSimulator(n, sc.randomized_states(s0, pid), params)
    sim.set_small_batch_shape(1)
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
    for n in (24576, 32768, 36864, 40960, 45056, 49152, 57344):
        out[n] = {"one_warp_blocks_uncapped": rate(n, 10 ** 9), "448x128": rate(n, 1)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
