#!/usr/bin/env python
"""Per-step timeline of the pipelined host-buffer arm (nps_step_host_async): where does a step lose time against the
device-resident arm?  Prints, per step, the host time spent in wait / issue and the device-side interval between
consecutive kernel completions.  python profiles/e2e_timeline.py [steps]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nuclear_sim_b200 import BatchedNuclearPlantSimulator, load_snapshot, scenarios as sc  # noqa: E402


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    n, ksub = 65536, 32
    dev = torch.device("cuda:0")
    s0, params = load_snapshot("pwr3000_reactor_dt1")
    pid = np.arange(n)
    sim = BatchedNuclearPlantSimulator(n, sc.randomized_states(s0, pid), params, device=str(dev))
    acts_h = torch.empty((K, ksub, n), dtype=torch.int8).pin_memory()
    mags_h = torch.empty((K, ksub, n), dtype=torch.float64).pin_memory()
    noise_h = torch.empty((K, ksub, 5, n), dtype=torch.float64).pin_memory()
    for i in range(K):
        a, m = sc.load_following_inputs(pid, i * ksub, ksub)
        acts_h[i] = torch.from_numpy(a); mags_h[i] = torch.from_numpy(m)
        noise_h[i] = torch.from_numpy(sc.noise_inputs(pid, i * ksub, ksub))
    obs_h = [torch.empty((22, n), dtype=torch.float64).pin_memory() for _ in range(sim.pipe_depth)]
    rew_h = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(sim.pipe_depth)]
    done_h = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(sim.pipe_depth)]
    modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["full"]
    for mode in [m for m in modes for _ in range(3)]:
        rep = mode
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        t_wait, t_issue, t_host = [], [], []
        tickets = []
        ev[0].record()
        t00 = time.perf_counter()
        for i in range(K):
            b = i % sim.pipe_depth
            t0 = time.perf_counter()
            if i >= sim.pipe_depth:
                sim.wait(tickets[i - sim.pipe_depth])
                float(rew_h[b].mean())
            t1 = time.perf_counter()
            if mode == "full":
                tickets.append(sim.step_host_async(acts_h[i], mags_h[i], noise_h[i], None, ksub, obs_h[b], rew_h[b], done_h[b]))
            elif mode == "noin":     # 2 MB of actions only: is the stall tied to the 100 MB input copy?
                tickets.append(sim.step_host_async(acts_h[i], None, None, None, ksub, obs_h[b], rew_h[b], done_h[b]))
            elif mode == "noout":    # inputs, but only the reward comes home
                tickets.append(sim.step_host_async(acts_h[i], mags_h[i], noise_h[i], None, ksub, None, rew_h[b], None))
            ev[i + 1].record()
            t2 = time.perf_counter()
            t_wait.append((t1 - t0) * 1e3); t_issue.append((t2 - t1) * 1e3); t_host.append((t0 - t00) * 1e3)
        for i in range(max(0, K - sim.pipe_depth), K):
            sim.wait(tickets[i])
        torch.cuda.synchronize()
        gaps = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
        print(f"rep {rep}: total {ev[0].elapsed_time(ev[K]):.1f} ms for {K} steps ({n * ksub * K / ev[0].elapsed_time(ev[K]) / 1e3:.4g} plant-steps/s)")
        print("  kernel-to-kernel ms:", " ".join(f"{g:.1f}" for g in gaps))
        print("  host wait ms       :", " ".join(f"{g:.1f}" for g in t_wait))
        print("  host issue ms      :", " ".join(f"{g:.1f}" for g in t_issue))


if __name__ == "__main__":
    main()
