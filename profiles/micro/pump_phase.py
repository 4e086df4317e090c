#!/usr/bin/env python
"""Driver for profiles/micro/pump_phase.cu: time the unit-parallel pump phase on the bench's 65,536-plant slab, check it
against the same functions on the host, and put it next to the monolithic step kernel's time per substep."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nuclear_sim_b200 import BatchedNuclearPlantSimulator, load_snapshot, scenarios as sc  # noqa: E402


def main():
    so = os.path.join(ROOT, "profiles", "micro", "libpump_phase.so")
    if not os.path.exists(so):      # the probe library is not part of the product build
        import subprocess
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-fmad=false",
                               "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-shared",
                               "-I", os.path.join(ROOT, "nuclear-sim_b200", "csrc", "plant"), "-o", so,
                               os.path.join(ROOT, "profiles", "micro", "pump_phase.cu")])
    L = ctypes.CDLL(so)
    L.pump_phase_launch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    L.pump_phase_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
    n = 65536
    s0, params = load_snapshot("pwr3000_reactor_dt1")
    pid = np.arange(n)
    sim = BatchedNuclearPlantSimulator(n, sc.randomized_states(s0, pid), params, device="cuda:0")
    for _ in range(2):
        sim.step(K=16)        # realistic mid-run state
    torch.cuda.synchronize()
    params = np.ascontiguousarray(params, dtype=np.float64)
    stream = torch.cuda.current_stream().cuda_stream
    # correctness of the probe kernel: 256 plants, one application, against the host build of the same functions
    ref = sim.state_numpy()[:256].copy()
    L.pump_phase_host(ref.ctypes.data, params.ctypes.data, 256, 1)
    small = sim.slab[:, :256].contiguous()
    L.pump_phase_launch(small.data_ptr(), params.ctypes.data, 256, 1, 1, stream)
    torch.cuda.synchronize()
    got = small.t().cpu().numpy()
    err = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
    print(f"probe kernel vs host functions, 256 plants: max rel err {np.nanmax(np.where(ref != 0, err, 0)):.2e}")
    # the monolithic kernel, per substep
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    sim.step(K=64); torch.cuda.synchronize()
    ev[0].record(); sim.step(K=64); sim.step(K=64); ev[1].record(); torch.cuda.synchronize()
    mono = ev[0].elapsed_time(ev[1]) / 128 * 1e3
    print(f"monolithic step kernel: {mono:.1f} us per substep for {n} plants (pumps ~19 % of samples = {0.19 * mono:.1f} us)")
    work = sim.slab.clone()
    for minb in (1, 2, 3, 4):
        for repeats in (1, 8):
            for _ in range(3):
                L.pump_phase_launch(work.data_ptr(), params.ctypes.data, n, repeats, minb, stream)
            torch.cuda.synchronize()
            ev[0].record()
            for _ in range(20):
                L.pump_phase_launch(work.data_ptr(), params.ctypes.data, n, repeats, minb, stream)
            ev[1].record(); torch.cuda.synchronize()
            us = ev[0].elapsed_time(ev[1]) / 20 * 1e3
            print(f"pump phase kernel, {minb} x 256 threads per SM, {repeats} updates per launch: {us:8.1f} us per launch, "
                  f"{us / repeats:7.1f} us per pump-phase of {n} plants")


if __name__ == "__main__":
    main()
