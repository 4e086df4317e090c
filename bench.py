#!/usr/bin/env python
"""bench.py — plant-steps/sec of the batched plant-dynamics hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE config #3 — 65,536 plants PER GPU with randomised initial conditions,
ReactorHeatSource at equilibrium, load-following / power-ramp rod actions plus the (inert) feedwater actions, dt = 1.0.
One bench "step" = ONE launch of the fused step kernel advancing every plant by SUBSTEPS timesteps, i.e.
plants x SUBSTEPS plant-steps.  Plants shard over ranks (nuclear_sim_b200.sharded.ShardedBatchedSimulator) with no
data-path collective (weak scaling); an NCCL all-gather of trajectory summaries runs after the timed region.

Keys of the JSON line (all plant-steps/s unless noted):
  value            device-resident inputs, physics only (CUDA events, max over ranks)
  full_step        what the reference's step() does besides the physics: maintenance thresholds evaluated after EVERY
                   substep inside the launch (331-row reference table, event list), trip / scram step stamps, per-substep
                   reward + done, and one full trajectory ring row (all exportable fields) per launch
  e2e              through the host-buffer C-ABI call (nps_step_host_async): pinned host inputs copied in and
                   observation / reward / done copied out and read on the host inside the timed region, every step
  value_at_power   control: BASELINE config #2 plants (constant heat source, steady 100 %, NO_ACTION), same plant count
  strong_65536     N > 1: the 65,536-plant batch split over the N GPUs (strong scaling point of the north star)
  small_batch      N = 1: 4,096 (config #2 size) and 16,384 (config #4 size) plants on one GPU
  cfg5_maintenance_loop   BASELINE config #5 end to end: 131,072 plants per GPU, 24 h at dt = 5 min, thresholds inside the
                   launches, work orders through the native work-order table and the maintenance kernel; each GPU's plants
                   as two independent batches whose host work and launches overlap (wall clock, max over ranks)
  roofline         SURVEY.md 8(d): frac = max(rate x 2e4 flop / measured FP64 peak, rate x (26,240 / K + 6,312 x logged
                   fraction) B / measured HBM peak), per GPU; the implementation's own traffic figures are named extras
  cpu_baseline     the host C restatement of the reference step on ALL host cores (count stated), bounded sample
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PLANTS_PER_GPU = 65536
INPUT_SETS = 4
SUBSTEPS = 128                       # fused substeps per launch (profiles/r01_tuning_variants.txt (20): 32 / 64 / 128)
FLOP_PER_PLANT_STEP = 2.0e4          # SURVEY.md 8(d) canonical figures
BYTES_PER_PLANT_STEP_K1 = 26240.0
BYTES_PER_LOGGED_ROW = 6312.0
METRIC = "plant-steps/sec"
UNIT = "plant-steps/s"
WORKLOAD = ("cfg3: 65,536 plants per GPU, randomized ICs, reactor heat source, load-following/power-ramp "
            f"rod + feedwater actions, dt=1.0, {SUBSTEPS} fused substeps per launch")
TRAFFIC_JSON = os.path.join(ROOT, "profiles", "r02_step_kernel_traffic.json")


def _n_live_fields():
    path = os.path.join(ROOT, "nuclear-sim_b200", "csrc", "plant", "live_fields.txt")
    return sum(1 for ln in open(path) if ln.strip() and not ln.startswith("#"))


def _capture():
    """The committed ncu capture of the CURRENT build's step kernel (profiles/r02_step_kernel_traffic.json)."""
    try:
        return json.load(open(TRAFFIC_JSON))
    except Exception:
        return None


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.tmp,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.strip().split(", ") for r in open(self.tmp.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.tmp.name)
        sm, reasons, mx = [], set(), None
        for r in rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the host C restatement of the reference step (oracle/cpu_port.cpp) on every host core
# ----------------------------------------------------------------------------------------------------------------------
def _cpu_proc(lo, hi, k, n_steps, barrier):
    """One host core's share of the CPU arm: plants [lo, hi) of the cfg3 workload, stepped by the host C restatement."""
    from nuclear_sim_b200 import load_snapshot
    from nuclear_sim_b200 import scenarios as sc
    from tests import _util as U
    L = U.oracle_lib()
    s0, params = load_snapshot("pwr3000_reactor_dt1")
    params = np.ascontiguousarray(params)
    pid = np.arange(lo, hi)
    st = np.ascontiguousarray(sc.randomized_states(s0, pid))
    acts, mags = sc.load_following_inputs(pid, 0, k)
    noise = sc.noise_inputs(pid, 0, k)
    a = np.ascontiguousarray(acts.T); m = np.ascontiguousarray(mags.T)
    z = np.ascontiguousarray(noise.transpose(2, 0, 1))      # [n, k, 5]
    for _ in range(n_steps):
        barrier.wait()
        L.nps_oracle_step(U.ptr(st), U.ptr(params), U.ptr(a), U.ptr(m), U.ptr(z), ctypes.c_int64(hi - lo), int(k))
        barrier.wait()


class CpuArm:
    """`procs` worker processes (one per host core; processes rather than threads so the figure does not depend on
    how the host schedules the threads of one process), each owning a contiguous share of the plants.  step() = one
    pass of k substeps over all plants, timed on the wall clock between two barriers."""

    def __init__(self, n_plants: int, k: int, procs: int, n_steps: int):
        import multiprocessing as mp
        ctx = mp.get_context("spawn")            # the GPU arm calls this with a live CUDA context: never fork it
        self.n, self.k, self.procs = n_plants, k, procs
        self.barrier = ctx.Barrier(procs + 1)
        cuts = np.linspace(0, n_plants, procs + 1).astype(int)
        self.ps = [ctx.Process(target=_cpu_proc, args=(int(cuts[i]), int(cuts[i + 1]), k, n_steps, self.barrier), daemon=True)
                   for i in range(procs)]
        for p in self.ps:
            p.start()

    def step(self) -> float:
        self.barrier.wait(timeout=600)           # workers have built their share and are ready
        t = time.perf_counter()
        self.barrier.wait(timeout=3600)
        return time.perf_counter() - t

    def close(self):
        for p in self.ps:
            p.join(timeout=30)


def _host_cores() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def cpu_baseline(target_seconds: float = 12.0):
    """Bounded sample of the cfg3 workload on ALL host cores."""
    procs = _host_cores()
    k = 8
    cal = CpuArm(64 * procs, k, procs, 1)
    rate = cal.n * k / cal.step()
    cal.close()
    n = int(max(64 * procs, min(PLANTS_PER_GPU, rate * target_seconds / k)))
    arm = CpuArm(n, k, procs, 1)
    el = arm.step()
    arm.close()
    return {"value": n * k / el, "unit": UNIT, "cores": procs, "kind": "port",
            "sample": f"{n} plants x {k} steps of the cfg3 workload ({el:.1f} s) on {procs} processes (all host cores), "
                      f"oracle/cpu_port.cpp (C restatement of the reference step; the Python reference itself, "
                      f"~140 plant-steps/s/core, cannot travel to this box)"}


def run_reference(args):
    """--impl reference: the CPU arm on the SAME config as the GPU arm — 65,536 plants x SUBSTEPS substeps per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = _host_cores()
    n, k = args.plants_per_gpu, args.substeps
    # one bench step of this arm is the full 65,536 x 128 plant-steps (~3 s on 16 cores); its `steps` are bounded so the
    # whole run stays within a few minutes
    W, K = min(args.warmup, 1), max(1, min(args.steps, 6))
    arm = CpuArm(n, k, threads, W + K)
    for _ in range(W):
        arm.step()
    el = [arm.step() for _ in range(K)]
    arm.close()
    v = n * k * K / float(np.sum(el))
    sample = (f"{n} plants x {k} substeps per step, {K} timed steps after {W} warm-up on {threads} processes (all host cores), "
              f"oracle/cpu_port.cpp (C restatement of the reference step)")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
            "warmup": W, "ms_per_step": 1e3 * float(np.mean(el)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "plants_per_gpu": n, "substeps_per_step": k},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def _bind_to_gpu_numa_node(torch, local: int):
    """Run this rank on the CPUs of its GPU's NUMA node, BEFORE the pinned staging buffers are allocated (first touch
    places them on that node).  No-op when the platform does not expose the topology (numa_node = -1)."""
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus |= set(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


class _StdoutToStderr:
    """NCCL prints its version banner on stdout when NCCL_DEBUG asks for it; rank 0's stdout carries exactly one JSON
    line, so file descriptor 1 points at stderr while the process group and its communicator come up."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--plants-per-gpu", type=int, default=PLANTS_PER_GPU)
    ap.add_argument("--substeps", type=int, default=SUBSTEPS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="value + full_step only (tuning runs)")
    ap.add_argument("--no-small", action="store_true", help="skip the small-batch / strong-scaling arm")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from nuclear_sim_b200 import N_STATE, load_snapshot, _clib
    from nuclear_sim_b200 import scenarios as sc
    from nuclear_sim_b200.sharded import ShardedBatchedSimulator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = _bind_to_gpu_numa_node(torch, local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        with _StdoutToStderr():
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
    W = max(3, args.warmup)
    K = args.steps
    n = args.plants_per_gpu
    ksub = args.substeps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    s0, params = load_snapshot("pwr3000_reactor_dt1")
    shard = ShardedBatchedSimulator(world * n, s0, params, rank=rank, world=world, device=str(dev))
    sim, pid = shard.sim, shard.plant_ids
    state_bytes = sim.slab.numel() * 8

    # per-launch inputs: INPUT_SETS distinct launches' worth, cycled; resident in HBM and mirrored in pinned host memory
    total = min(W + K, INPUT_SETS)
    acts_h = torch.empty((total, ksub, n), dtype=torch.int8).pin_memory()
    mags_h = torch.empty((total, ksub, n), dtype=torch.float64).pin_memory()
    noise_h = torch.empty((total, ksub, 5, n), dtype=torch.float64).pin_memory()

    def fill(i):
        a, m = sc.load_following_inputs(pid, i * ksub, ksub)
        acts_h[i] = torch.from_numpy(a); mags_h[i] = torch.from_numpy(m)
        for j in range(0, ksub, 16):
            noise_h[i, j:j + 16] = torch.from_numpy(sc.noise_inputs(pid, i * ksub + j, min(16, ksub - j)))
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(min(total, 4)) as ex:
        list(ex.map(fill, range(total)))
    acts_d, mags_d, noise_d = acts_h.to(dev), mags_h.to(dev), noise_h.to(dev)

    def timed_launches(s, count, after=None, kk=ksub, sets=(acts_d, mags_d, noise_d), warm=W):
        """`count` timed launches of sim `s` after `warm` warm-up launches; returns (total ms max over ranks, per-launch ms)."""
        A, M, Z = sets
        for i in range(warm):
            s.step(actions=A[i % total][:kk], magnitudes=M[i % total][:kk], noise=Z[i % total][:kk], K=kk)
            if after:
                after(s)
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(count + 1)]
        ev[0].record()
        for i in range(count):
            j = (warm + i) % total
            s.step(actions=A[j][:kk], magnitudes=M[j][:kk], noise=Z[j][:kk], K=kk)
            if after:
                after(s)
            ev[i + 1].record()
        barrier()
        return max_over_ranks(ev[0].elapsed_time(ev[count])), np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(count)])

    # ------------------------------------------------------------------ device-resident arm (value)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = sim.n_launches
    t_ms, per_launch_ms = timed_launches(sim, K)
    clocks = sampler.stop() if rank == 0 else None
    launches = sim.n_launches - launches0 - W
    value = world * n * ksub * K / (t_ms * 1e-3)

    # ------------------------------------------------------------------ full step: monitoring every substep + ring row per launch
    from nuclear_sim_b200 import maintenance as M
    cfg = json.load(open(os.path.join(ROOT, "nuclear-sim_b200", "data", "maintenance_system_template.json")))
    sim.reset()
    sim.set_thresholds(M.ThresholdTable(cfg).device_rows())
    sim.enable_monitor(per_substep=True, max_k=ksub)
    n_logged = sim.set_logged_columns(None, ring_rows=2)
    n_events = [0]

    def after_full(s):
        s.log_row()
        n_events[0] += int(s._mon["n_events"].item())      # the host reads the event count of every launch
        s._mon["n_events"].zero_()
    Kf = max(3, K // 2)
    tf_ms, _ = timed_launches(sim, Kf, after=after_full, warm=3)
    full_value = world * n * ksub * Kf / (tf_ms * 1e-3)
    log_fraction = 1.0 / ksub
    sim.disable_monitor()
    sim.clear_thresholds()

    e2e_value = e2e_rng_value = None
    h2d = ksub * n * (1 + 8 + 5 * 8)
    d2h = n * (22 * 8 + 8 + 1)
    loss_check = None
    extra = {}
    if not args.quick:
        # -------------------------------------------------------------- end-to-end arm (host buffers through the C ABI)
        D = sim.pipe_depth
        obs_h = [torch.empty((22, n), dtype=torch.float64).pin_memory() for _ in range(D)]
        rew_h = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(D)]
        done_h = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(D)]

        def e2e_loop(noise_sets):
            sim.reset()
            warm = []
            for i in range(W):       # warm-up through the SAME entry point: its staging sets are allocated on first use
                if i >= D:
                    sim.wait(warm[i - D])
                warm.append(sim.step_host_async(acts_h[i % total], mags_h[i % total], None if noise_sets is None else noise_sets[i % total],
                                                None, ksub, obs_h[i % D], rew_h[i % D], done_h[i % D]))
            for t in warm[-D:]:
                sim.wait(t)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reward_sum, tickets = 0.0, []
            e0.record()
            for i in range(K):
                b = i % D
                if i >= D:
                    sim.wait(tickets[i - D])
                    reward_sum += float(rew_h[b].mean())
                j = (W + i) % total
                tickets.append(sim.step_host_async(acts_h[j], mags_h[j], None if noise_sets is None else noise_sets[j], None, ksub,
                                                   obs_h[b], rew_h[b], done_h[b]))
            for i in range(max(0, K - D), K):
                sim.wait(tickets[i])
                reward_sum += float(rew_h[i % D].mean())
            e1.record()
            barrier()
            return world * n * ksub * K / (max_over_ranks(e0.elapsed_time(e1)) * 1e-3), reward_sum / K
        e2e_value, loss_check = e2e_loop(noise_h)
        shard.set_device_rng(20260118)
        e2e_rng_value, _ = e2e_loop(None)
        shard.set_device_rng(None)

        # -------------------------------------------------------------- control: config #2 plants at power
        s2, p2 = load_snapshot("pwr3000_steady_dt1")
        ctl = ShardedBatchedSimulator(world * n, s2, p2, rank=rank, world=world, device=str(dev)).sim
        none_acts = torch.full_like(acts_d, 8)
        Kc = max(3, K // 4)
        tc_ms, _ = timed_launches(ctl, Kc, sets=(none_acts, mags_d, noise_d), warm=3)
        extra["value_at_power"] = {"value": world * n * ksub * Kc / (tc_ms * 1e-3), "unit": UNIT,
                                   "workload": "cfg2 control: constant heat source, steady 100 %, NO_ACTION, same plant count",
                                   "mean_power_percent_after_run": float(ctl.state.power_level.mean())}
        del ctl, none_acts
        torch.cuda.empty_cache()

    # -------------------------------------------------------------- strong scaling / small batches
    def sub_batch(n_sub, count, shape=0):
        sub = ShardedBatchedSimulator(world * n_sub, s0, params, rank=rank, world=world, device=str(dev)).sim
        sub.set_small_batch_shape(shape)
        sets = (acts_d[:, :, :n_sub].contiguous(), mags_d[:, :, :n_sub].contiguous(), noise_d[:, :, :, :n_sub].contiguous())
        ms, _ = timed_launches(sub, count, sets=sets, warm=3)
        return world * n_sub * ksub * count / (ms * 1e-3)
    if args.no_small:
        pass
    elif world > 1:
        extra["strong_65536"] = {"value": sub_batch(PLANTS_PER_GPU // world, max(3, K // 2)), "unit": UNIT,
                                 "one_thread_per_plant": sub_batch(PLANTS_PER_GPU // world, max(3, K // 2), 1) if PLANTS_PER_GPU // world <= 18944 else None,
                                 "plants_per_gpu": PLANTS_PER_GPU // world,
                                 "note": "the 65,536-plant batch of the north star split over the GPUs (strong scaling)"}
    else:
        extra["small_batch"] = {"unit": UNIT, "plants_4096": sub_batch(4096, 3), "plants_8192": sub_batch(8192, 3),
                                "plants_16384": sub_batch(16384, 3), "plants_32768": sub_batch(32768, 3),
                                "one_thread_per_plant": {"plants_4096": sub_batch(4096, 3, 1), "plants_8192": sub_batch(8192, 3, 1),
                                                         "plants_16384": sub_batch(16384, 3, 1), "plants_32768": sub_batch(32768, 3, 1)},
                                "note": "config #2 / strong-scaled config #3 share per GPU / config #4 sizes on one GPU; default "
                                        "launch shape up to 18,944 plants = two threads per plant (source / sink halves pipelined "
                                        "by one substep), bit-identical to one thread per plant"}


    # ------------------------------------------------------------------ config #5: the loop with automatic maintenance
    # 131,072 plants per GPU, 24 h at dt = 5 min, every pump crossing its oil thresholds: monitored fused launches cut at
    # the 15-minute gate, native work-order table, maintenance kernel (profiles/run_cfg5_maintenance.py is the same loop)
    if not args.no_small and not args.quick:
        import importlib.util
        spec = importlib.util.spec_from_file_location("run_cfg5_maintenance", os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                                                      "profiles", "run_cfg5_maintenance.py"))
        cfg5 = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(cfg5)
        rec = cfg5.run_loop(131072, 24.0, "native", rank, world, local, parts=2, threads=True)
        if rank == 0:
            extra["cfg5_maintenance_loop"] = {
                "value": rec["plant_steps_per_s_whole_loop"], "unit": UNIT, "plants_per_gpu": rec["plants_per_gpu"],
                "steps": rec["steps"], "parts_per_gpu": rec["parts_per_gpu"], "work_orders_executed": rec["work_orders_executed"],
                "seconds_total": rec["seconds_total_max_over_ranks"],
                "note": "BASELINE config #5 end to end: thresholds every step inside the launches, work orders created / executed "
                        "by the native work-order table and the maintenance kernel; each GPU's plants as two independent batches "
                        "whose host work and launches overlap; wall clock, max over ranks"}

    # ------------------------------------------------------------------ trajectory summaries: the only collective
    summary = shard.gather_summaries(["pri.power_level", "sec.electrical_power_output", "pri.fuel_temperature", "pri.scram_status"])
    mean_power = float(summary[:, 0].mean())
    assert summary.shape[0] == world * n

    if rank == 0:
        hbm_peak, peak_src = _peaks()
        fp64 = ctypes.c_double(0.0); fp64_ms = ctypes.c_double(0.0)
        _clib.check(sim.L.nps_measure_fp64_peak(local, 4096, ctypes.byref(fp64), ctypes.byref(fp64_ms)))
        fp64_peak = float(fp64.value)
        avg_launch_s = float(per_launch_ms.mean()) * 1e-3
        rate_gpu = n * ksub / avg_launch_s                                  # dominant kernel, per GPU, measured live
        fp64_tflops = rate_gpu * FLOP_PER_PLANT_STEP / 1e12
        hbm_gbs = rate_gpu * (BYTES_PER_PLANT_STEP_K1 / ksub) / 1e9
        frac_fp64, frac_hbm = fp64_tflops / fp64_peak, hbm_gbs / hbm_peak
        rate_full = full_value / world
        frac_full = max(rate_full * FLOP_PER_PLANT_STEP / 1e12 / fp64_peak,
                        rate_full * (BYTES_PER_PLANT_STEP_K1 / ksub + BYTES_PER_LOGGED_ROW * log_fraction) / 1e9 / hbm_peak)
        n_live = _n_live_fields()
        live_bytes = 2 * n_live * 8 + 49 + ((N_STATE - n_live) * 8 + 185) / ksub
        cap = _capture()
        traffic = issue = None
        if cap:
            per = cap["plants"] * cap["substeps"]
            traffic = (cap["dram_bytes_read"] + cap["dram_bytes_write"]) / per * n * ksub
            issue = cap["warp_instructions_executed"] / per * n * ksub / (avg_launch_s * 148 * 4 * (clocks["sm_mhz"] or 1965.0) * 1e6)
        bound = "fp64" if frac_fp64 >= frac_hbm else "hbm"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": t_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "plants_per_gpu": n, "substeps_per_step": ksub, "n_state_fields": N_STATE,
                       "state_bytes_per_gpu": state_bytes,
                       "l2": "inputs larger than L2 (state slab %.0f MB per GPU is streamed every launch)" % (state_bytes / 1e6),
                       "mean_power_percent_after_run": mean_power, "rank0_numa_node": numa_node,
                       "distinct_input_sets": total},
            "full_step": {"value": full_value, "unit": UNIT, "steps": Kf, "thresholds_rows": 331, "threshold_events": n_events[0],
                          "logged_fields_per_row": n_logged, "log_fraction": log_fraction, "roofline_frac": frac_full,
                          "note": "thresholds + trip/scram stamps + reward/done after EVERY substep inside the launch, one "
                                  "full ring row (TMA copy kernel) and an event-count read-back per launch"},
            "e2e": None if e2e_value is None else {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                                                   "d2h_bytes_per_step": d2h, "mean_reward_readback": loss_check},
            "e2e_device_rng": None if e2e_rng_value is None else {
                "value": e2e_rng_value, "unit": UNIT, "h2d_bytes_per_step": ksub * n * (1 + 8), "d2h_bytes_per_step": d2h,
                "note": "same pipelined host-buffer loop, the five random draws per plant-step generated on the device"},
            "gpu_launches": launches,
            "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                       "samples": clocks["samples"]},
            "roofline": {"bound": bound, "kernel": "nps_step_kernel", "avg_launch_ms": avg_launch_s * 1e3,
                         "achieved": fp64_tflops if bound == "fp64" else hbm_gbs,
                         "peak": fp64_peak if bound == "fp64" else hbm_peak,
                         "unit": "TFLOP/s" if bound == "fp64" else "GB/s",
                         "frac": max(frac_fp64, frac_hbm), "traffic": traffic,
                         "definition": "SURVEY.md 8(d): max(rate x 2e4 flop / FP64 peak, rate x 26,240 B / K / HBM peak), per GPU",
                         "fp64": {"achieved_tflops": fp64_tflops, "peak_tflops": fp64_peak, "frac": frac_fp64,
                                  "peak_source": f"measured live: nps_measure_fp64_peak, DFMA chains, {fp64_ms.value:.2f} ms "
                                                 "(profiles/r02_fp64_peak.json)"},
                         "hbm": {"achieved_gbs": hbm_gbs, "peak_gbs": hbm_peak, "frac": frac_hbm, "peak_source": peak_src,
                                 "bytes_per_plant_step": BYTES_PER_PLANT_STEP_K1 / ksub},
                         "extras": {"note": "implementation-side figures, NOT the roofline fraction: the per-thread state "
                                            "frame streams through L1/L2/HBM every substep (DESIGN.md 5)",
                                    "live_set_bytes_per_plant_step": live_bytes,
                                    "live_set_gbs": rate_gpu * live_bytes / 1e9,
                                    "live_set_frac_of_hbm_peak": rate_gpu * live_bytes / 1e9 / hbm_peak,
                                    "dram_gbs_actual": (traffic / avg_launch_s / 1e9) if traffic else None,
                                    "issue_slot_frac": issue,
                                    "capture": os.path.basename(TRAFFIC_JSON) if cap else None}},
        }
        line.update(extra)
        if not args.no_cpu_baseline and world == 1 and not args.quick:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
