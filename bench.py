#!/usr/bin/env python
"""bench.py — plant-steps/sec of the batched plant-dynamics hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE config #3 — 65,536 plants PER GPU with randomised initial
conditions, ReactorHeatSource at equilibrium, load-following / power-ramp rod actions plus the
(inert) feedwater actions, dt = 1.0.  One bench "step" = ONE launch of the fused step kernel
advancing every plant by SUBSTEPS (=32) timesteps, i.e. plants x 32 plant-steps.  Plants shard over
ranks with no data-path collective (weak scaling); an NCCL all-gather of trajectory summaries runs
after the timed region.

value  : plant-steps/s with the per-step inputs already resident in HBM (CUDA events, max over ranks)
e2e    : the same metric through the host-buffer C-ABI call (nps_step_host_async, two launches in flight): pinned
         host inputs are copied in and observation/reward/done copied out and read on the host inside the timed
         region, every step
roofline / cpu_baseline: see DESIGN.md §Measurement.
"""
from __future__ import annotations

import argparse
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
INPUT_SETS = 12
SUBSTEPS = 128                       # fused substeps per launch (profiles/r01_tuning_variants.txt (20): 32 / 64 / 128)
FLOP_PER_PLANT_STEP = 2.0e4          # SURVEY.md §8d canonical figure (FP64 flop-equivalents)
METRIC = "plant-steps/sec"
UNIT = "plant-steps/s"
WORKLOAD = ("cfg3: 65,536 plants per GPU, randomized ICs, reactor heat source, load-following/power-ramp "
            f"rod + feedwater actions, dt=1.0, {SUBSTEPS} fused substeps per launch")


def _n_live_fields():
    """Fields that are read before written within a step (csrc/plant/live_fields.txt): the per-substep working set."""
    path = os.path.join(ROOT, "nuclear-sim_b200", "csrc", "plant", "live_fields.txt")
    return sum(1 for ln in open(path) if ln.strip() and not ln.startswith("#"))


def _measured_issue_fraction(n, ksub, launch_s, sm_mhz):
    """Share of the SMs' issue slots (148 SMs x 4 warp-instructions per cycle) the step kernel uses: executed
    warp-instructions per plant-substep from the committed ncu capture x this run's plants and launch time."""
    path = os.path.join(ROOT, "profiles", "r01_step_kernel_traffic.json")
    try:
        t = json.load(open(path))
        per = t["warp_instructions_executed"] / (t["plants"] * t["substeps"])
        return per * n * ksub / (launch_s * 148 * 4 * sm_mhz * 1e6)
    except Exception:
        return None


def _measured_traffic(n, ksub):
    """DRAM bytes per launch of nps_step_kernel from the committed ncu capture (profiles/r01_step_kernel_traffic.json),
    scaled per plant-substep; None when the capture is for a different build."""
    path = os.path.join(ROOT, "profiles", "r01_step_kernel_traffic.json")
    try:
        t = json.load(open(path))
        per = (t["dram_bytes_read"] + t["dram_bytes_write"]) / (t["plants"] * t["substeps"])
        return per * n * ksub
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


def cpu_baseline(steps_cpu: int = 4, threads: int = 1, target_seconds: float = 12.0):
    """Host oracle (C port of the reference step, oracle/cpu_port.cpp) on a bounded sample of the same workload."""
    import ctypes
    from concurrent.futures import ThreadPoolExecutor
    from nuclear_sim_b200 import load_snapshot
    from nuclear_sim_b200 import scenarios as sc
    from tests import _util as U
    L = U.oracle_lib()
    s0, params = load_snapshot("pwr3000_reactor_dt1")
    params = np.ascontiguousarray(params)
    # calibrate on a small sample, then size the timed sample for ~target_seconds of work
    n_cal = 64 * threads
    pid = np.arange(n_cal)

    def run(pid, k):
        st = np.ascontiguousarray(sc.randomized_states(s0, pid))
        acts, mags = sc.load_following_inputs(pid, 0, k)
        noise = sc.noise_inputs(pid, 0, k)
        a = np.ascontiguousarray(acts.T); m = np.ascontiguousarray(mags.T)
        z = np.ascontiguousarray(noise.transpose(2, 0, 1))      # [n, k, 5]
        chunks = np.array_split(np.arange(len(pid)), threads)

        def work(c):
            if len(c) == 0:
                return
            lo, hi = c[0], c[-1] + 1
            L.nps_oracle_step(U.ptr(st[lo:hi]), U.ptr(params), U.ptr(a[lo:hi]), U.ptr(m[lo:hi]), U.ptr(z[lo:hi]),
                              ctypes.c_int64(hi - lo), int(k))
        t = time.perf_counter()
        if threads == 1:
            work(chunks[0])
        else:
            with ThreadPoolExecutor(threads) as ex:
                list(ex.map(work, chunks))
        return time.perf_counter() - t
    tcal = run(pid, steps_cpu)
    rate = n_cal * steps_cpu / tcal
    n = int(max(n_cal, min(PLANTS_PER_GPU, rate * target_seconds / steps_cpu)))
    n -= n % threads
    el = run(np.arange(n), steps_cpu)
    return {"value": n * steps_cpu / el, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} plants x {steps_cpu} steps of the cfg3 workload ({el:.1f} s), oracle/cpu_port.cpp (C restatement "
                      f"of the reference step; the Python reference itself cannot travel to this box)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_baseline(steps_cpu=SUBSTEPS, threads=threads, target_seconds=max(2.0, 60.0 / max(1, args.warmup + args.steps)))
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    n_plants = int(vals[-1]["sample"].split(" plants")[0])
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * n_plants * SUBSTEPS / v, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "plants_per_step_sample": n_plants, "substeps_per_step": SUBSTEPS},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": vals[-1]["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def _bind_to_gpu_numa_node(torch, local: int):
    """Run this rank on the CPUs of its GPU's NUMA node, BEFORE the pinned staging buffers are allocated (first touch
    places them on that node): with eight ranks on a two-socket host the e2e arm otherwise pulls half of its input
    stream across the socket interconnect.  No-op when the platform does not expose the topology (numa_node = -1)."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--plants-per-gpu", type=int, default=PLANTS_PER_GPU)
    ap.add_argument("--substeps", type=int, default=SUBSTEPS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg2"],
                    help="cfg3 (headline): reactor heat source + load-following actions; cfg2: constant heat source, "
                         "steady 100 %%, NO_ACTION (same plant count; diagnostic for branch-mix sensitivity)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from nuclear_sim_b200 import BatchedNuclearPlantSimulator, N_STATE, load_snapshot
    from nuclear_sim_b200 import scenarios as sc

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
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    K = args.steps
    n = args.plants_per_gpu
    ksub = args.substeps

    s0, params = load_snapshot("pwr3000_reactor_dt1" if args.workload == "cfg3" else "pwr3000_steady_dt1")
    pid = np.arange(rank * n, (rank + 1) * n)
    sim = BatchedNuclearPlantSimulator(n, sc.randomized_states(s0, pid), params, device=str(dev))
    state_bytes = sim.slab.numel() * 8

    # per-launch inputs (distinct for up to INPUT_SETS consecutive launches), resident in HBM and mirrored in pinned host memory
    # at most INPUT_SETS distinct launches' worth of inputs, cycled: keeps pinned host memory (and its device mirror)
    # bounded for any --steps / --warmup the caller chooses (one set is 411 MB at 128 substeps x 65,536 plants)
    total = min(W + K, INPUT_SETS)
    acts_h = torch.empty((total, ksub, n), dtype=torch.int8).pin_memory()
    mags_h = torch.empty((total, ksub, n), dtype=torch.float64).pin_memory()
    noise_h = torch.empty((total, ksub, 5, n), dtype=torch.float64).pin_memory()
    for i in range(total):
        a, m = sc.load_following_inputs(pid, i * ksub, ksub)
        if args.workload == "cfg2":
            a[:] = 8
        acts_h[i] = torch.from_numpy(a); mags_h[i] = torch.from_numpy(m)
        noise_h[i] = torch.from_numpy(sc.noise_inputs(pid, i * ksub, ksub))
    acts_d, mags_d, noise_d = acts_h.to(dev), mags_h.to(dev), noise_h.to(dev)
    D = sim.pipe_depth       # launches nps_step_host_async keeps in flight; one set of pinned result buffers per slot
    obs_h = [torch.empty((22, n), dtype=torch.float64).pin_memory() for _ in range(D)]
    rew_h = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(D)]
    done_h = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(D)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ device-resident arm (value)
    for i in range(W):
        sim.step(actions=acts_d[i % total], magnitudes=mags_d[i % total], noise=noise_d[i % total], K=ksub)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    launches0 = sim.n_launches
    barrier()
    ev[0].record()
    for i in range(K):
        sim.step(actions=acts_d[(W + i) % total], magnitudes=mags_d[(W + i) % total], noise=noise_d[(W + i) % total], K=ksub)
        ev[i + 1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = sim.n_launches - launches0
    t_total_ms = ev[0].elapsed_time(ev[K])
    per_launch_ms = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(K)])
    t = torch.tensor([t_total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_ms = float(t.item())
    value = world * n * ksub * K / (t_ms * 1e-3)

    # ------------------------------------------------------------------ end-to-end arm (host buffers through the C ABI)
    # Every step: pinned host inputs -> device, the fused substeps, observation/reward/done -> pinned host, and the host reads
    # the step's reward.  nps_step_host_async keeps up to D launches in flight: the input copies of later steps overlap
    # the kernel of step i, and the host consumes step i's result while the next steps run.
    sim.reset()
    warm = []
    for i in range(W):       # warm-up through the SAME entry point: its staging sets are allocated on first use
        if i >= D:
            sim.wait(warm[i - D])
        warm.append(sim.step_host_async(acts_h[i % total], mags_h[i % total], noise_h[i % total], None, ksub, obs_h[i % D], rew_h[i % D], done_h[i % D]))
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
        tickets.append(sim.step_host_async(acts_h[(W + i) % total], mags_h[(W + i) % total], noise_h[(W + i) % total], None, ksub, obs_h[b], rew_h[b], done_h[b]))
    for i in range(max(0, K - D), K):
        sim.wait(tickets[i])
        reward_sum += float(rew_h[i % D].mean())
    e1.record()
    barrier()
    te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n * ksub * K / (float(te.item()) * 1e-3)
    h2d = ksub * n * (1 + 8 + 5 * 8)
    d2h = n * (22 * 8 + 8 + 1)
    loss_check = reward_sum / K

    # the same loop with device-side noise (nps_set_device_rng): only actions and magnitudes travel (9 B per plant-step)
    sim.reset()
    sim.set_device_rng(20260118, plant_offset=rank * n)
    tickets = []
    for i in range(min(W, D)):
        tickets.append(sim.step_host_async(acts_h[i % total], mags_h[i % total], None, None, ksub, obs_h[i % D], rew_h[i % D], done_h[i % D]))
    for t_ in tickets:
        sim.wait(t_)
    barrier()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tickets = []
    r0.record()
    for i in range(K):
        b = i % D
        if i >= D:
            sim.wait(tickets[i - D])
            reward_sum += float(rew_h[b].mean())
        tickets.append(sim.step_host_async(acts_h[(W + i) % total], mags_h[(W + i) % total], None, None, ksub, obs_h[b], rew_h[b], done_h[b]))
    for i in range(max(0, K - D), K):
        sim.wait(tickets[i])
    r1.record()
    barrier()
    tr = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tr, op=dist.ReduceOp.MAX)
    e2e_rng_value = world * n * ksub * K / (float(tr.item()) * 1e-3)
    sim.set_device_rng(None)

    # ------------------------------------------------------------------ trajectory summaries: the only collective
    summary = torch.stack([sim.state.power_level, sim.state.electrical_power_output, sim.state.fuel_temperature,
                           sim.state.scram_status]).t().contiguous()
    if world > 1:
        gathered = [torch.empty_like(summary) for _ in range(world)]
        dist.all_gather(gathered, summary)
        summary = torch.cat(gathered)
    mean_power = float(summary[:, 0].mean())

    if rank == 0:
        peak, peak_src = _peaks()
        # Algorithmic bytes of one launch: a plant's state is 10.4 KB and 448 plants are resident per
        # SM (4.7 MB against 0.5 MB of registers + shared memory), so EVERY substep must stream the fields that are
        # live on entry in from HBM and the same number back out; fields that are only outputs leave once per launch.
        n_live = _n_live_fields()
        per_substep = 2 * n_live * 8 + (1 + 8 + 40)                        # live state in + out, per-substep inputs
        per_launch = (N_STATE - n_live) * 8 + (22 * 8 + 8 + 1)              # output-only fields, obs/reward/done
        bytes_per_launch = n * (ksub * per_substep + per_launch)
        resident_bytes_per_launch = 2 * state_bytes + ksub * n * (1 + 8 + 40) + n * (22 * 8 + 8 + 1)
        avg_launch_s = float(per_launch_ms.mean()) * 1e-3
        achieved = bytes_per_launch / avg_launch_s / 1e9
        traffic = _measured_traffic(n, ksub)
        fp64_tflops = n * ksub * FLOP_PER_PLANT_STEP / avg_launch_s / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": t_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD if args.workload == "cfg3" else WORKLOAD.replace("cfg3", "cfg2 (steady 100 %, constant heat source, NO_ACTION)"),
                       "plants_per_gpu": n, "substeps_per_step": ksub, "n_state_fields": N_STATE,
                       "state_bytes_per_gpu": state_bytes,
                       "l2": "inputs larger than L2 (state slab %.0f MB per GPU is streamed every launch)" % (state_bytes / 1e6),
                       "mean_power_percent_after_run": mean_power, "rank0_numa_node": numa_node,
                       "distinct_input_sets": total},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "mean_reward_readback": loss_check},
            "e2e_device_rng": {"value": e2e_rng_value, "unit": UNIT, "h2d_bytes_per_step": ksub * n * (1 + 8),
                               "d2h_bytes_per_step": d2h,
                               "note": "same pipelined host-buffer loop, the five random draws per plant-step generated "
                                       "on the device (Philox4x32-10) instead of copied from the host"},
            "gpu_launches": launches,
            "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                       "samples": clocks["samples"]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "nps_step_kernel",
                         "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_launch_s * 1e3,
                         "algorithmic_bytes_per_plant_step": bytes_per_launch / (n * ksub), "n_live_fields": n_live,
                         "note": "per-substep streaming of the live state (it cannot be resident: 4.8 MB per SM); "
                                 "frac_if_state_resident is the K-amortised figure of SURVEY 8d; issue_slot_frac is the "
                                 "share of the issue roofline (148 SMs x 4 warp-instructions/cycle), the bound that "
                                 "actually binds this scalar FP64 path (DESIGN.md 5)",
                         "frac_if_state_resident": resident_bytes_per_launch / avg_launch_s / 1e9 / peak,
                         "dram_gbs_actual": (traffic / avg_launch_s / 1e9) if traffic else None,
                         "issue_slot_frac": _measured_issue_fraction(n, ksub, avg_launch_s, clocks["sm_mhz"] or 1965.0),
                         "fp64_tflops_at_2e4_flop_per_plant_step": fp64_tflops,
                         "fp64_frac_of_37_tflops": fp64_tflops / 37.0},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
