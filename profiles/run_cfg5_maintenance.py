#!/usr/bin/env python
"""BASELINE config #5: long-horizon maintenance degradation with the FULL loop, sharded over the GPUs of one node.

    python profiles/run_cfg5_maintenance.py [--plants-per-gpu 131072] [--hours 24]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 profiles/run_cfg5_maintenance.py

dt = 5 min; every rank owns a contiguous plant range (ShardedBatchedSimulator) and runs its own work-order
bookkeeping (ColumnarAutoMaintenance.advance): thresholds are evaluated after every substep inside the launch, a
launch ends on each 15-minute gate step, due work orders are applied on the device, the gate step is checked by the
event-list flag kernel.  No collective on the path; per-action counts are all-gathered at the end.
Prints one JSON line: whole-job plant-steps/s, per-rank wall-clock split (step kernel vs bookkeeping), work orders."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nuclear_sim_b200 import load_snapshot, field_index  # noqa: E402
from nuclear_sim_b200 import scenarios as sc  # noqa: E402
from nuclear_sim_b200 import maintenance as M  # noqa: E402
from nuclear_sim_b200.maintenance import ThresholdTable  # noqa: E402
from nuclear_sim_b200.sharded import ShardedBatchedSimulator  # noqa: E402


def run_loop(plants_per_gpu=131072, hours=24.0, bookkeeping="native", rank=0, world=1, local=0, parts=1, threads=True):
    """The loop itself (torch.distributed already initialised when world > 1); returns the result record on rank 0,
    None elsewhere.  bench.py calls this for the `cfg5_maintenance_loop` key of its line."""
    args = argparse.Namespace(plants_per_gpu=plants_per_gpu, hours=hours, bookkeeping=bookkeeping)
    n, dt = args.plants_per_gpu, 5.0
    s0, params = load_snapshot("pwr3000_oil_top_off_dt5")
    ix = field_index()

    def states(pid):
        """Initial conditions positioned near maintenance thresholds: oil levels just above 58 %,
        contamination near 15.2 ppm — a pure function of the global plant id."""
        st = sc.randomized_states(s0, pid)
        u = sc.noise_inputs(pid, 0, 2, seed=7)[:, 2:, :]          # [2, 3, n] uniforms keyed by plant id
        for p in range(4):
            st[:, ix[f"fw.pump[{p}].lub.oil_level"]] = 58.0 + 6.0 * u[p // 3, p % 3]
            st[:, ix[f"fw.pump[{p}].lub.oil_contamination_level"]] = 15.2 - 0.6 * u[(p + 1) // 3 % 2, (p + 1) % 3]
        return st
    cfg = json.load(open(os.path.join(ROOT, "nuclear-sim_b200", "data", "maintenance_system_template.json")))
    cls = M.NativeAutoMaintenance if args.bookkeeping == "native" else M.ColumnarAutoMaintenance
    if parts > 1:
        return _run_interleaved(n, dt, args, parts, rank, world, local, s0, params, states, cfg, cls, threads)
    shard = ShardedBatchedSimulator(world * n, s0, params, rank=rank, world=world, device=f"cuda:{local}", states=states)
    maint = cls(shard.sim, ThresholdTable(cfg), aggressive=True)
    shard.sim.enable_monitor(event_capacity=8 * n)
    steps = int(args.hours * 60 / dt)
    maint.advance(3)                                    # warm-up: first launches, allocations
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    timers = {}
    maint.seconds_host = maint.seconds_device_calls = 0.0
    t0 = time.perf_counter()
    maint.advance(steps, timers=timers)
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    dev_calls = maint.seconds_device_calls + timers.get("event_drain", 0.0)
    t = torch.tensor([total, timers["step_kernel"], timers["bookkeeping"], maint.seconds_host, dev_calls], dtype=torch.float64,
                     device=f"cuda:{local}")
    counts = torch.tensor([maint.n_work_orders_created, maint.n_work_orders_executed, sum(len(c["plant"]) for c in maint.event_cols)],
                          dtype=torch.int64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    summary = shard.gather_summaries(["fw.pump[0].lub.oil_level", "pri.power_level"])
    if rank == 0:
        total, t_dev, t_host, t_numpy, t_devcalls = (float(x) for x in t)
        return {
            "workload": "cfg5: long-horizon maintenance degradation, dt=5 min, launches cut at the 15-min gate, full loop",
            "n_gpus": world, "plants": world * n, "plants_per_gpu": n, "simulated_hours": args.hours, "steps": steps,
            "launches_per_rank": timers["launches"], "plant_steps": world * n * steps,
            "plant_steps_per_s_whole_loop": world * n * steps / total, "seconds_total_max_over_ranks": total,
            "seconds_step_kernel_max_over_ranks": t_dev, "seconds_bookkeeping_max_over_ranks": t_host,
            "bookkeeping_over_kernel": t_host / t_dev,
            "seconds_host_numpy_max_over_ranks": t_numpy, "host_numpy_over_kernel": t_numpy / t_dev,
            "seconds_device_side_of_bookkeeping_max_over_ranks": t_devcalls,
            "split": "bookkeeping = host numpy (decisions, dedupe, work-order columns) + device-side calls (event drains, "
                     "maintenance kernel with its request / status copies, gate-step flag kernel)",
            "threshold_events": int(counts[2]), "work_orders_created": int(counts[0]), "work_orders_executed": int(counts[1]),
            "by_action_rank0": maint.counts_by_action(), "mean_oil_level_pump0": float(summary[:, 0].mean()),
            "bookkeeping": type(maint).__name__ + " (in-launch threshold events + event-list flag kernel at gate steps; native = nps_wo_* work-order table in the library, columnar = numpy columns)"}
    return None


def _run_interleaved(n, dt, args, parts, rank, world, local, s0, params, states, cfg, cls, threads=True):
    """The same loop with this rank's plants cut into `parts` independent batches (contiguous global id ranges, own
    simulator, stream and books each) resumed in turn: one part's host work overlaps another part's launch."""
    shards = [ShardedBatchedSimulator(world * n, s0, params, rank=rank * parts + j, world=world * parts, device=f"cuda:{local}",
                                      states=states) for j in range(parts)]
    maints = [cls(sh.sim, ThresholdTable(cfg), aggressive=True) for sh in shards]
    for sh in shards:
        sh.sim.enable_monitor(event_capacity=8 * sh.n_plants)
    steps = int(args.hours * 60 / dt)
    run = M.advance_threaded if threads else M.advance_interleaved
    run(maints, 3)                                      # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    for m in maints:
        m.seconds_host = m.seconds_device_calls = 0.0
    t0 = time.perf_counter()
    run(maints, steps)
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    t = torch.tensor([total, sum(m.seconds_host for m in maints)], dtype=torch.float64, device=f"cuda:{local}")
    counts = torch.tensor([sum(m.n_work_orders_created for m in maints), sum(m.n_work_orders_executed for m in maints),
                           sum(len(c["plant"]) for m in maints for c in m.event_cols)], dtype=torch.int64, device=f"cuda:{local}")
    oil = torch.cat([sh.sim.state["fw.pump[0].lub.oil_level"] for sh in shards]).sum().reshape(1)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        dist.all_reduce(oil, op=dist.ReduceOp.SUM)
    if rank == 0:
        by = {}
        for m in maints:
            for k, v in m.counts_by_action().items():
                by[k] = by.get(k, 0) + v
        return {"workload": "cfg5: long-horizon maintenance degradation, dt=5 min, launches cut at the 15-min gate, full loop",
                "n_gpus": world, "plants": world * n, "plants_per_gpu": n, "parts_per_gpu": parts, "simulated_hours": args.hours,
                "steps": steps, "plant_steps": world * n * steps, "plant_steps_per_s_whole_loop": world * n * steps / float(t[0]),
                "seconds_total_max_over_ranks": float(t[0]), "seconds_host_numpy_max_over_ranks": float(t[1]),
                "threshold_events": int(counts[2]), "work_orders_created": int(counts[0]), "work_orders_executed": int(counts[1]),
                "by_action_rank0": by, "mean_oil_level_pump0": float(oil[0]) / (world * n),
                "bookkeeping": cls.__name__ + f", {parts} independent plant batches per GPU, " +
                               ("one host thread and CUDA stream each (advance_threaded)" if threads else
                                "resumed in turn by one host thread (advance_interleaved)") +
                               ": one batch's host work runs while another batch's launch occupies the GPU"}
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--plants-per-gpu", type=int, default=131072)
    ap.add_argument("--hours", type=float, default=24.0)
    ap.add_argument("--parts", type=int, default=2, help="independent plant batches per GPU whose host work and launches overlap (1: one batch)")
    ap.add_argument("--one-thread", action="store_true", help="resume the batches in turn from one host thread instead of one thread per batch")
    ap.add_argument("--bookkeeping", choices=["native", "columnar"], default="native",
                    help="native: the library's work-order table (nps_wo_*); columnar: numpy columns")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)            # NCCL's version banner goes to stderr: stdout carries the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
        torch.cuda.synchronize()
        os.dup2(saved, 1)
        os.close(saved)
    rec = run_loop(args.plants_per_gpu, args.hours, args.bookkeeping, rank, world, local, args.parts, not args.one_thread)
    if rank == 0:
        print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
