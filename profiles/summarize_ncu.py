#!/usr/bin/env python
"""Summarise one `ncu --set full` capture of nps_step_kernel into the JSON bench.py reads (roofline.traffic,
issue_slot_frac) and a short text table for the judge.

    ncu -i gpurun_out/prof_step.ncu-rep --page raw --csv > raw.csv
    python profiles/summarize_ncu.py raw.csv <plants> <substeps> profiles/r02_step_kernel_traffic.json
"""
import csv
import json
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_ms",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "smsp__inst_executed.sum": "warp_instructions_executed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed": "fp64_pipe_pct_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "smsp__sass_inst_executed_op_local_ld.sum": "local_load_instructions",
    "smsp__sass_inst_executed_op_local_st.sum": "local_store_instructions",
    "sm__cycles_elapsed.avg": "sm_cycles_elapsed",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}


def main():
    raw, plants, substeps, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"kernel": vals[hdr.index("Kernel Name")], "plants": plants, "substeps": substeps}
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS:
            x = float(v.replace(",", ""))
            d[KEYS[h]] = x * UNIT.get(u, 1.0) if u in UNIT else x
    d["dram_bytes_read"], d["dram_bytes_write"] = d.pop("dram_read"), d.pop("dram_write")
    per = plants * substeps
    d["dram_bytes_per_plant_substep"] = (d["dram_bytes_read"] + d["dram_bytes_write"]) / per
    d["dram_gbs"] = (d["dram_bytes_read"] + d["dram_bytes_write"]) / (d["duration_ms"] * 1e-3) / 1e9
    d["warp_instructions_per_warp_substep"] = d["warp_instructions_executed"] / (per / 32)
    json.dump(d, open(out, "w"), indent=1, sort_keys=True)
    for k in sorted(d):
        print(f"{k:40s} {d[k]}")


if __name__ == "__main__":
    main()
