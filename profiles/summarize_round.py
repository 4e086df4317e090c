#!/usr/bin/env python
"""Turn the scratch outputs of profiles/round_end.sh (gpurun_out/) into the tracked summaries under profiles/:
   r01_ncu_full_step_kernel_final.csv, r01_step_kernel_traffic.json, r01_launches_bench_final.csv,
   r01_attribution_step_kernel_final.txt, r01_bench_n1.json, r01_bench_reference_arm.json."""
import csv
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"


def main():
    rows = list(csv.reader(open(os.path.join(G, "prof_step_final_raw.csv"))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    g = lambda k: vals[hdr.index(k)]
    keys = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
            "launch__registers_per_thread", "launch__waves_per_multiprocessor", "l1tex__t_sector_hit_rate.pct",
            "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__cycles_elapsed.avg", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
    keys += [k for k in hdr if k.startswith("smsp__pcsamp_warps_issue_stalled") and not k.endswith("not_issued")]
    bench = json.loads(open(os.path.join(G, "bench_n1.log")).read().strip().splitlines()[-1])
    n, k = bench["config"]["plants_per_gpu"], bench["config"]["substeps_per_step"]
    with open(os.path.join(P, f"{TAG}_ncu_full_step_kernel_final.csv"), "w") as fh:
        fh.write(f"# ncu --set full, nps_step_kernel, {TAG} final build, {n} plants x {k} substeps per launch (one launch)\n")
        fh.write("# command: ncu --set full --clock-control none --import-source on -k regex:nps_step_kernel -s 3 -c 1 "
                 "python bench.py --steps 3 --warmup 3 --no-cpu-baseline\nmetric,unit,value\n")
        for key in keys:
            if key in hdr:
                fh.write(f"{key},{units[hdr.index(key)]},{g(key)}\n")
    r, w = float(g("dram__bytes_read.sum")) * 1e9, float(g("dram__bytes_write.sum")) * 1e9
    json.dump({"kernel": "nps_step_kernel<448,1>", "plants": n, "substeps": k, "dram_bytes_read": r, "dram_bytes_write": w,
               "duration_ms_under_ncu": float(g("gpu__time_duration.sum")),
               "warp_instructions_executed": float(g("smsp__inst_executed.sum")),
               "source": f"profiles/{TAG}_ncu_full_step_kernel_final.csv (ncu --set full, one launch)"},
              open(os.path.join(P, f"{TAG}_step_kernel_traffic.json"), "w"), indent=1)
    shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, f"{TAG}_launches_bench_final.csv"))
    shutil.copy(os.path.join(G, "bench_n1.log"), os.path.join(P, f"{TAG}_bench_n1.json"))
    shutil.copy(os.path.join(G, "bench_ref.log"), os.path.join(P, f"{TAG}_bench_reference_arm.json"))
    # attribution: disassemble the library that was profiled (the in-tree build)
    tmp = tempfile.mkdtemp()
    lib = os.path.join(ROOT, "nuclear-sim_b200", "_lib", "libnps_b200.so")
    subprocess.check_call(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    with open(os.path.join(tmp, "dis.txt"), "w") as fh:
        subprocess.call(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], stdout=fh, stderr=subprocess.DEVNULL)
    out = subprocess.run([sys.executable, os.path.join(P, "attribute_sass.py"), os.path.join(tmp, "dis.txt"),
                          os.path.join(G, "prof_step_final_sass.csv"), str(n // 32 * k), "nps_step_kernelILi448"],
                         capture_output=True, text=True).stdout
    with open(os.path.join(P, f"{TAG}_attribution_step_kernel_final.txt"), "w") as fh:
        fh.write(f"# nps_step_kernel<448,1>, {TAG} final build, {n} plants x {k} substeps ({n // 32 * k} warp-substeps)\n")
        fh.write("# profiles/attribute_sass.py: ncu source page of the full capture joined with nvdisasm -gi line info\n")
        fh.write("\n".join(out.splitlines()[:80]) + "\n")
    print("per plant-substep DRAM bytes:", (r + w) / (n * k), " value", bench["value"], " e2e", bench["e2e"]["value"])


if __name__ == "__main__":
    main()
