#!/bin/bash
# round 2, GPU call 23: values per memory round trip in the in-launch monitor (8 / 16 / 32)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for v in mc8 mc16 mc32; do
  NPS_B200_LIB=$PWD/nuclear-sim_b200/_lib/libnps_b200_$v.so timeout 600 python profiles/monitor_cost.py > gpurun_out/monitor_cost_$v.json 2> gpurun_out/monitor_cost_$v.err; echo "$v rc=$?"; cat gpurun_out/monitor_cost_$v.json
done
