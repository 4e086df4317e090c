#!/bin/bash
# round 2, GPU call 10: cache operator of the monitor's loads (plain / .cs / .lu), one box
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for v in base moncs monlu; do
  case $v in base) unset NPS_B200_LIB;; *) export NPS_B200_LIB=$PWD/nuclear-sim_b200/_lib/libnps_b200_$v.so;; esac
  timeout 300 python profiles/monitor_cost.py > gpurun_out/monitor_cost_$v.json 2> gpurun_out/monitor_cost_$v.err; echo "$v rc=$?"; cat gpurun_out/monitor_cost_$v.json
done
