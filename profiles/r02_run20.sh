#!/bin/bash
# round 2, GPU call 20: ncu --set full of the two-threads-per-plant kernel at 4,096 and 16,384 plants
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python profiles/split_profile.py > gpurun_out/split_profile.log 2>&1; rc=$?; echo "plain rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:nps_step_split_kernel -c 4 -o gpurun_out/prof_split python profiles/split_profile.py > gpurun_out/ncu_split.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_split.log
fi
ls -la gpurun_out | tail -5
