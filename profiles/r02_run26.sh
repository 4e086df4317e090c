#!/bin/bash
# round 2, GPU call 26: fresh ncu --set full capture of nps_step_kernel<448,1> (per-line attribution of the final build)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
CMD="python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline --no-small"
$CMD > gpurun_out/plain26.log 2>&1; rc=$?; echo "plain rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:nps_step_kernel -s 3 -c 1 -o gpurun_out/prof_step26 $CMD > gpurun_out/ncu26.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu26.log
fi
