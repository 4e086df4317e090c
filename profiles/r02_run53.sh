#!/bin/bash
# round 2, GPU call 53: launch list of a bench run with the final build (ncu --metrics gpu__time_duration.sum)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
CMD="python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline --no-small"
$CMD > gpurun_out/plain53.log 2>&1; rc=$?; echo "plain rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu53.log 2>&1; echo "ncu rc=$?"
fi
