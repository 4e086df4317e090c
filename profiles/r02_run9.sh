#!/bin/bash
# round 2, GPU call 9: monitor cost after grouping its loads; event / split / maintenance tests
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python profiles/monitor_cost.py > gpurun_out/monitor_cost2.json 2> gpurun_out/monitor_cost2.err; echo "rc=$?"; cat gpurun_out/monitor_cost2.json; tail -3 gpurun_out/monitor_cost2.err
timeout 900 python -m pytest tests -m gpu -q -k "fused or split or maintenance or status or threshold or embedded" > gpurun_out/pytest_gpu9.log 2>&1; tail -4 gpurun_out/pytest_gpu9.log
