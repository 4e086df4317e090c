#!/bin/bash
# round 2, GPU call 8: cost of each part of the in-launch monitor + ncu capture of a monitored launch; smoke with the fixture check
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke8.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke8.log; tail -3 gpurun_out/smoke8.log
timeout 600 python profiles/monitor_cost.py > gpurun_out/monitor_cost.json 2> gpurun_out/monitor_cost.err; echo "rc=$?"; cat gpurun_out/monitor_cost.json; tail -3 gpurun_out/monitor_cost.err
timeout 600 python -m pytest tests -m gpu -q -k "split or export or error_conventions" > gpurun_out/pytest_gpu8.log 2>&1; tail -3 gpurun_out/pytest_gpu8.log
CMD="python bench.py --quick --no-small --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain8.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:nps_step_kernel -s 7 -c 1 -o gpurun_out/prof_step8_monitored $CMD > gpurun_out/ncu_full8.log 2>&1
cp nuclear-sim_b200/_lib/libnps_b200.so gpurun_out/libnps_b200_call8.so
