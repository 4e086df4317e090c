#!/bin/bash
# round 2, GPU call 7: final-build measurements on one GPU (suite, both bench arms, cfg5 split timers, ncu launch list + full capture)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu7.log
tail -6 gpurun_out/pytest_gpu7.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench7_ref.json 2> gpurun_out/bench7_ref.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/bench7_n1.json 2> gpurun_out/bench7_n1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench7_n1.err
timeout 600 python profiles/run_cfg5_maintenance.py > gpurun_out/cfg5_7_n1.json 2> gpurun_out/cfg5_7_n1.err; echo "cfg5 rc=$?"; cat gpurun_out/cfg5_7_n1.json
CMD="python bench.py --quick --no-small --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain7.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches7.csv $CMD > gpurun_out/ncu_launches7.log 2>&1
$CMD > gpurun_out/plain7b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:nps_step_kernel -s 3 -c 1 -o gpurun_out/prof_step7 $CMD > gpurun_out/ncu_full7.log 2>&1
ls -la gpurun_out | tail -5
