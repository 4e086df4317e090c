#!/bin/bash
# round 2, GPU call 12: final-build verification on one GPU: smoke, whole GPU suite, both bench arms, cfg5 loop, launch list + full ncu capture (+ the matching .so for attribution)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke12.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke12.log; tail -3 gpurun_out/smoke12.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu12.log; tail -5 gpurun_out/pytest_gpu12.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench12_ref.json 2> gpurun_out/bench12_ref.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/bench12_n1.json 2> gpurun_out/bench12_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench12_n1.err
timeout 600 python profiles/run_cfg5_maintenance.py > gpurun_out/cfg5_12_n1.json 2> gpurun_out/cfg5_12_n1.err; echo "cfg5 rc=$?"
CMD="python bench.py --quick --no-small --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain12.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches12.csv $CMD > gpurun_out/ncu_launches12.log 2>&1
$CMD > gpurun_out/plain12b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:nps_step_kernel -s 3 -c 1 -o gpurun_out/prof_step12 $CMD > gpurun_out/ncu_full12.log 2>&1
cp nuclear-sim_b200/_lib/libnps_b200.so gpurun_out/libnps_b200_call12.so
