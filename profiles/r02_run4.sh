#!/bin/bash
# round 2, GPU call 4: whole GPU suite (split kernel now serves every small batch), full bench, cfg5 loop, ncu of the current build
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke4.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke4.log; tail -2 gpurun_out/smoke4.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu4.log
tail -12 gpurun_out/pytest_gpu4.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench4_n1.json 2> gpurun_out/bench4_n1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench4_n1.err
timeout 600 python profiles/run_cfg5_maintenance.py > gpurun_out/cfg5_4_n1.json 2> gpurun_out/cfg5_4_n1.err; echo "cfg5 rc=$?"; cat gpurun_out/cfg5_4_n1.json
CMD="python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:nps_step_kernel -s 3 -c 1 -o gpurun_out/prof_step4 $CMD > gpurun_out/ncu_full4.log 2>&1
ls -la gpurun_out | tail -8
