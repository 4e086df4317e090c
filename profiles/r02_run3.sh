#!/bin/bash
# round 2, GPU call 3: new maintenance tests, cfg5 full loop on one GPU, quick bench of the current build
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "maintenance or fused or status or env_cooling" > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log
tail -8 gpurun_out/pytest_gpu3.log
timeout 600 python profiles/run_cfg5_maintenance.py > gpurun_out/cfg5_n1.json 2> gpurun_out/cfg5_n1.err; echo "cfg5 rc=$?"; tail -3 gpurun_out/cfg5_n1.err; cat gpurun_out/cfg5_n1.json
timeout 300 python bench.py --quick --steps 6 --warmup 3 > gpurun_out/quick_base3.json 2> gpurun_out/quick_base3.err; echo "bench rc=$?"
