#!/bin/bash
# round 2, GPU call 22: full GPU suite (catalogue sweep, new maintenance targets, frame_skip) + smoke + quick bench
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke22.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke22.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu22.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu22.log
timeout 600 python bench.py --quick --steps 3 --warmup 3 > gpurun_out/bench_quick22.json 2> gpurun_out/bench_quick22.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_quick22.json
