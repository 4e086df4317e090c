#!/bin/bash
# round 2, GPU call 25: full GPU suite with the new fixtures (cfg9, 16 trip cases, catalogue sweep) + smoke
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke25.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke25.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu25.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu25.log
