#!/bin/bash
# round 2, GPU call 15: auxiliary kernels + maintenance / threshold tests after the event-list flag kernel rewrite
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python profiles/aux_kernels.py > gpurun_out/aux_kernels.json 2> gpurun_out/aux_kernels.err; echo "rc=$?"; cat gpurun_out/aux_kernels.json; tail -3 gpurun_out/aux_kernels.err
timeout 900 python -m pytest tests -m gpu -q -k "maintenance or threshold or fused" > gpurun_out/pytest_gpu15.log 2>&1; tail -3 gpurun_out/pytest_gpu15.log
