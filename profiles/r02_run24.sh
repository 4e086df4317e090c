#!/bin/bash
# round 2, GPU call 24: trip fixtures on the device (state parity + latch steps inside one fused launch)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -k "trip_ or fixture" > gpurun_out/pytest_gpu24.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu24.log
