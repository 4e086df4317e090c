#!/bin/bash
# round 2, GPU call 16: where to switch from uncapped one-warp blocks to the 448 x 128-register shape; env frame_skip test
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python profiles/mid_batch.py > gpurun_out/mid_batch.json 2> gpurun_out/mid_batch.err; echo "rc=$?"; cat gpurun_out/mid_batch.json; tail -3 gpurun_out/mid_batch.err
timeout 600 python -m pytest tests -m gpu -q -k "env or checkpoint" > gpurun_out/pytest_gpu16.log 2>&1; tail -3 gpurun_out/pytest_gpu16.log
