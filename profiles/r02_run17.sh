#!/bin/bash
# round 2, GPU call 17: ring-coupled halves in the two-threads-per-plant kernel (bitwise tests, then rates); frame_skip test
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "split or env or checkpoint or fused or embedded" > gpurun_out/pytest_gpu17.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu17.log
timeout 600 python profiles/small_batch.py > gpurun_out/small_batch_ring.json 2> gpurun_out/small_batch.err; echo "rc=$?"; cat gpurun_out/small_batch_ring.json; tail -3 gpurun_out/small_batch.err
