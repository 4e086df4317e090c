#!/bin/bash
# round 2, GPU call 38: next-stage carried fields fetched one turbine stage ahead (volatile loads into registers) vs the previous build
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "fixture or split_launch or fused" 2>&1 | tail -2
for rep in 1 2; do for v in _prev ""; do
  NPS_B200_LIB=$PWD/nuclear-sim_b200/_lib/libnps_b200$v.so timeout 600 python profiles/large_batch_ab.py 2>/dev/null | tee -a gpurun_out/stage_ahead_ab.jsonl
done; done
