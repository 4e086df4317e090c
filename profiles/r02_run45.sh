#!/bin/bash
# round 2, GPU call 45: tube-scale carried fields fetched before the TSP update (shipped candidate) vs the previous build
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for rep in 1 2 3; do for v in _prev ""; do
  NPS_B200_LIB=$PWD/nuclear-sim_b200/_lib/libnps_b200$v.so timeout 600 python profiles/large_batch_ab.py 2>/dev/null | tee -a gpurun_out/tif_ab.jsonl | cut -c1-80
done; done
timeout 900 python -m pytest tests -m gpu -q -x -k "fixture or split_launch or fused" 2>&1 | tail -2
