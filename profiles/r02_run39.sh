#!/bin/bash
# round 2, GPU call 39: more first-touch round trips merged in the turbine body (stage outlet temperatures with the thermal
# tracker's group, bearing loads with the rotor group, the four bearing records together) vs stage-ahead only vs before
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for rep in 1; do for v in _sa "" _sa ""; do
  NPS_B200_LIB=$PWD/nuclear-sim_b200/_lib/libnps_b200$v.so timeout 600 python profiles/large_batch_ab.py 2>/dev/null | tee -a gpurun_out/touch_merge_ab.jsonl
done; done
timeout 900 python -m pytest tests -m gpu -q -x -k "fixture or split_launch or fused" 2>&1 | tail -2
