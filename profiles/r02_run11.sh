#!/bin/bash
# round 2, GPU call 11: how busy each half of the split kernel is (tuning build with clock64 accounting)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
NPS_B200_LIB=$PWD/nuclear-sim_b200/_lib/libnps_b200_splittiming.so timeout 300 python profiles/split_timing.py > gpurun_out/split_timing.log 2>&1; cat gpurun_out/split_timing.log | tail -12
