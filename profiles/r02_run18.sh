#!/bin/bash
# round 2, GPU call 18: ring depth / shared-memory carve-out variants of the two-threads-per-plant kernel against the barrier version
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for v in barrier ring2 ring3 ring4c100; do
  NPS_B200_LIB=$PWD/nuclear-sim_b200/_lib/libnps_b200_$v.so timeout 600 python profiles/small_batch.py > gpurun_out/small_batch_$v.json 2> gpurun_out/small_batch_$v.err; echo "$v rc=$?"; cat gpurun_out/small_batch_$v.json
done
