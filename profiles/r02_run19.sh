#!/bin/bash
# round 2, GPU call 19: shared-memory carve-out (L1 size) of the small-batch kernels; each variant measured twice
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for rep in 1 2; do
for v in barrier ring2 r2c28 r2c28_ow14; do
  NPS_B200_LIB=$PWD/nuclear-sim_b200/_lib/libnps_b200_$v.so timeout 600 python profiles/small_batch.py > gpurun_out/small_batch_${v}_$rep.json 2> gpurun_out/small_batch_$v.err; echo "$v rc=$?"; cat gpurun_out/small_batch_${v}_$rep.json
done
done
