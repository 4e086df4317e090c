#!/bin/bash
# round 2, GPU call 43: stage outlet temperatures touched together with the thermal tracker's own arrays (thm) vs shipped
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for rep in 1 2 3; do for v in "" _thm; do
  NPS_B200_LIB=$PWD/nuclear-sim_b200/_lib/libnps_b200$v.so timeout 600 python profiles/large_batch_ab.py 2>/dev/null | tee -a gpurun_out/thm_ab.jsonl | cut -c1-80
done; done
