#!/bin/bash
# round 2, GPU call 41: feedwater pump loop - a touch group at the top of the lubrication update (fwt), the next pump's first six
# fields fetched one pump ahead and stored where the code reads them (fwa), both (fwta), against the shipped build
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for rep in 1 2; do for v in "" _fwt _fwa _fwta; do
  NPS_B200_LIB=$PWD/nuclear-sim_b200/_lib/libnps_b200$v.so timeout 600 python profiles/large_batch_ab.py 2>/dev/null | tee -a gpurun_out/fw_ab.jsonl
done; done
