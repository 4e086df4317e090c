#!/bin/bash
# round 2, GPU call 27: config #5 maintenance loop on one GPU, numpy-column bookkeeping vs the native work-order table
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for b in columnar native columnar native; do
  timeout 600 python profiles/run_cfg5_maintenance.py --bookkeeping $b 2> gpurun_out/cfg5_$b.err | tee -a gpurun_out/cfg5_bookkeeping.jsonl | cut -c1-900; tail -2 gpurun_out/cfg5_$b.err
done
timeout 600 python -m pytest tests -m gpu -q -k "maint or columnar" 2>&1 | tail -3
