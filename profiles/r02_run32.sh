#!/bin/bash
# round 2, GPU call 32: config #5 loop with the rank's plants cut into independent batches resumed in turn (host work of one
# overlaps the launch of another): parts = 1 / 2 / 3 / 4 on one GPU; the device test of the interleaved mode
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "interleaved" 2>&1 | tail -3
for parts in 1 2 3 4 2; do
  timeout 600 python profiles/run_cfg5_maintenance.py --parts $parts 2> gpurun_out/cfg5_parts$parts.err | tee -a gpurun_out/cfg5_parts.jsonl | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('parts_per_gpu','plant_steps_per_s_whole_loop','seconds_total_max_over_ranks','seconds_host_numpy_max_over_ranks','work_orders_executed','mean_oil_level_pump0')})"; tail -2 gpurun_out/cfg5_parts$parts.err
done
