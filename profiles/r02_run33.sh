#!/bin/bash
# round 2, GPU call 33: interleaved (one host thread, generators) vs threaded (one host thread per batch) for 2 parts
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for mode in --one-thread "" --one-thread ""; do
  timeout 600 python profiles/run_cfg5_maintenance.py --parts 2 $mode 2> gpurun_out/cfg5_thr.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('mode=$mode', {k:d.get(k) for k in ('parts_per_gpu','plant_steps_per_s_whole_loop','seconds_total_max_over_ranks','work_orders_executed','mean_oil_level_pump0')})"; tail -2 gpurun_out/cfg5_thr.err
done
