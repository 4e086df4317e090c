#!/bin/bash
# round 2, GPU call 29 (8 GPUs): bench.py --gpus 8 (weak scaling + strong_65536 + cfg5 key) and the cfg5 loop at 1,048,576 plants
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > gpurun_out/bench_n8_final.json 2> gpurun_out/bench_n8_final.err; echo "bench8 rc=$?"; tail -c 300 gpurun_out/bench_n8_final.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n8_final.json')); print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in list(v.items())[:5]}) for k,v in d.items() if k in ('value','full_step','e2e','cfg5_maintenance_loop','strong_65536','n_gpus')})"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 profiles/run_cfg5_maintenance.py > gpurun_out/cfg5_n8_final.json 2> gpurun_out/cfg5_n8_final.err; echo "cfg5 rc=$?"; cut -c1-700 gpurun_out/cfg5_n8_final.json
