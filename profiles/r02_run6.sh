#!/bin/bash
# round 2, GPU call 6 (8 GPUs): weak-scaling bench + strong_65536 on 8 ranks, BASELINE config #5 at 1,048,576 plants
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi8.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench6_n8.json 2> gpurun_out/bench6_n8.err; echo "bench8 rc=$?"; tail -c 300 gpurun_out/bench6_n8.err
python -c "
import json
d=json.load(open('gpurun_out/bench6_n8.json')); print(d['value'], d['full_step']['value'], d['e2e']['value'], d.get('strong_65536'))"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
    profiles/run_cfg5_maintenance.py > gpurun_out/cfg5_6_n8.json 2> gpurun_out/cfg5_6_n8.err; echo "cfg5 rc=$?"; tail -c 300 gpurun_out/cfg5_6_n8.err; cat gpurun_out/cfg5_6_n8.json
