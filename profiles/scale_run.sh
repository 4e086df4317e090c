#!/bin/bash
# weak-scaling bench at N = 1, 2, 4, 8 on one 8-GPU box (same launch line the driver uses)
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n1.log 2>&1; tail -1 gpurun_out/scale_n1.log | cut -c1-160
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/scale_n$n.log 2>&1
  tail -1 gpurun_out/scale_n$n.log | cut -c1-160
done
