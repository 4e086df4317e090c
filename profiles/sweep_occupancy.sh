#!/bin/bash
# occupancy sweep: register budget (threads/SM) x plants per GPU, 8 fused substeps
mkdir -p gpurun_out
for v in "" b64m10 b64m14 b64m16 b128m5; do
  for n in 65536 132608 265216; do
    if [ -z "$v" ]; then lib=""; else lib="NPS_B200_LIB=$PWD/nuclear-sim_b200/_lib/libnps_b200_$v.so"; fi
    echo -n "variant=${v:-b64m7} plants=$n  "
    env $lib python bench.py --steps 6 --warmup 3 --no-cpu-baseline --plants-per-gpu $n 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.3e  e2e %.3e  ms/launch %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
  done
done
