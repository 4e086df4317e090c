#!/bin/bash
# usage: sweep_libs.sh variant1 variant2 ...   (libnps_b200_<variant>.so under nuclear-sim_b200/_lib); prints one line each
for v in "$@"; do
  lib="NPS_B200_LIB=$PWD/nuclear-sim_b200/_lib/libnps_b200_$v.so"
  echo -n "variant=$v  "
  env $lib python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.3e  e2e %.3e  ms/launch %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
