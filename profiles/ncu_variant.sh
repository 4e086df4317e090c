#!/bin/bash
# usage: ncu_variant.sh <variant> ; full ncu capture of one nps_step_kernel launch for libnps_b200_<variant>.so, raw metrics as CSV
v=$1
NPS_B200_LIB=$PWD/nuclear-sim_b200/_lib/libnps_b200_$v.so ncu --set full --clock-control none --import-source on -k regex:nps_step_kernel -s 3 -c 1 -f -o gpurun_out/prof_$v python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$v.log 2>&1
ncu -i gpurun_out/prof_$v.ncu-rep --page raw --csv > gpurun_out/prof_${v}_raw.csv 2>/dev/null
