#!/bin/bash
# round 2, GPU call 5: A/B of the large-batch kernel on ONE box (pow memo on/off, chemistry before/after the sink half)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for v in base nomemo chemlast nomemo_chemlast base2; do
  case $v in base|base2) unset NPS_B200_LIB;; *) export NPS_B200_LIB=$PWD/nuclear-sim_b200/_lib/libnps_b200_$v.so;; esac
  timeout 300 python bench.py --quick --no-small --steps 8 --warmup 3 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err; echo "$v rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/ab_$v.json'));print('$v', d['value'], d['full_step']['value'])"
done
unset NPS_B200_LIB
timeout 600 python -m pytest tests -m gpu -q -k "error_conventions or split" > gpurun_out/pytest_gpu5.log 2>&1; tail -3 gpurun_out/pytest_gpu5.log
