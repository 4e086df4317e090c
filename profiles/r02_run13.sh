#!/bin/bash
# round 2, GPU call 13: power function A/B on one box (table-driven vs series form, pow(1, y) shortcut), pow accuracy test
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "pow or split_launch_shape_on or cuda_matches_reference_fixture" > gpurun_out/pytest_gpu13.log 2>&1; tail -4 gpurun_out/pytest_gpu13.log
for v in base powseries powone base2; do
  case $v in base|base2) unset NPS_B200_LIB;; *) export NPS_B200_LIB=$PWD/nuclear-sim_b200/_lib/libnps_b200_$v.so;; esac
  timeout 300 python bench.py --quick --steps 8 --warmup 3 > gpurun_out/pw_$v.json 2> gpurun_out/pw_$v.err; echo "$v rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/pw_$v.json'));sb=d['small_batch'];print('$v', '%.4e %.4e' % (d['value'], d['full_step']['value']), sb['plants_4096'], sb['plants_16384'], sb['one_thread_per_plant']['plants_32768'])"
done
