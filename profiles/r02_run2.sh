#!/bin/bash
# round 2, GPU call 2: GPU test-suite after the monitor / pow-memo / small-batch launch-shape changes, quick bench of build variants
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
tail -8 gpurun_out/pytest_gpu2.log
for v in base unroll inl; do
  if [ "$v" = base ]; then unset NPS_B200_LIB; else export NPS_B200_LIB=$PWD/nuclear-sim_b200/_lib/libnps_b200_$v.so; fi
  timeout 300 python bench.py --quick --steps 6 --warmup 3 > gpurun_out/quick_$v.json 2> gpurun_out/quick_$v.err; echo "$v rc=$?"
done
