#!/bin/bash
# round 2, GPU call 28: final single-GPU verification: smoke, GPU suite, bench (own arm with the cfg5 key, reference arm)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi28.txt
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke28.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke28.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu28.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu28.log
timeout 900 python bench.py > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_n1_final.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1_final.json')); print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in list(v.items())[:6]}) for k,v in d.items() if k in ('value','full_step','e2e','cfg5_maintenance_loop','roofline','cpu_baseline','small_batch','ms_per_step')})"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref_final.json
