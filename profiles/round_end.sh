#!/bin/bash
# end-of-round GPU evidence: tests, smoke, bench (ours + reference arm), ncu launch list and one full capture
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-300
python bench.py > gpurun_out/bench_n1.log 2>&1 && tail -1 gpurun_out/bench_n1.log | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nps_step_kernel -s 3 -c 1 -f -o gpurun_out/prof_step_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_step_final.ncu-rep --page raw --csv > gpurun_out/prof_step_final_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_step_final.ncu-rep --page source --csv > gpurun_out/prof_step_final_sass.csv 2>/dev/null
ls -la gpurun_out | tail -6
