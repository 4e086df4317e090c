#!/bin/bash
# e2e jitter hunt: N bench runs with the pipeline trace on; per run the bench line's value / e2e and the spread of the
# per-step device intervals (input copy, copy-done -> kernel start, kernel, result copy).
mkdir -p gpurun_out
for i in $(seq 1 ${1:-6}); do
  NPS_PIPE_TRACE=1 python bench.py --no-cpu-baseline > gpurun_out/trace_$i.out 2> gpurun_out/trace_$i.err
  python - "$i" <<'PY'
import json, re, sys
i = sys.argv[1]
d = json.loads(open(f"gpurun_out/trace_{i}.out").read().strip().splitlines()[-1])
rows = [list(map(float, re.findall(r"(?:in|gap|kernel|out) ([0-9.]+)", l))) for l in open(f"gpurun_out/trace_{i}.err") if l.startswith("[nps pipe]")]
rows = [r for r in rows if len(r) == 4][5:]   # skip the warm-up launches
col = lambda k: [r[k] for r in rows]
fmt = lambda v: f"mean {sum(v)/len(v):6.2f} max {max(v):6.2f}"
print(f"run {i}: value {d['value']:.4g} e2e {d['e2e']['value']:.4g} | in {fmt(col(0))} | gap {fmt(col(1))} | kernel {fmt(col(2))} | out {fmt(col(3))}")
PY
done
