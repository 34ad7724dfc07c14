#!/bin/bash
# round 2q: split-aware arena estimate of k_pug_build — hand-back counts and C3/C4/C5 times
mkdir -p gpurun_out
for c in C3 C4 C5; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/r2q_bench_$c.json 2> gpurun_out/r2q_bench_$c.err
  python - <<PY
import json
j=json.loads(open('gpurun_out/r2q_bench_$c.json').read().strip().splitlines()[-1])
print('$c', round(j['value']), round(j['ms_per_step'],2), round(j['e2e']['value']), round(j['roofline']['frac'],4), j['roofline'].get('regions_ms'))
PY
  AFQ_DEBUG_CTL=1 timeout 600 python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline --no-others 2>&1 >/dev/null | grep -v "^\[" | tail -3 | cut -c1-400
  AFQ_DEBUG_CTL=1 timeout 600 python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline --no-others 2> gpurun_out/r2q_ctl_$c.err >/dev/null
  tail -2 gpurun_out/r2q_ctl_$c.err | cut -c1-500
done
