#!/bin/bash
# round 2f: whole GPU tier (split parsimony path, infer / dump-eqclasses, device lists, RAD fixtures, full-size every-cell parity), default bench line
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2f_pytest_gpu.log 2>&1 ) 2>&1 | grep real
tail -5 gpurun_out/r2f_pytest_gpu.log
( time timeout 1500 python bench.py > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err ) 2>&1 | grep real
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2f_bench_default.json').read().strip().splitlines()[-1])
print('C2', round(j['value']), round(j['e2e']['value']), j['cpu_baseline'] and round(j['cpu_baseline']['value']), round(j['roofline']['frac'],4))
for k,v in j.get('other_configs',{}).items(): print(k, round(v['value']), round(v['e2e']['value']), v['cpu_baseline'] and round(v['cpu_baseline']['value']), round(v['roofline']['frac'],4), {a:round(b,2) for a,b in v['roofline']['per_kernel_ms'].items() if b>0.3})
PY
