#!/bin/bash
# round 2n: whole GPU tier on the tree with the EM split, default bench line (C2 + other_configs), reference arm
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2n_pytest_gpu.log 2>&1 ) 2>&1 | grep real
tail -4 gpurun_out/r2n_pytest_gpu.log
( time timeout 1500 python bench.py > gpurun_out/r2n_bench_default.json 2> gpurun_out/r2n_bench_default.err ) 2>&1 | grep real
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2n_bench_default.json').read().strip().splitlines()[-1])
print('C2', round(j['value']), round(j['ms_per_step'],2), round(j['e2e']['value']), j['cpu_baseline'] and round(j['cpu_baseline']['value']), round(j['roofline']['frac'],4))
for k,v in j.get('other_configs',{}).items(): print(k, round(v['value']), round(v['ms_per_step'],2), round(v['e2e']['value']), v['cpu_baseline'] and round(v['cpu_baseline']['value']), round(v['roofline']['frac'],4), v.get('tie_census') and round(v['tie_census']['frac_molecules_changed_worst_case'],5))
PY
