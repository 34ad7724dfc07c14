#!/bin/bash
# round 2, final evidence on the frozen tree: whole GPU tier, smoke(), default bench line (C2 + other_configs), reference arm
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2final_pytest_gpu.log 2>&1 ) 2>&1 | grep real
tail -3 gpurun_out/r2final_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 1500 python bench.py > gpurun_out/r2final_bench_default.json 2> gpurun_out/r2final_bench_default.err ) 2>&1 | grep real
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2final_bench_default.json').read().strip().splitlines()[-1])
def show(k, v): print(k, 'value', round(v['value']), 'ms', round(v['ms_per_step'],2), 'e2e', round(v['e2e']['value']), round(v['e2e']['ms_per_step'],2), 'cpu', v['cpu_baseline'] and round(v['cpu_baseline']['value']), 'frac', round(v['roofline']['frac'],4), 'traffic', v['roofline']['traffic'], 'issue', v['roofline'].get('issue',{}).get('frac'), v.get('tie_census') and round(v['tie_census']['frac_molecules_changed_worst_case'],5))
show('C2', j)
for k,v in j.get('other_configs',{}).items(): show(k, v)
PY
( time timeout 900 python bench.py --impl reference > gpurun_out/r2final_bench_reference.json 2> gpurun_out/r2final_bench_reference.err ) 2>&1 | grep real
python -c "
import json
j=json.loads(open('gpurun_out/r2final_bench_reference.json').read().strip().splitlines()[-1]); print('reference', round(j['value']), j['cpu_baseline']['cores'], {k: round(v['value']) for k,v in j.get('other_configs',{}).items()})"
