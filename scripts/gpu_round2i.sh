#!/bin/bash
# round 2i: split path for the EM resolutions (k_pug_build -> k_pug_cover* -> k_pug_back) + bulk-copy loads in ps_cell: whole GPU tier, default bench
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2i_pytest_gpu.log 2>&1 ) 2>&1 | grep real
tail -5 gpurun_out/r2i_pytest_gpu.log
( time timeout 1500 python bench.py > gpurun_out/r2i_bench_default.json 2> gpurun_out/r2i_bench_default.err ) 2>&1 | grep real
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2i_bench_default.json').read().strip().splitlines()[-1])
print('C2', round(j['value']), round(j['e2e']['value']), j['cpu_baseline'] and round(j['cpu_baseline']['value']), round(j['roofline']['frac'],4))
for k,v in j.get('other_configs',{}).items(): print(k, round(v['value']), round(v['ms_per_step'],2), round(v['e2e']['value']), v['cpu_baseline'] and round(v['cpu_baseline']['value']), round(v['roofline']['frac'],4), {a:round(b,2) for a,b in v['roofline']['per_kernel_ms'].items() if b>0.3})
PY
for cfg in C4 C5; do
AFQ_NO_PS_SPLIT=1 timeout 900 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/r2i_bench_${cfg}_nosplit.json 2>/dev/null
python -c "
import json
j=json.loads(open('gpurun_out/r2i_bench_${cfg}_nosplit.json').read().strip().splitlines()[-1]); print('$cfg no-split ms', round(j['ms_per_step'],2))"
done
