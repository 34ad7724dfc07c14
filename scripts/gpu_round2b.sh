#!/bin/bash
# round 2b: dynamic cover queues + forked k_pug_smem lanes + conflict-free compaction (variant a) vs thread-count variants (b: no-spill
# thread counts for the big arenas, c: 56 KB x 4 x 192 threads for the small arena, d: both); then the new default bench line (other_configs)
bash scripts/gpu_ab.sh r2b "C3" a b c d
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "not full_size" > gpurun_out/r2b_pytest_gpu.log 2>&1
tail -3 gpurun_out/r2b_pytest_gpu.log
( time timeout 1500 python bench.py > gpurun_out/r2b_bench_default.json 2> gpurun_out/r2b_bench_default.err ) 2>&1 | grep real
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2b_bench_default.json').read().strip().splitlines()[-1])
print('C2', round(j['value']), round(j['e2e']['value']), j['cpu_baseline'] and round(j['cpu_baseline']['value']))
for k,v in j.get('other_configs',{}).items(): print(k, round(v['value']), round(v['e2e']['value']), v['cpu_baseline'] and round(v['cpu_baseline']['value']), round(v['roofline']['frac'],4))
PY
( time timeout 900 python bench.py --impl reference > gpurun_out/r2b_bench_reference.json 2> gpurun_out/r2b_bench_reference.err ) 2>&1 | grep real
tail -c 600 gpurun_out/r2b_bench_reference.json
