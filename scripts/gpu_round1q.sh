#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python scripts/cli_bench.py C2 100000 cr-like 3 > gpurun_out/cli_bench_c2_100k.json 2> gpurun_out/cli_bench2.err
cat gpurun_out/cli_bench_c2_100k.json | cut -c1-1500
tail -3 gpurun_out/cli_bench2.err
timeout 900 python scripts/cli_bench.py C3 50000 parsimony 2 > gpurun_out/cli_bench_c3_50k.json 2> gpurun_out/cli_bench3.err
cat gpurun_out/cli_bench_c3_50k.json | cut -c1-1200
# full-size single-GPU shares of C3 and C5 (bench.py defaults for those configs)
timeout 1200 python bench.py --config C3 --steps 3 --warmup 3 --cpu-sample-seconds 6 > gpurun_out/bench_C3_full.json 2> gpurun_out/bench_C3_full.err
python -c "
import json
j=json.loads(open('gpurun_out/bench_C3_full.json').read().strip().splitlines()[-1]); print('C3 full value',round(j['value']),'ms',round(j['ms_per_step'],2),'e2e',round(j['e2e']['value']),'cpu',j['cpu_baseline'] and round(j['cpu_baseline']['value']), 'roof', j['roofline']['frac'])"
timeout 1200 python bench.py --config C5 --steps 3 --warmup 3 --cpu-sample-seconds 6 > gpurun_out/bench_C5_full.json 2> gpurun_out/bench_C5_full.err
python -c "
import json
j=json.loads(open('gpurun_out/bench_C5_full.json').read().strip().splitlines()[-1]); print('C5 full value',round(j['value']),'ms',round(j['ms_per_step'],2),'e2e',round(j['e2e']['value']),'cpu',j['cpu_baseline'] and round(j['cpu_baseline']['value']), 'roof', j['roofline']['frac'])"
