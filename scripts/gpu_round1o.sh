#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_c2.json').read().strip().splitlines()[-1]); print('C2 value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']))"
timeout 900 python scripts/cli_bench.py C2 20000 cr-like 3 > gpurun_out/cli_bench_c2_20k.json 2> gpurun_out/cli_bench.err
cat gpurun_out/cli_bench_c2_20k.json | cut -c1-1500
tail -3 gpurun_out/cli_bench.err
