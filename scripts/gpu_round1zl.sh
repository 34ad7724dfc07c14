#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
tail -2 gpurun_out/pytest_gpu.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python -c "
import json
j=json.loads(open('gpurun_out/bench_c2.json').read().strip().splitlines()[-1]); print('C2', round(j['value']), round(j['ms_per_step'],3), round(j['e2e']['value']), j['roofline']['frac'])"
