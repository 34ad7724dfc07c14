#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "na8 or cli or subset" > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
tail -c 900 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
timeout 900 python bench.py --e2e-offsets --no-cpu-baseline > gpurun_out/bench_c2_offsets.json 2> gpurun_out/bench_c2_offsets.err
python -c "
import json
j=json.load(open('gpurun_out/bench_c2_offsets.json')); print('OFFSETS e2e',j['e2e'])"
