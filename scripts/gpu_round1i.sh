#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
for cfg in C3 C4 C5; do
  timeout 900 python bench.py --config $cfg --cells 20000 --steps 3 --warmup 1 --cpu-sample-seconds 4 > gpurun_out/bench_${cfg}_20k.json 2> gpurun_out/bench_${cfg}_20k.err
  python -c "
import json
j=json.load(open('gpurun_out/bench_${cfg}_20k.json')); print('$cfg value',round(j['value']),'ms',round(j['ms_per_step'],2),'e2e',round(j['e2e']['value']),'cpu',round(j['cpu_baseline']['value']))"
done
