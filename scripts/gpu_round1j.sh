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
# DRAM traffic of the resolve family for one full-size C2 step (8 launches), and the launch list of the bench command
AFQ_NO_LANES=1 timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_resolve -s 8 -c 8 --csv --log-file gpurun_out/traffic_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1
tail -2 gpurun_out/ncu_traffic.log | cut -c1-200
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
tail -1 gpurun_out/ncu_launch_bench.log | cut -c1-200
