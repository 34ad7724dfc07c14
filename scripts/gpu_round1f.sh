#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
tail -c 2700 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
for cfg in C3 C4 C5; do
  timeout 900 python bench.py --config $cfg --cells 20000 --steps 3 --warmup 1 --cpu-sample-seconds 4 > gpurun_out/bench_${cfg}_20k.json 2> gpurun_out/bench_${cfg}_20k.err
  python -c "
import json
j=json.load(open('gpurun_out/bench_${cfg}_20k.json')); print('$cfg value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e']['value'],'cpu',j['cpu_baseline']['value'],'frac',j['roofline']['frac']); print({k:round(v,3) for k,v in j['roofline']['per_kernel_ms'].items() if v>0.05})"
done
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_gene_eqc -s 1 -c 1 -f -o gpurun_out/prof_gene_eqc_v3 python bench.py --config C3 --steps 1 --warmup 1 --cells 5000 --no-cpu-baseline > gpurun_out/ncu_full_ge.log 2>&1
tail -1 gpurun_out/ncu_full_ge.log | cut -c1-200
