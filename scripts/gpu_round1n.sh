#!/bin/bash
# converged insert tail + lane-per-vertex PUG neighbour passes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_c2.json').read().strip().splitlines()[-1]); print('C2 value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']))"
for cfg in C3 C4 C5; do
  timeout 900 python bench.py --config $cfg --cells 20000 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_${cfg}_20k.json 2> gpurun_out/bench_${cfg}_20k.err
  python -c "
import json
j=json.loads(open('gpurun_out/bench_${cfg}_20k.json').read().strip().splitlines()[-1]); print('$cfg value',round(j['value']),'ms',round(j['ms_per_step'],2),'e2e',round(j['e2e']['value']))"
done
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_resolve_smem -s 2 -c 3 -f -o gpurun_out/prof_resolve_r1n python bench.py --steps 1 --warmup 1 --cells 20000 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-200
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_gene_eqc -s 1 -c 1 -f -o gpurun_out/prof_gene_eqc_r1n python bench.py --config C3 --steps 1 --warmup 1 --cells 5000 --no-cpu-baseline > gpurun_out/ncu_full_ge.log 2>&1
tail -1 gpurun_out/ncu_full_ge.log | cut -c1-200
