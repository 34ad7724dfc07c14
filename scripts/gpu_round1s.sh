#!/bin/bash
# re-entry baseline: GPU parity tier, default bench, fresh per-line ncu captures of both hot kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_c2.json').read().strip().splitlines()[-1]); print('C2 value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']))"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_resolve_smem -s 2 -c 3 -f -o gpurun_out/prof_resolve_r1s python bench.py --steps 1 --warmup 1 --cells 20000 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gene_eqc -s 1 -c 1 -f -o gpurun_out/prof_gene_eqc_r1s python bench.py --config C3 --steps 1 --warmup 1 --cells 5000 --no-cpu-baseline > gpurun_out/ncu_full_ge.log 2>&1
tail -1 gpurun_out/ncu_full_ge.log | cut -c1-200
ls -la gpurun_out
