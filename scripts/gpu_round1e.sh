#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
tail -c 2700 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
AFQ_NO_LANES=1 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_c2_nolanes.json 2> gpurun_out/bench_c2_nolanes.err
python -c "
import json
j=json.load(open('gpurun_out/bench_c2_nolanes.json')); print('NO_LANES value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e']['value']); print(j['roofline']['per_kernel_ms'])"
NC=20000 timeout 1200 python scripts/gpu_check.py cr-like,parsimony,cr-like-em,parsimony-em > gpurun_out/gpu_check_all.log 2>&1
grep -c "^\[OK\]" gpurun_out/gpu_check_all.log; grep -E "FAIL|SOME|ALL OK" gpurun_out/gpu_check_all.log; grep -E "^(cr-like|parsimony|cr-like-em|parsimony-em):|k_gene_eqc|host API" gpurun_out/gpu_check_all.log
AFQ_NO_LANES=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_resolve_smem -s 2 -c 3 -f -o gpurun_out/prof_resolve_v4 python bench.py --steps 1 --warmup 1 --cells 20000 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_gene_eqc -s 1 -c 1 -f -o gpurun_out/prof_gene_eqc_v2 python bench.py --config C3 --steps 1 --warmup 1 --cells 5000 --no-cpu-baseline > gpurun_out/ncu_full_ge.log 2>&1
tail -1 gpurun_out/ncu_full_ge.log | cut -c1-200
