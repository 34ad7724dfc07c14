#!/bin/bash
# A/B of the probe-window width (AFQ_UMI_WINDOW 8 vs 1) + fresh ncu captures (resolve, k_gene_eqc)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
tail -2 gpurun_out/pytest_gpu.log
for v in w8 w1 w8 w1; do
  if [ $v = w1 ]; then cp alevin_fry_b200/libafq.so /tmp/libafq_w8.so; cp alevin_fry_b200/libafq_w1.so alevin_fry_b200/libafq.so; fi
  timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_c2_$v.json 2> gpurun_out/bench_c2_$v.err
  python -c "
import json
d=json.loads(open('gpurun_out/bench_c2_$v.json').read().strip().splitlines()[-1]); print('$v value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']))"
  if [ $v = w1 ]; then cp /tmp/libafq_w8.so alevin_fry_b200/libafq.so; fi
done
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_resolve_smem -s 2 -c 3 -f -o gpurun_out/prof_resolve_w8 python bench.py --steps 1 --warmup 1 --cells 20000 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-200
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_gene_eqc -s 1 -c 1 -f -o gpurun_out/prof_gene_eqc_r1m python bench.py --config C3 --steps 1 --warmup 1 --cells 5000 --no-cpu-baseline > gpurun_out/ncu_full_ge.log 2>&1
tail -1 gpurun_out/ncu_full_ge.log | cut -c1-200
