#!/bin/bash
# k_gene_eqc occupancy A/B: CTAs/SM 2,3,4 (default build), 5,6 (builds with fewer registers)
mkdir -p gpurun_out
cp alevin_fry_b200/libafq.so /tmp/libafq_ge4.so
for m in 2 3 4 5 6; do
  if [ $m -le 4 ]; then cp /tmp/libafq_ge4.so alevin_fry_b200/libafq.so; else cp alevin_fry_b200/libafq_ge$m.so alevin_fry_b200/libafq.so; fi
  for cfg in C3 C5; do
    AFQ_GE_OCC=$m timeout 600 python bench.py --config $cfg --cells 20000 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_${cfg}_occ$m.json 2> gpurun_out/bench_${cfg}_occ$m.err
    python -c "
import json
j=json.loads(open('gpurun_out/bench_${cfg}_occ$m.json').read().strip().splitlines()[-1]); print('occ$m $cfg value',round(j['value']),'ms',round(j['ms_per_step'],2))"
  done
done
cp /tmp/libafq_ge4.so alevin_fry_b200/libafq.so
