#!/bin/bash
# A/B: warp-aggregated distinct counter in table_insert (cr-like, C2)
mkdir -p gpurun_out
show() { python -c "
import json,sys
j=json.loads(open('$1').read().strip().splitlines()[-1]); pk=j['roofline']['per_kernel_ms']
print('$2 value',round(j['value']),'ms',round(j['ms_per_step'],3),'e2e',round(j['e2e']['value']), 'region', round(pk.get('resolve_region(wall)',0),3))"; }
cp alevin_fry_b200/libafq.so /tmp/libafq_a.so
for v in a b a b; do
  if [ $v = b ]; then cp alevin_fry_b200/libafq_pf.so alevin_fry_b200/libafq.so; else cp /tmp/libafq_a.so alevin_fry_b200/libafq.so; fi
  timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_$v.json 2> gpurun_out/bench_c2_$v.err
  show gpurun_out/bench_c2_$v.json "$v C2"
done
cp alevin_fry_b200/libafq_pf.so alevin_fry_b200/libafq.so
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "c2 or skew or edge or forced or c1" 2>&1 | tail -2
cp /tmp/libafq_a.so alevin_fry_b200/libafq.so
