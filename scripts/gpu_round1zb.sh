#!/bin/bash
# GPU tier incl. golden vectors; A/B: k_pug_smem<0> with a 56 KB arena x 4 CTAs/SM (64 registers) vs 72 KB x 3 (80 registers)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
show() { python -c "
import json,sys
j=json.loads(open('$1').read().strip().splitlines()[-1]); pk=j['roofline']['per_kernel_ms']
print('$2 value',round(j['value']),'ms',round(j['ms_per_step'],2),'e2e',round(j['e2e']['value']), {k:round(v,2) for k,v in pk.items() if v>0.05})"; }
cp alevin_fry_b200/libafq.so /tmp/libafq_a.so
for v in a b a b; do
  if [ $v = b ]; then cp alevin_fry_b200/libafq_v0b.so alevin_fry_b200/libafq.so; else cp /tmp/libafq_a.so alevin_fry_b200/libafq.so; fi
  for cfg in C3 C5; do
    timeout 900 python bench.py --config $cfg --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_${cfg}_$v.json 2> gpurun_out/bench_${cfg}_$v.err
    show gpurun_out/bench_${cfg}_$v.json "$v $cfg"
  done
done
cp /tmp/libafq_a.so alevin_fry_b200/libafq.so
