#!/bin/bash
# A/B harness: usage gpu_ab.sh TAG "cfgs" variants... ; variant x = alevin_fry_b200/libafq_x.so ("a" = the built libafq.so)
TAG=$1; CFGS=$2; shift 2
mkdir -p gpurun_out
cp alevin_fry_b200/libafq.so /tmp/libafq_a.so
show() { python -c "
import json,sys
j=json.loads(open('$1').read().strip().splitlines()[-1]); pk=j['roofline']['per_kernel_ms']
print('$2 value',round(j['value']),'ms',round(j['ms_per_step'],2),'e2e',round(j['e2e']['value']),'kern_ms',round(j['roofline']['kernel_ms_per_step'],2), {k:round(v,2) for k,v in pk.items() if v>0.05})"; }
for rep in 1 2; do
for v in "$@"; do
  if [ $v = a ]; then cp /tmp/libafq_a.so alevin_fry_b200/libafq.so; else cp alevin_fry_b200/libafq_$v.so alevin_fry_b200/libafq.so; fi
  for cfg in $CFGS; do
    timeout 900 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/${TAG}_bench_${cfg}_${v}_$rep.json 2> gpurun_out/${TAG}_bench_${cfg}_${v}_$rep.err
    show gpurun_out/${TAG}_bench_${cfg}_${v}_$rep.json "$v $cfg #$rep"
  done
done
done
cp /tmp/libafq_a.so alevin_fry_b200/libafq.so
