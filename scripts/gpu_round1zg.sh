#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import json,sys
j=json.loads(open('$1').read().strip().splitlines()[-1]); pk=j['roofline']['per_kernel_ms']
print('$2 value',round(j['value']),'ms',round(j['ms_per_step'],2),'e2e',round(j['e2e']['value']),'frac',round(j['roofline']['frac'],4), {k:round(v,2) for k,v in pk.items() if v>0.05})"; }
for cfg in C3 C5; do
timeout 1200 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${cfg}_full.json 2> gpurun_out/bench_${cfg}_full.err
show gpurun_out/bench_${cfg}_full.json "$cfg full"
done
