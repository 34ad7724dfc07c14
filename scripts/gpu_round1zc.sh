#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import json,sys
j=json.loads(open('$1').read().strip().splitlines()[-1]); pk=j['roofline']['per_kernel_ms']
print('$2 value',round(j['value']),'ms',round(j['ms_per_step'],2),'e2e',round(j['e2e']['value']),'frac',round(j['roofline']['frac'],4),'cpu',j['cpu_baseline'] and round(j['cpu_baseline']['value']), {k:round(v,2) for k,v in pk.items() if v>0.05})"; }
timeout 1200 python bench.py --config C3 --steps 3 --warmup 3 --cpu-sample-seconds 6 > gpurun_out/bench_C3_full.json 2> gpurun_out/bench_C3_full.err
show gpurun_out/bench_C3_full.json "C3 full"
timeout 1200 python bench.py --config C5 --steps 3 --warmup 3 --cpu-sample-seconds 6 > gpurun_out/bench_C5_full.json 2> gpurun_out/bench_C5_full.err
show gpurun_out/bench_C5_full.json "C5 full"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
