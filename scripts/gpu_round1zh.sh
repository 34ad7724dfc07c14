#!/bin/bash
# final round-1 evidence: GPU tier, smoke, default bench (C2) + reference arm, full C3 / C5 / C4 shares
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
tail -2 gpurun_out/pytest_gpu.log
show() { python -c "
import json,sys
j=json.loads(open('$1').read().strip().splitlines()[-1]); pk=j['roofline']['per_kernel_ms']
print('$2 value',round(j['value']),'ms',round(j['ms_per_step'],2),'e2e',round(j['e2e']['value']),'frac',round(j['roofline']['frac'],4),'cpu',j['cpu_baseline'] and round(j['cpu_baseline']['value']), {k:round(v,2) for k,v in pk.items() if v>0.05})"; }
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
show gpurun_out/bench_c2.json "C2 default"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
cut -c1-400 gpurun_out/bench_reference.json
timeout 1200 python bench.py --config C3 --steps 3 --warmup 3 --cpu-sample-seconds 6 > gpurun_out/bench_C3_full.json 2> gpurun_out/bench_C3_full.err
show gpurun_out/bench_C3_full.json "C3 full"
for cfg in C5 C4; do
  timeout 1200 python bench.py --config $cfg --steps 3 --warmup 3 --cpu-sample-seconds 6 > gpurun_out/bench_${cfg}_full.json 2> gpurun_out/bench_${cfg}_full.err
  show gpurun_out/bench_${cfg}_full.json "$cfg full"
done
