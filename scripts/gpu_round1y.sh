#!/bin/bash
# round-1 evidence run: GPU tier, default bench (C2), full C3 / C5 / C4, launch list + ncu capture of k_pug_smem<0>
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
show() { python -c "
import json,sys
j=json.loads(open('$1').read().strip().splitlines()[-1]); pk=j['roofline']['per_kernel_ms']
print('$2 value',round(j['value']),'ms',round(j['ms_per_step'],2),'e2e',round(j['e2e']['value']),'frac',round(j['roofline']['frac'],4),'cpu',j['cpu_baseline'] and round(j['cpu_baseline']['value']), {k:round(v,2) for k,v in pk.items() if v>0.05})"; }
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
show gpurun_out/bench_c2.json "C2 default"
timeout 1200 python bench.py --config C3 --steps 3 --warmup 3 --cpu-sample-seconds 6 > gpurun_out/bench_C3_full.json 2> gpurun_out/bench_C3_full.err
show gpurun_out/bench_C3_full.json "C3 full"
for cfg in C5 C4; do
  timeout 1200 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${cfg}_full.json 2> gpurun_out/bench_${cfg}_full.err
  show gpurun_out/bench_${cfg}_full.json "$cfg full"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_bench_c3.csv python bench.py --config C3 --cells 20000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_c3.log 2>&1
tail -2 gpurun_out/ncu_launches_c3.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pug_smem --launch-skip 3 -c 1 -f -o gpurun_out/prof_pug_smem0_r1y python bench.py --config C3 --steps 1 --warmup 1 --cells 5000 --no-cpu-baseline > gpurun_out/ncu_full_ps.log 2>&1
tail -1 gpurun_out/ncu_full_ps.log | cut -c1-200
ls -la gpurun_out | head -30
