#!/bin/bash
# k_pug_smem v5 + EM back end without per-iteration searches
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
show() { python -c "
import json,sys
j=json.loads(open('$1').read().strip().splitlines()[-1]); pk=j['roofline']['per_kernel_ms']
print('$2 value',round(j['value']),'ms',round(j['ms_per_step'],2),'e2e',round(j['e2e']['value']),'frac',round(j['roofline']['frac'],4), {k:round(v,2) for k,v in pk.items() if v>0.05})"; }
for cfg in C3 C5 C4; do
  timeout 1200 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${cfg}_full.json 2> gpurun_out/bench_${cfg}_full.err
  show gpurun_out/bench_${cfg}_full.json "$cfg full"
done
AFQ_NO_PS=1 timeout 1200 python bench.py --config C4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C4_full_nops.json 2> gpurun_out/bench_C4_full_nops.err
show gpurun_out/bench_C4_full_nops.json "C4 full no-ps"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pug_smem -c 4 -f -o gpurun_out/prof_pug_smem_c4_r1x python bench.py --config C4 --steps 1 --warmup 1 --cells 5000 --no-cpu-baseline > gpurun_out/ncu_full_ps4.log 2>&1
tail -1 gpurun_out/ncu_full_ps4.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pug_smem -c 4 -f -o gpurun_out/prof_pug_smem_c5_r1x python bench.py --config C5 --steps 1 --warmup 1 --cells 5000 --no-cpu-baseline > gpurun_out/ncu_full_ps5.log 2>&1
tail -1 gpurun_out/ncu_full_ps5.log | cut -c1-200
