#!/bin/bash
# round-2 entry run: GPU tier after the ADVICE fixes, baselines (C2 default, C3), ncu --set full captures of the
# SHIPPED k_resolve_smem<2..4> and k_pug_smem<0..2>
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc > gpurun_out/r2a_nproc.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2a_pytest_gpu.log 2>&1
tail -3 gpurun_out/r2a_pytest_gpu.log
show() { python -c "
import json,sys
j=json.loads(open('$1').read().strip().splitlines()[-1]); pk=j['roofline']['per_kernel_ms']
print('$2 value',round(j['value']),'ms',round(j['ms_per_step'],2),'e2e',round(j['e2e']['value']),'frac',round(j['roofline']['frac'],4),'cpu',j['cpu_baseline'] and round(j['cpu_baseline']['value']), {k:round(v,2) for k,v in pk.items() if v>0.05})"; }
timeout 900 python bench.py > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err
show gpurun_out/r2a_bench_c2.json "C2 default"
timeout 1200 python bench.py --config C3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_C3_full.json 2> gpurun_out/r2a_bench_C3_full.err
show gpurun_out/r2a_bench_C3_full.json "C3 full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_resolve_smem --launch-skip 18 -c 6 -f -o gpurun_out/r2a_prof_resolve python bench.py --steps 1 --warmup 1 --cells 10000 --no-cpu-baseline > gpurun_out/r2a_ncu_resolve.log 2>&1
tail -1 gpurun_out/r2a_ncu_resolve.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pug_smem --launch-skip 12 -c 4 -f -o gpurun_out/r2a_prof_pug python bench.py --config C3 --steps 1 --warmup 1 --cells 10000 --no-cpu-baseline > gpurun_out/r2a_ncu_pug.log 2>&1
tail -1 gpurun_out/r2a_ncu_pug.log | cut -c1-200
ls -la gpurun_out | head -40
