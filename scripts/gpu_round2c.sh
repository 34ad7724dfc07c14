#!/bin/bash
# round 2c: the SPLIT parsimony path (k_pug_build -> flat k_pug_cover* -> k_pug_count) vs the single-kernel k_pug_smem (AFQ_NO_PS_SPLIT=1)
mkdir -p gpurun_out
show() { python -c "
import json,sys
j=json.loads(open('$1').read().strip().splitlines()[-1]); pk=j['roofline']['per_kernel_ms']
print('$2 value',round(j['value']),'ms',round(j['ms_per_step'],2),'e2e',round(j['e2e']['value']),'kern_ms',round(j['roofline']['kernel_ms_per_step'],2), {k:round(v,2) for k,v in pk.items() if v>0.05})"; }
timeout 1200 python -m pytest tests -m gpu -q -x -k "not full_size" > gpurun_out/r2c_pytest_gpu.log 2>&1
tail -3 gpurun_out/r2c_pytest_gpu.log
for rep in 1 2; do
  timeout 900 python bench.py --config C3 --steps 3 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/r2c_bench_C3_split_$rep.json 2> gpurun_out/r2c_bench_C3_split_$rep.err
  show gpurun_out/r2c_bench_C3_split_$rep.json "split C3 #$rep"
  AFQ_NO_PS_SPLIT=1 timeout 900 python bench.py --config C3 --steps 3 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/r2c_bench_C3_single_$rep.json 2> gpurun_out/r2c_bench_C3_single_$rep.err
  show gpurun_out/r2c_bench_C3_single_$rep.json "single C3 #$rep"
done
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full_size and C3" > gpurun_out/r2c_pytest_full_c3.log 2>&1
tail -3 gpurun_out/r2c_pytest_full_c3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_pug_build|k_pug_cover|k_pug_count' --launch-skip 27 -c 9 -f -o gpurun_out/r2c_prof_split python bench.py --config C3 --steps 1 --warmup 1 --cells 10000 --no-cpu-baseline --no-others > gpurun_out/r2c_ncu_split.log 2>&1
tail -1 gpurun_out/r2c_ncu_split.log | cut -c1-200
