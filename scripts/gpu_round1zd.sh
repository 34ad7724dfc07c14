#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pug_smem --launch-skip 3 -c 1 -f -o gpurun_out/prof_pug_smem0_c5_r1zd python bench.py --config C5 --steps 1 --warmup 1 --cells 5000 --no-cpu-baseline > gpurun_out/ncu_full_ps5.log 2>&1
tail -1 gpurun_out/ncu_full_ps5.log | cut -c1-200
