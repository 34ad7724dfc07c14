#!/bin/bash
# round 2k: where does k_pug_back spend its time on C4 (USA cr-like-em)?
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_pug_back' --launch-skip 12 -c 1 -f -o gpurun_out/r2k_prof_back_c4 python bench.py --config C4 --steps 1 --warmup 1 --cells 8000 --no-cpu-baseline --no-others > gpurun_out/r2k_ncu.log 2>&1
tail -1 gpurun_out/r2k_ncu.log | cut -c1-200
