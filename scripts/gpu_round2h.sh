#!/bin/bash
# round 2h: where do the EM configurations spend their time? ncu --set full of k_pug_smem<0> on C5 (parsimony-em) and k_gene_eqc on C4 (USA cr-like-em)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_pug_smem' --launch-skip 15 -c 1 -f -o gpurun_out/r2h_prof_c5 python bench.py --config C5 --steps 1 --warmup 1 --cells 10000 --no-cpu-baseline --no-others > gpurun_out/r2h_ncu_c5.log 2>&1
tail -1 gpurun_out/r2h_ncu_c5.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gene_eqc' --launch-skip 7 -c 1 -f -o gpurun_out/r2h_prof_c4 python bench.py --config C4 --steps 1 --warmup 1 --cells 6000 --no-cpu-baseline --no-others > gpurun_out/r2h_ncu_c4.log 2>&1
tail -1 gpurun_out/r2h_ncu_c4.log | cut -c1-200
