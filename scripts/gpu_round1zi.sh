#!/bin/bash
# DRAM traffic of the resolve-family launches of one full-size C3 / C5 step (for roofline.traffic)
mkdir -p gpurun_out
for cfg in C3 C5; do
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_pug_smem|k_gene_eqc' -c 14 --csv --log-file gpurun_out/ncu_traffic_${cfg}.csv python bench.py --config $cfg --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_traffic_${cfg}.log 2>&1
tail -3 gpurun_out/ncu_traffic_${cfg}.csv | cut -c1-200
done
