#!/bin/bash
# DRAM traffic + warp instructions of one full-size device-resident step (C2, C3) -> gpurun_out/<tag>_profiles/ncu_traffic.json
mkdir -p gpurun_out
TAG=${1:-r2t}
NCU="ncu --clock-control none"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum
timeout 1500 $NCU --metrics $M -k regex:'k_resolve' -c 64 --csv --log-file gpurun_out/${TAG}_traffic_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-others > gpurun_out/${TAG}_traffic_c2.log 2>&1
timeout 1500 $NCU --metrics $M -k regex:'k_pug_build|k_pug_cover|k_pug_count|k_gene_eqc|k_resolve' -c 110 --csv --log-file gpurun_out/${TAG}_traffic_c3.csv python bench.py --config C3 --steps 1 --warmup 1 --no-cpu-baseline --no-others > gpurun_out/${TAG}_traffic_c3.log 2>&1
BUILD=$(python -c "import bench; print(bench.build_hash())")
PROFILES_OUT=gpurun_out/${TAG}_profiles python scripts/profiles_r2.py ${TAG} ${BUILD} 2>&1 | tail -8
