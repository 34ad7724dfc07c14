#!/bin/bash
# round-2 evidence for profiles/: ncu --set full of the kernels as shipped, launch lists of the bench command, DRAM traffic of one full-size step
mkdir -p gpurun_out
TAG=${1:-r2p}
NCU="ncu --clock-control none"
# launch lists (gpu__time_duration.sum): C2 (headline) and C3
timeout 900 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/${TAG}_launches_c2.csv python bench.py --steps 2 --warmup 1 --cells 20000 --no-cpu-baseline --no-others > gpurun_out/${TAG}_launches_c2.log 2>&1
timeout 900 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/${TAG}_launches_c3.csv python bench.py --config C3 --steps 2 --warmup 1 --cells 20000 --no-cpu-baseline --no-others > gpurun_out/${TAG}_launches_c3.log 2>&1
# --set full
timeout 900 $NCU --set full --import-source on -k regex:k_resolve_smem --launch-skip 18 -c 6 -f -o gpurun_out/${TAG}_full_resolve python bench.py --steps 1 --warmup 1 --cells 10000 --no-cpu-baseline --no-others > gpurun_out/${TAG}_full_resolve.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:'k_pug_build|k_pug_cover|k_pug_count' --launch-skip 27 -c 9 -f -o gpurun_out/${TAG}_full_pug python bench.py --config C3 --steps 1 --warmup 1 --cells 10000 --no-cpu-baseline --no-others > gpurun_out/${TAG}_full_pug.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:'k_pug_back|k_em_cells' --launch-skip 24 -c 8 -f -o gpurun_out/${TAG}_full_em_c4 python bench.py --config C4 --steps 1 --warmup 1 --cells 8000 --no-cpu-baseline --no-others > gpurun_out/${TAG}_full_em_c4.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:'k_pug_back|k_em_cells' --launch-skip 24 -c 8 -f -o gpurun_out/${TAG}_full_em_c5 python bench.py --config C5 --steps 1 --warmup 1 --cells 10000 --no-cpu-baseline --no-others > gpurun_out/${TAG}_full_em_c5.log 2>&1
# DRAM traffic + warp instructions of the resolve family: every launch of the first steps is captured, profiles_r2.py keeps the 4th step (the timed one)
timeout 1500 $NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum -k regex:'k_resolve' -c 64 --csv --log-file gpurun_out/${TAG}_traffic_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-others > gpurun_out/${TAG}_traffic_c2.log 2>&1
timeout 1500 $NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum -k regex:'k_pug_build|k_pug_cover|k_pug_count|k_gene_eqc|k_resolve' -c 110 --csv --log-file gpurun_out/${TAG}_traffic_c3.csv python bench.py --config C3 --steps 1 --warmup 1 --no-cpu-baseline --no-others > gpurun_out/${TAG}_traffic_c3.log 2>&1
# summarise on the box (the reports together exceed what gpurun copies back); keep the parsimony report only
BUILD=$(python -c "import bench; print(bench.build_hash())")
PROFILES_OUT=gpurun_out/${TAG}_profiles python scripts/profiles_r2.py ${TAG} ${BUILD} 2>&1 | tail -8
rm -f gpurun_out/${TAG}_full_resolve.ncu-rep gpurun_out/${TAG}_full_em_c4.ncu-rep gpurun_out/${TAG}_full_em_c5.ncu-rep
ls -la gpurun_out | grep ${TAG}
