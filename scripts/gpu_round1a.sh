#!/bin/bash
# bench + ncu launch list + ncu full capture of the resolve kernels + cr-like GPU tests
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "(cr-like or trivial or pipelined) and not em and not edge and not skew and not forced" > gpurun_out/pytest_crlike.log 2>&1
tail -5 gpurun_out/pytest_crlike.log
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
tail -c 3000 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --cells 20000 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_resolve_smem -s 2 -c 3 -f -o gpurun_out/prof_resolve python bench.py --steps 1 --warmup 1 --cells 20000 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
