#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
NC=20000 timeout 1200 python scripts/gpu_check.py cr-like,parsimony,cr-like-em,parsimony-em,parsimony-gene > gpurun_out/gpu_check_all.log 2>&1
tail -75 gpurun_out/gpu_check_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
for cfg in C3 C5 C4; do
  timeout 900 python bench.py --config $cfg --cells 20000 --steps 3 --warmup 1 --cpu-sample-seconds 4 > gpurun_out/bench_${cfg}_20k.json 2> gpurun_out/bench_${cfg}_20k.err
  tail -c 2500 gpurun_out/bench_${cfg}_20k.json; tail -3 gpurun_out/bench_${cfg}_20k.err
done
