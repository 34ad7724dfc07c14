#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python scripts/cli_bench.py C2 20000 cr-like 3 > gpurun_out/cli_bench_c2_20k.json 2> gpurun_out/cli_bench.err
cat gpurun_out/cli_bench_c2_20k.json | cut -c1-1500
tail -3 gpurun_out/cli_bench.err
timeout 900 python scripts/cli_bench.py C2 100000 cr-like 2 > gpurun_out/cli_bench_c2_100k.json 2> gpurun_out/cli_bench2.err
cat gpurun_out/cli_bench_c2_100k.json | cut -c1-1500
tail -3 gpurun_out/cli_bench2.err
