#!/bin/bash
# round 2ah: CLI without the teardown (process_exits): CLI GPU tests + end-to-end timing
mkdir -p gpurun_out
( timeout 120 python -m pytest tests/test_host_cli.py -m gpu -q -x > gpurun_out/r2ah_pytest_cli.log 2>&1 ); tail -2 gpurun_out/r2ah_pytest_cli.log
timeout 60 python scripts/cli_bench.py C2 100000 cr-like 3 > gpurun_out/r2ah_cli_bench_c2_100kcells.json 2> gpurun_out/r2ah_cli_c2.err; python -c "
import json; j=json.loads(open('gpurun_out/r2ah_cli_bench_c2_100kcells.json').read().strip().splitlines()[-1]); print('C2', j.get('wall_s'), round(j.get('cells_per_s',0)), j.get('host_threads')); print(j.get('timing_log'))"
