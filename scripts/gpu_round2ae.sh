#!/bin/bash
# round 2ae: host-side speed-ups (formatter, parser, prefault, background teardown, fast exit): CLI GPU tests + CLI end-to-end timings
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_host_cli.py tests/test_infer.py -m gpu -q -x > gpurun_out/r2ae_pytest_cli.log 2>&1 ) 2>&1 | grep real
tail -3 gpurun_out/r2ae_pytest_cli.log
timeout 600 python scripts/cli_bench.py C2 100000 cr-like 3 > gpurun_out/r2ae_cli_bench_c2_100kcells.json 2> gpurun_out/r2ae_cli_c2.err; python -c "
import json; j=json.loads(open('gpurun_out/r2ae_cli_bench_c2_100kcells.json').read().strip().splitlines()[-1]); print('C2', j.get('wall_s'), round(j.get('cells_per_s',0)), j.get('host_threads')); print(j.get('timing_log'))"
timeout 600 python scripts/cli_bench.py C3 50000 parsimony 2 > gpurun_out/r2ae_cli_bench_c3_50kcells.json 2> gpurun_out/r2ae_cli_c3.err; python -c "
import json; j=json.loads(open('gpurun_out/r2ae_cli_bench_c3_50kcells.json').read().strip().splitlines()[-1]); print('C3', j.get('wall_s'), round(j.get('cells_per_s',0))); print(j.get('timing_log'))"
timeout 300 python scripts/host_stage_bench.py C2 100000 > gpurun_out/r2ae_host_stage_bench_c2_100kcells.json 2>/dev/null; cat gpurun_out/r2ae_host_stage_bench_c2_100kcells.json
