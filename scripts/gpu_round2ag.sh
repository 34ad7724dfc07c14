#!/bin/bash
# round 2ag: CLI end to end with CUDA's default 8 hardware queues (the CLI is host-bound; 32 queues cost 1-2 s of context creation)
mkdir -p gpurun_out
timeout 200 python scripts/cli_bench.py C2 100000 cr-like 3 > gpurun_out/r2ag_cli_bench_c2_100kcells.json 2> gpurun_out/r2ag_cli_c2.err; python -c "
import json; j=json.loads(open('gpurun_out/r2ag_cli_bench_c2_100kcells.json').read().strip().splitlines()[-1]); print('C2', j.get('wall_s'), round(j.get('cells_per_s',0)), j.get('host_threads')); print(j.get('timing_log'))"
timeout 100 python scripts/cli_bench.py C3 50000 parsimony 2 > gpurun_out/r2ag_cli_bench_c3_50kcells.json 2> gpurun_out/r2ag_cli_c3.err; python -c "
import json; j=json.loads(open('gpurun_out/r2ag_cli_bench_c3_50kcells.json').read().strip().splitlines()[-1]); print('C3', j.get('wall_s'), round(j.get('cells_per_s',0))); print(j.get('timing_log'))"
