#!/bin/bash
mkdir -p gpurun_out
AFQ_DEBUG_CTL=1 timeout 900 python bench.py --config C3 --steps 1 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/r2e_bench_C3.json 2> gpurun_out/r2e_bench_C3.err
grep "afq ctl" gpurun_out/r2e_bench_C3.err | sort | uniq -c | head
AFQ_DEBUG_CTL=1 timeout 900 python bench.py --config C3 --cells 60000 --steps 1 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/r2e_bench_C3_60k.json 2> gpurun_out/r2e_bench_C3_60k.err
grep "afq ctl" gpurun_out/r2e_bench_C3_60k.err | sort | uniq -c | head
python -c "
import json
for f in ('gpurun_out/r2e_bench_C3.json','gpurun_out/r2e_bench_C3_60k.json'):
    j=json.loads(open(f).read().strip().splitlines()[-1]); print({k:round(v,2) for k,v in j['roofline']['per_kernel_ms'].items() if v>0.05})"
timeout 600 python -m pytest tests/test_infer.py -m gpu -q -x > gpurun_out/r2e_pytest_infer.log 2>&1
tail -5 gpurun_out/r2e_pytest_infer.log
