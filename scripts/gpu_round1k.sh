#!/bin/bash
# v5 A/B: parity on GPU, then bench default vs AFQ_RESOLVE=5
mkdir -p gpurun_out
AFQ_RESOLVE=5 timeout 300 python scripts/gpu_check.py cr-like,trivial > gpurun_out/v5_check.log 2>&1
echo "v5 check exit $?" >> gpurun_out/v5_check.log
tail -3 gpurun_out/v5_check.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err
AFQ_RESOLVE=5 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err
for f in v3 v5; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("kernel_ms"))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/bench_$f.err").read()[-1500:])
PY
done
