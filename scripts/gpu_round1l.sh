#!/bin/bash
# flat (lane-per-alignment) phase 1 + 24-bit wire arrays: parity, bench, ncu
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_c2.json").read().strip().splitlines()[-1])
    print("C2 value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "h2d", d["e2e"]["h2d_bytes_per_step"], "roof", d["roofline"]["frac"], "cpu", round(d["cpu_baseline"]["value"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_c2.err").read()[-2000:])
PY
timeout 600 python bench.py --e2e-u32 --no-cpu-baseline > gpurun_out/bench_c2_u32.json 2> gpurun_out/bench_c2_u32.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_c2_u32.json').read().strip().splitlines()[-1]); print('u32 e2e', round(d['e2e']['value']))"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_resolve_smem -s 2 -c 3 -f -o gpurun_out/prof_resolve_flat python bench.py --steps 1 --warmup 1 --cells 20000 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
