#!/bin/bash
# round 2x: record form of phase 1 in the cr-like kernels: parity tests, C2 / C1-like timings, per-kernel times
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "record_form or mini_c2 or c1_tiny or edge or skewed or forced or giant or pipelined or flat_alignment or tiny_cells or (full_size and C2)" > gpurun_out/r2x_pytest.log 2>&1 ) 2>&1 | grep real
tail -3 gpurun_out/r2x_pytest.log
show() { python - "$1" "$2" <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
pk=j['roofline']['per_kernel_ms']
print(sys.argv[2], 'value', round(j['value']), 'ms', round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value']), 'e2e ms', round(j['e2e']['ms_per_step'],2), 'frac', round(j['roofline']['frac'],4), {k:round(v,2) for k,v in pk.items() if v>0.05})
PY
}
for i in 1 2; do
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/r2x_C2_$i.json 2>gpurun_out/r2x_C2_$i.err; show gpurun_out/r2x_C2_$i.json "C2 record form"
done
timeout 300 python bench.py --config C3 --steps 5 --warmup 3 --no-cpu-baseline --no-others --e2e-batches 4 > gpurun_out/r2x_C3_nb4.json 2>/dev/null; show gpurun_out/r2x_C3_nb4.json "C3 nb=4"
timeout 300 python bench.py --config C5 --steps 5 --warmup 3 --no-cpu-baseline --no-others --e2e-batches 4 > gpurun_out/r2x_C5_nb4.json 2>/dev/null; show gpurun_out/r2x_C5_nb4.json "C5 nb=4"
