#!/bin/bash
# round 2z: two host pipelines (consecutive batches overlap on the device): GPU tier, e2e with 2 pipelines vs 1
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2z_pytest_gpu.log 2>&1 ) 2>&1 | grep real
tail -4 gpurun_out/r2z_pytest_gpu.log
show() { python - "$1" "$2" <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], 'value', round(j['value']), 'ms', round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value']), 'e2e ms', round(j['e2e']['ms_per_step'],2), 'nb', j['e2e']['host_batches_per_step'])
PY
}
run() { cfg=$1; tag=$2; shift; shift; env "$@" timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/r2z_${cfg}_$tag.json 2>gpurun_out/r2z_${cfg}_$tag.err; show gpurun_out/r2z_${cfg}_$tag.json "$cfg $tag"; }
for c in C4 C3 C5 C2; do
  run $c pipes2 A=1
  run $c pipes1 AFQ_HOST_PIPES=1
done
timeout 300 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline --no-others --e2e-batches 8 > gpurun_out/r2z_C4_pipes2_nb8.json 2>/dev/null; show gpurun_out/r2z_C4_pipes2_nb8.json "C4 pipes2 nb=8"
