#!/bin/bash
# round 2u: CUDA_DEVICE_MAX_CONNECTIONS 8 (default) vs 32 on every configuration, value and e2e
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], 'value', round(j['value']), 'ms', round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value']), 'e2e ms', round(j['e2e']['ms_per_step'],2))
PY
}
run() { cfg=$1; tag=$2; shift; shift; env "$@" timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/r2u_${cfg}_$tag.json 2>gpurun_out/r2u_${cfg}_$tag.err; show gpurun_out/r2u_${cfg}_$tag.json "$cfg $tag"; }
for c in C2 C3 C4 C5; do
  run $c conn8 A=1
  run $c conn32 CUDA_DEVICE_MAX_CONNECTIONS=32
  run $c conn16 CUDA_DEVICE_MAX_CONNECTIONS=16
done
run C4 conn8b A=1
run C4 conn32b CUDA_DEVICE_MAX_CONNECTIONS=32
