#!/bin/bash
# round 2t: why C4's e2e got slower without the read-back: hardware queue aliasing between the copy stream and the lanes?
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], 'value', round(j['value']), 'ms', round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value']), 'e2e ms', round(j['e2e']['ms_per_step'],2))
PY
}
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/r2t_$tag.json 2>gpurun_out/r2t_$tag.err; show gpurun_out/r2t_$tag.json "C4 $tag"; }
run base A=1
run conn32 CUDA_DEVICE_MAX_CONNECTIONS=32
run nolanes AFQ_NO_LANES=1
run conn32_sync CUDA_DEVICE_MAX_CONNECTIONS=32 AFQ_SYNC_SIZING=1
run pool4g AFQ_POOL_MB=4096
