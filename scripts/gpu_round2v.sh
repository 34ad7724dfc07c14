#!/bin/bash
# round 2v: 32 hardware queues by default + lane priorities (biggest arenas first)
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
pk=j['roofline']['per_kernel_ms']
print(sys.argv[2], 'value', round(j['value']), 'ms', round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value']), 'e2e ms', round(j['e2e']['ms_per_step'],2), {k:round(v,2) for k,v in pk.items() if 'region' in k})
PY
}
run() { cfg=$1; tag=$2; shift; shift; env "$@" timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/r2v_${cfg}_$tag.json 2>gpurun_out/r2v_${cfg}_$tag.err; show gpurun_out/r2v_${cfg}_$tag.json "$cfg $tag"; }
for c in C4 C2 C3 C5; do
  run $c prio A=1
  run $c noprio AFQ_LANE_PRIO=0
done
