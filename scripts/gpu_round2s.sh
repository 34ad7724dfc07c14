#!/bin/bash
# round 2s: afq_submit without the control-block read-back (arenas planned on the device): GPU tier, e2e of C2..C5
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2s_pytest_gpu.log 2>&1 ) 2>&1 | grep real
tail -4 gpurun_out/r2s_pytest_gpu.log
show() { python - "$1" "$2" <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], 'value', round(j['value']), 'ms', round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value']), 'e2e ms', round(j['e2e']['ms_per_step'],2))
PY
}
for c in C2 C3 C4 C5; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/r2s_$c.json 2>gpurun_out/r2s_$c.err; show gpurun_out/r2s_$c.json "$c async nb=8"
done
for c in C3 C4; do
  AFQ_SYNC_SIZING=1 timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/r2s_$c.sync.json 2>/dev/null; show gpurun_out/r2s_$c.sync.json "$c sync nb=8"
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-others --e2e-batches 4 > gpurun_out/r2s_$c.nb4.json 2>/dev/null; show gpurun_out/r2s_$c.nb4.json "$c async nb=4"
done
