#!/bin/bash
# round 2r: where the e2e time goes — small-batch device time, e2e vs host batches per step, carry-over across steps
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], 'value', round(j['value']), 'ms', round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value']), 'e2e ms', round(j['e2e']['ms_per_step'],2))
PY
}
for c in C4 C3; do
  timeout 300 python bench.py --config $c --cells 15625 --steps 5 --warmup 3 --no-cpu-baseline --no-others --e2e-batches 1 > gpurun_out/r2r_$c.small.json 2>/dev/null; show gpurun_out/r2r_$c.small.json "$c 15625 cells, 1 batch"
  for nb in 2 4 8 16; do
    timeout 300 python bench.py --config $c --steps 4 --warmup 3 --no-cpu-baseline --no-others --e2e-batches $nb > gpurun_out/r2r_$c.nb$nb.json 2>/dev/null; show gpurun_out/r2r_$c.nb$nb.json "$c carry nb=$nb"
  done
  timeout 300 python bench.py --config $c --steps 4 --warmup 3 --no-cpu-baseline --no-others --e2e-batches 8 --e2e-drain > gpurun_out/r2r_$c.drain8.json 2>/dev/null; show gpurun_out/r2r_$c.drain8.json "$c drain nb=8"
done
for nb in 4 8; do
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-others --e2e-batches $nb > gpurun_out/r2r_C2.nb$nb.json 2>/dev/null; show gpurun_out/r2r_C2.nb$nb.json "C2 carry nb=$nb"
done
