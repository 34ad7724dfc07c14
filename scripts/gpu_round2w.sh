#!/bin/bash
# round 2w: ncu --set full of k_pug_build<0..3> as they are now (C3), for the source-line reading; e2e vs host batches at 32 queues
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 900 $NCU --set full --import-source on -k regex:'k_pug_build' --launch-skip 12 -c 4 -f -o gpurun_out/r2w_full_build python bench.py --config C3 --steps 1 --warmup 1 --cells 10000 --no-cpu-baseline --no-others > gpurun_out/r2w_full_build.log 2>&1
ls -la gpurun_out/*.ncu-rep
show() { python - "$1" "$2" <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], 'value', round(j['value']), 'ms', round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value']), 'e2e ms', round(j['e2e']['ms_per_step'],2))
PY
}
for c in C4 C2; do for nb in 2 4 6; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-others --e2e-batches $nb > gpurun_out/r2w_${c}_nb$nb.json 2>/dev/null; show gpurun_out/r2w_${c}_nb$nb.json "$c nb=$nb"
done; done
