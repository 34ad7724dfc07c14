#!/bin/bash
# round 2j: arena tiers of k_pug_back (AFQ_BACK_MAX_TIER = 0: 48 KB shared memory or global arenas; 1: + 100 KB; 2: + 224 KB) on the EM configurations
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pug or tiny or mini" > gpurun_out/r2j_pytest.log 2>&1
tail -3 gpurun_out/r2j_pytest.log
run() { AFQ_BACK_MAX_TIER=$2 timeout 900 python bench.py --config $1 $3 --steps 3 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/r2j_$1_t$2$4.json 2>/dev/null
python -c "
import json
j=json.loads(open('gpurun_out/r2j_$1_t$2$4.json').read().strip().splitlines()[-1]); pk=j['roofline']['per_kernel_ms']
print('$1 $3 tier<=$2 ms', round(j['ms_per_step'],2), {k:round(v,2) for k,v in pk.items() if v>0.3 and ('back' in k or 'region' in k or 'gene_eqc' in k)})"; }
for t in 0 1 2; do run C4 $t "" ""; done
for t in 0 1; do run C5 $t "" ""; done
for t in 0 1 2; do run C3 $t "--resolution parsimony-em" "_em"; done
