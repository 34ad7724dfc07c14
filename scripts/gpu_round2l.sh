#!/bin/bash
# round 2l: stage C of the EM resolutions in its own kernel (k_em_cells) vs ge_back's stage C inside k_pug_back (AFQ_NO_EM_SPLIT=1)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "not full_size" > gpurun_out/r2l_pytest.log 2>&1
tail -3 gpurun_out/r2l_pytest.log
run() { env $2 timeout 900 python bench.py --config $1 $3 --steps 3 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/r2l_$1$4.json 2>/dev/null
python -c "
import json
j=json.loads(open('gpurun_out/r2l_$1$4.json').read().strip().splitlines()[-1]); pk=j['roofline']['per_kernel_ms']
print('$1 $3 $2 ms', round(j['ms_per_step'],2), {k:round(v,2) for k,v in pk.items() if v>0.3 and ('region' in k or 'gene_eqc' in k or 'bin' in k)})"; }
run C4 "AFQ_X=0" "" ""
run C4 "AFQ_NO_EM_SPLIT=1" "" "_noemsplit"
run C5 "AFQ_X=0" "" ""
run C5 "AFQ_NO_EM_SPLIT=1" "" "_noemsplit"
run C3 "AFQ_X=0" "--resolution parsimony-em" "_em"
run C3 "AFQ_NO_EM_SPLIT=1" "--resolution parsimony-em" "_em_noemsplit"
run C2 "AFQ_X=0" "--resolution cr-like-em" "_em"
run C2 "AFQ_NO_PS_SPLIT=1" "--resolution cr-like-em" "_em_nosplit"
