#!/bin/bash
# round 2ad (2 GPUs): the driver's N=2 command as is (default flags: C2 + other_configs under torchrun) and the reference arm under torchrun
mkdir -p gpurun_out
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29921 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2ad_bench_default_n2.json 2> gpurun_out/r2ad_bench_default_n2.err ) 2>&1 | grep real
python - <<'PY' || tail -20 gpurun_out/r2ad_bench_default_n2.err
import json
j=json.loads(open('gpurun_out/r2ad_bench_default_n2.json').read().strip().splitlines()[-1])
def show(k, v): print(k, 'value', round(v['value']), 'ms', round(v['ms_per_step'],2), 'e2e', round(v['e2e']['value']), round(v['e2e']['ms_per_step'],2))
show('C2', j)
for k,v in j.get('other_configs',{}).items(): show(k, v)
PY
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29922 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2ad_bench_reference_n2.json 2> gpurun_out/r2ad_bench_reference_n2.err ) 2>&1 | grep real
tail -c 600 gpurun_out/r2ad_bench_reference_n2.json
