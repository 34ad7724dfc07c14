#!/bin/bash
# round 2aa (2 GPUs): the NCCL assembly overlapped with the next step's kernels
mkdir -p gpurun_out
P=29811
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 5 --warmup 3 --no-others --no-cpu-baseline > gpurun_out/r2ab_bench_c2_n2.json 2> gpurun_out/r2ab_bench_c2_n2.err ) 2>&1 | grep real
python -c "
import json
j=json.loads(open('gpurun_out/r2ab_bench_c2_n2.json').read().strip().splitlines()[-1]); print('N=2 C2 value',round(j['value']),'ms',round(j['ms_per_step'],2),'e2e',round(j['e2e']['value']),'e2e ms',round(j['e2e']['ms_per_step'],2), j.get('assembly','')[:60])" || tail -20 gpurun_out/r2ab_bench_c2_n2.err
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 2 --config C3 --steps 3 --warmup 3 --no-others --no-cpu-baseline > gpurun_out/r2ab_bench_c3_n2.json 2> gpurun_out/r2ab_bench_c3_n2.err ) 2>&1 | grep real
python -c "
import json
j=json.loads(open('gpurun_out/r2ab_bench_c3_n2.json').read().strip().splitlines()[-1]); print('N=2 C3 value',round(j['value']),'ms',round(j['ms_per_step'],2),'e2e',round(j['e2e']['value']),'e2e ms',round(j['e2e']['ms_per_step'],2))" || tail -20 gpurun_out/r2ab_bench_c3_n2.err
