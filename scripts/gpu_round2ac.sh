#!/bin/bash
# round 2ac (8 GPUs): bench C2 at N=8 with the NCCL assembly overlapped with the next step's kernels (r2y without: 12.89 ms/step)
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 8 --steps 5 --warmup 3 --no-others --no-cpu-baseline > gpurun_out/r2ac_bench_c2_n8.json 2> gpurun_out/r2ac_bench_c2_n8.err ) 2>&1 | grep real
python -c "
import json
j=json.loads(open('gpurun_out/r2ac_bench_c2_n8.json').read().strip().splitlines()[-1]); print('N=8 C2 value',round(j['value']),'ms',round(j['ms_per_step'],2),'e2e',round(j['e2e']['value']),'e2e ms',round(j['e2e']['ms_per_step'],2), round(j['roofline']['per_kernel_ms']['resolve_region(wall)'],2))" || tail -20 gpurun_out/r2ac_bench_c2_n8.err
