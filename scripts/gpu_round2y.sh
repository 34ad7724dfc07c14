#!/bin/bash
# round 2y (8 GPUs): the box's H2D ceiling at 1/2/4/8 concurrent processes, bench C2 at N=8 and N=4 (NCCL matrix assembly in the timed step)
mkdir -p gpurun_out
P=29611
nvidia-smi topo -m > gpurun_out/r2y_topo.txt 2>&1; nproc > gpurun_out/r2y_nproc.txt; free -g | head -2 >> gpurun_out/r2y_nproc.txt
timeout 200 python scripts/h2d_ceiling.py > gpurun_out/r2y_h2d_n1.json 2>/dev/null; cat gpurun_out/r2y_h2d_n1.json
for n in 2 4 8; do
  P=$((P+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P scripts/h2d_ceiling.py > gpurun_out/r2y_h2d_n$n.json 2>/dev/null; cat gpurun_out/r2y_h2d_n$n.json
done
for n in 8 4; do
  P=$((P+1))
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --steps 5 --warmup 3 --no-others --no-cpu-baseline > gpurun_out/r2y_bench_c2_n$n.json 2> gpurun_out/r2y_bench_c2_n$n.err ) 2>&1 | grep real
  python -c "
import json
j=json.loads(open('gpurun_out/r2y_bench_c2_n$n.json').read().strip().splitlines()[-1]); print('N=$n C2 value',round(j['value']),'ms',round(j['ms_per_step'],2),'e2e',round(j['e2e']['value']),'e2e ms',round(j['e2e']['ms_per_step'],2), j.get('clocks'))" || tail -5 gpurun_out/r2y_bench_c2_n$n.err
done
