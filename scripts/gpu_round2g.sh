#!/bin/bash
# round 2g (2 GPUs): new GPU tests (arena growth, giant split cell), bench at N=2 with the NCCL matrix assembly in the timed step, H2D ceiling at N=1/2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "giant or grows or forced_large" > gpurun_out/r2g_pytest.log 2>&1
tail -3 gpurun_out/r2g_pytest.log
P=29511
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 5 --warmup 3 --no-others > gpurun_out/r2g_bench_c2_n2.json 2> gpurun_out/r2g_bench_c2_n2.err
python -c "
import json
j=json.loads(open('gpurun_out/r2g_bench_c2_n2.json').read().strip().splitlines()[-1]); print('N=2 C2 value',round(j['value']),'ms',round(j['ms_per_step'],2),'e2e',round(j['e2e']['value']))"
timeout 300 python scripts/h2d_ceiling.py > gpurun_out/r2g_h2d_n1.json 2>/dev/null; cat gpurun_out/r2g_h2d_n1.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) scripts/h2d_ceiling.py > gpurun_out/r2g_h2d_n2.json 2>/dev/null; cat gpurun_out/r2g_h2d_n2.json
nvidia-smi topo -m > gpurun_out/r2g_topo.txt 2>&1; nproc
