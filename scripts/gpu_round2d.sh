#!/bin/bash
# round 2d: why does k_gene_eqc take 15 ms behind the split path? (work-list sizes), and the small build arena at 4 CTAs/SM
mkdir -p gpurun_out
AFQ_DEBUG_CTL=1 python scripts/gpu_dbg.py 60000 2>&1 | tail -4
AFQ_DEBUG_CTL=1 AFQ_NO_PS_SPLIT=1 python scripts/gpu_dbg.py 60000 2>&1 | tail -4
bash scripts/gpu_ab.sh r2d "C3" a e
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pug" > gpurun_out/r2d_pytest_pug.log 2>&1
tail -3 gpurun_out/r2d_pytest_pug.log
