#!/bin/bash
# round 2af: is the CLI's 2.3 s set-up (CUDA context creation) caused by the 32 hardware work queues?
mkdir -p gpurun_out
nvidia-smi -L | head -3; nproc
for q in 8 32; do
  CUDA_DEVICE_MAX_CONNECTIONS=$q timeout 300 python scripts/cli_bench.py C2 3000 cr-like 2 > gpurun_out/r2af_cli_q$q.json 2>/dev/null
  python -c "
import json; j=json.loads(open('gpurun_out/r2af_cli_q$q.json').read().strip().splitlines()[-1]); print('queues $q', [round(x,2) for x in j.get('wall_s')]); print(j.get('timing_log'))"
done
python - <<'PY'
import ctypes, time, os
t0=time.time(); l=ctypes.CDLL('/usr/local/cuda/lib64/libcudart.so'); p=ctypes.c_void_p(); r=l.cudaFree(0); print('cudaFree(0) [context creation], default queues:', round(time.time()-t0,3), 's rc', r)
PY
CUDA_DEVICE_MAX_CONNECTIONS=32 python - <<'PY'
import ctypes, time, os
t0=time.time(); l=ctypes.CDLL('/usr/local/cuda/lib64/libcudart.so'); r=l.cudaFree(0); print('cudaFree(0) [context creation], 32 queues:', round(time.time()-t0,3), 's rc', r)
PY
