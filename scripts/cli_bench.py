#!/usr/bin/env python
"""End-to-end CLI throughput: synthetic collated RAD on disk (page cache) -> `bin/alevin-fry quant`
-> alevin/quants_mat.{mtx,rows,cols} + featureDump + quant.json. Prints one JSON line.
usage: cli_bench.py [CONFIG=C2] [N_CELLS=20000] [RESOLUTION=cr-like] [REPEATS=3]"""
import json, os, shutil, subprocess, sys, tempfile, time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from alevin_fry_b200 import host
import synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
n_cells = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
res = sys.argv[3] if len(sys.argv) > 3 else "cr-like"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
spec = synth.config_spec(cfg)
tmp = tempfile.mkdtemp(prefix="afq_cli_", dir=os.environ.get("AFQ_TMP", "/dev/shm" if os.path.isdir("/dev/shm") else None))
try:
    t0 = time.time()
    b = synth.generate(spec, 0, n_cells)
    names = host.write_synth_t2g(os.path.join(tmp, "t2g.tsv"), spec)
    host.write_collated_rad(os.path.join(tmp, "in"), b, host.make_barcodes(0, n_cells), names, 16, spec.umi_len)
    rad_bytes = os.path.getsize(os.path.join(tmp, "in", "map.collated.rad"))
    gen_s = time.time() - t0
    times, logs = [], []
    for r in range(reps):
        out = os.path.join(tmp, "out")
        shutil.rmtree(out, ignore_errors=True)
        t0 = time.time()
        p = subprocess.run([host.CLI_PATH, "quant", "-i", os.path.join(tmp, "in"), "-m", os.path.join(tmp, "t2g.tsv"), "-o", out, "-r", res],
                           capture_output=True, text=True, env=dict(os.environ, AFQ_TIMING="1"))
        dt = time.time() - t0
        if p.returncode != 0:
            print(json.dumps({"error": p.stderr[-2000:]})); sys.exit(1)
        times.append(dt); logs.append(p.stderr.strip().splitlines()[-12:])
    mtx_bytes = os.path.getsize(os.path.join(tmp, "out", "alevin", "quants_mat.mtx"))
    best = min(times)
    print(json.dumps({"what": "alevin-fry quant CLI end to end (RAD in page cache -> output files)", "config": cfg, "resolution": res,
                      "cells": n_cells, "records": int(b.n_records), "rad_bytes": rad_bytes, "mtx_bytes": mtx_bytes,
                      "wall_s": times, "best_s": best, "cells_per_s": n_cells / best, "records_per_s": b.n_records / best,
                      "host_threads": os.cpu_count(), "timing_log": logs[times.index(best)], "gen_s": gen_s}))
finally:
    shutil.rmtree(tmp, ignore_errors=True)
