#!/usr/bin/env python
"""Host stages of `bin/alevin-fry quant` without a GPU: chunk index walk + parallel parse of a synthetic collated RAD, and the
parallel text formatting of a synthetic result of the same shape. Prints one JSON line.
usage: host_stage_bench.py [CONFIG=C2] [N_CELLS=20000] [THREADS=all] [REPEATS=3]"""
import json, os, shutil, sys, tempfile, time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from alevin_fry_b200 import host
import synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
n_cells = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
threads = int(sys.argv[3]) if len(sys.argv) > 3 else (os.cpu_count() or 2)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
spec = synth.config_spec(cfg)
tmp = tempfile.mkdtemp(prefix="afq_host_", dir=os.environ.get("AFQ_TMP", "/dev/shm" if os.path.isdir("/dev/shm") else None))
try:
    b = synth.generate(spec, 0, n_cells)
    names = host.write_synth_t2g(os.path.join(tmp, "t2g.tsv"), spec)
    host.write_collated_rad(os.path.join(tmp, "in"), b, host.make_barcodes(0, n_cells), names, 16, spec.umi_len)
    rad = os.path.join(tmp, "in", "map.collated.rad")
    ref = host.rad_summary(rad)
    best = None
    for _ in range(reps):
        s = host.host_stage_bench(rad, threads, 0)
        assert (s.sum_umi, s.sum_refs, s.n_records, s.n_alignments) == (ref.sum_umi, ref.sum_refs, ref.n_records, ref.n_alignments)
        if best is None or s.parse_s + s.format_s < best.parse_s + best.format_s:
            best = s
    print(json.dumps({"what": "host stages of alevin-fry quant, no GPU (RAD in page cache)", "config": cfg, "cells": n_cells, "records": int(best.n_records),
                      "rad_bytes": os.path.getsize(rad), "threads": best.threads, "walk_s": best.walk_s, "parse_s": best.parse_s, "parse_warm_s": best.parse_warm_s,
                      "parse_GBps": os.path.getsize(rad) / best.parse_s / 1e9, "format_s": best.format_s, "nnz": int(best.nnz), "mtx_bytes": int(best.mtx_bytes),
                      "format_Mnnz_s": best.nnz / best.format_s / 1e6, "pack24": bool(best.pack24)}))
finally:
    shutil.rmtree(tmp, ignore_errors=True)
