#!/usr/bin/env python
"""Regenerates tests/golden/*.npz: small seeded inputs (the synthetic generator of SURVEY.md §8(d)) together
with the ORACLE's outputs for every resolution. The reference itself (Rust) cannot be built or run in
this environment, so these vectors pin the oracle (and through it the CUDA path) against regressions;
the oracle in turn is pinned by the reference's own known-answer tests (tests/test_oracle_kat.py).

usage: python tests/golden/make_golden.py      (from the repo root, after `make`)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib  # noqa: E402
from alevin_fry_b200 import QuantOpts  # noqa: E402
import synth  # noqa: E402

ALL_RES = ["trivial", "cr-like", "cr-like-em", "parsimony", "parsimony-em", "parsimony-gene", "parsimony-gene-em"]
CASES = {
    # name: (spec, first cell, cells, resolutions, extra QuantOpts)
    "c1_tiny_cells": (synth.config_spec("C1"), 0, 64, ["cr-like", "parsimony", "cr-like-em"], {}),
    "c2_mini": (synth.SynthSpec(reads_mean=400.0), 0, 12, ALL_RES, {}),
    "c4_mini_usa": (synth.SynthSpec(usa_mode=True, reads_mean=400.0, n_genes=2000), 0, 10,
                    ["cr-like", "cr-like-em", "parsimony", "parsimony-em"], {}),
    "c5_mini_high_dup": (synth.SynthSpec(n_genes=5000, reads_mean=600.0, reads_per_umi=40.0, p_multi2=0.15, p_multi3=0.15, umi_err=0.02),
                         0, 10, ["parsimony-em", "parsimony", "cr-like"], {}),
    "dense_umi_components": (synth.SynthSpec(n_genes=300, umi_len=6, reads_mean=600.0, reads_per_umi=1.5, umi_err=0.05), 0, 6,
                             ["parsimony", "parsimony-em", "parsimony-gene"], {"small_thresh": 0}),
}


def main():
    for name, (spec, c0, nc, resolutions, extra) in CASES.items():
        b = synth.generate(spec, c0, nc)
        t2g = synth.tid_to_gid(spec)
        out = {"cell_rec_offsets": b.cell_rec_offsets, "rec_umi32": b.rec_umi32, "rec_ref_offsets": b.rec_ref_offsets,
               "refs": b.refs, "tid_to_gid": t2g,
               "meta": np.array([int(spec.usa_mode), spec.num_gene_ids, spec.num_rows, spec.umi_len, extra.get("small_thresh", 100)], dtype=np.int64),
               "resolutions": np.array(resolutions)}
        for res in resolutions:
            o = QuantOpts(resolution=res, usa_mode=spec.usa_mode, num_gene_ids=spec.num_gene_ids, num_rows=spec.num_rows,
                          umi_len=spec.umi_len, **extra)
            r = oracle_lib.oracle_quant(o, t2g, b, n_threads=2)
            for f in ("row_ptr", "col", "val", "sum_umi", "max_umi", "num_expr", "num_over_mean", "flags"):
                out[f"{res}/{f}"] = getattr(r, f)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, b.n_cells, "cells", b.n_records, "records ->", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
