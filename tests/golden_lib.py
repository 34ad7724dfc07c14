"""Loader for the committed golden vectors under tests/golden/ (made by tests/golden/make_golden.py)."""
import glob
import os

import numpy as np

from alevin_fry_b200 import CellBatch, QuantOpts

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = ("row_ptr", "col", "val", "sum_umi", "max_umi", "num_expr", "num_over_mean", "flags")


def cases():
    out = []
    for path in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        z = np.load(path)
        for res in z["resolutions"]:
            out.append((os.path.basename(path)[:-4], str(res)))
    return out


def load(name, res):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    usa, gene_ids, rows, umi_len, small_thresh = (int(x) for x in z["meta"])
    batch = CellBatch(z["cell_rec_offsets"], z["rec_umi32"], z["rec_ref_offsets"], z["refs"], 0)
    opts = QuantOpts(resolution=res, usa_mode=bool(usa), num_gene_ids=gene_ids, num_rows=rows, umi_len=umi_len,
                     small_thresh=small_thresh)
    want = {f: z[f"{res}/{f}"] for f in FIELDS}
    return opts, z["tid_to_gid"], batch, want


def assert_matches(got, want, exact, ctx):
    for f in FIELDS:
        g = getattr(got, f)
        if exact or f not in ("val", "sum_umi", "max_umi"):
            if f == "num_over_mean" and not exact:
                continue
            assert np.array_equal(g, want[f]), f"{ctx}: {f} differs from the golden vector"
        else:
            np.testing.assert_allclose(g, want[f], rtol=1e-5 if f != "sum_umi" else 1e-4, atol=0, err_msg=f"{ctx}: {f}")
