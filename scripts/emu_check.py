import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib, emu_lib
from alevin_fry_b200 import QuantOpts
import synth

def cmp(got, want, tag, exact=True):
    ok = np.array_equal(got.row_ptr, want.row_ptr) and np.array_equal(got.col, want.col)
    if ok: ok = np.array_equal(got.val, want.val) if exact else np.allclose(got.val, want.val, rtol=1e-5, atol=0)
    ok2 = np.array_equal(got.num_expr, want.num_expr) and np.array_equal(got.flags, want.flags) and np.array_equal(got.num_over_mean, want.num_over_mean)
    if exact: ok2 = ok2 and np.array_equal(got.sum_umi, want.sum_umi) and np.array_equal(got.max_umi, want.max_umi)
    bit = np.array_equal(got.val, want.val) and np.array_equal(got.sum_umi, want.sum_umi) if got.val.shape == want.val.shape else False
    print(f"[{'OK' if ok and ok2 else 'FAIL'}] {tag}: nnz emu={got.nnz} cpu={want.nnz} bit_identical={bit}", flush=True)
    if not (ok and ok2):
        bad = np.nonzero(got.num_expr != want.num_expr)[0]
        print("   cells with different num_expr:", bad[:10], got.num_expr[bad[:5]], want.num_expr[bad[:5]])
        for c in range(min(got.n_cells, want.n_cells)):
            gc, gv = got.row(c); wc, wv = want.row(c)
            if not (np.array_equal(gc, wc) and np.array_equal(gv, wv)):
                print("   first differing cell", c, "flags", got.flags[c], want.flags[c])
                sg = dict(zip(gc.tolist(), gv.tolist())); sw = dict(zip(wc.tolist(), wv.tolist()))
                diff = [(k, sg.get(k), sw.get(k)) for k in sorted(set(sg) | set(sw)) if sg.get(k) != sw.get(k)]
                print("   (col, emu, cpu):", diff[:12]); break
    return ok and ok2

res_list = sys.argv[1].split(",") if len(sys.argv) > 1 else ["cr-like"]
cfgs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["C2"]
ncell = int(sys.argv[3]) if len(sys.argv) > 3 else 20
allok = True
for cfgname in cfgs:
    spec = synth.config_spec(cfgname)
    if cfgname == "C2S": spec = synth.SynthSpec(reads_mean=300.0)
    b = synth.generate(spec, 0, ncell); t2g = synth.tid_to_gid(spec)
    for res in res_list:
        for st in (100, 0):
            o = QuantOpts(resolution=res, usa_mode=spec.usa_mode, num_gene_ids=spec.num_gene_ids, num_rows=spec.num_rows, small_thresh=st)
            t0 = time.time(); got = emu_lib.emu_quant(o, t2g, b); dt = time.time() - t0
            allok &= cmp(got, oracle_lib.oracle_quant(o, t2g, b), f"{cfgname}/{res}/st{st} ({dt:.1f}s)", exact=not res.endswith("-em"))
print("ALL OK" if allok else "SOME FAILED")
