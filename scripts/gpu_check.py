"""Quick GPU sanity: parity of the implemented resolutions vs the oracle + rough timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import oracle_lib
from alevin_fry_b200 import QuantOpts, Quantifier
import synth

def cmp(got, want, tag, exact=True):
    ok = np.array_equal(got.row_ptr, want.row_ptr) and np.array_equal(got.col, want.col)
    if ok:
        ok = np.array_equal(got.val, want.val) if exact else np.allclose(got.val, want.val, rtol=1e-5, atol=0)
    ok2 = np.array_equal(got.num_expr, want.num_expr) and np.array_equal(got.flags, want.flags) and (not exact or (np.array_equal(got.sum_umi, want.sum_umi) and np.array_equal(got.max_umi, want.max_umi) and np.array_equal(got.num_over_mean, want.num_over_mean)))
    print(f"[{'OK' if ok and ok2 else 'FAIL'}] {tag}: nnz gpu={got.nnz} cpu={want.nnz}", flush=True)
    if not (ok and ok2):
        bad = np.nonzero(got.num_expr != want.num_expr)[0]
        print("   first bad cells (num_expr):", bad[:10], got.num_expr[bad[:5]], want.num_expr[bad[:5]])
        if len(bad) == 0 and got.nnz == want.nnz:
            d = np.nonzero((got.col != want.col) | (got.val != want.val))[0]
            print("   first diffs at nnz idx", d[:10], got.col[d[:5]], want.col[d[:5]], got.val[d[:5]], want.val[d[:5]])
    return ok and ok2

resolutions = sys.argv[1].split(",") if len(sys.argv) > 1 else ["cr-like", "trivial"]
allok = True
for cfgname, ncell in (("C2", 400), ("C1", 1000), ("C4", 200), ("C5", 200)):
    spec = synth.config_spec(cfgname)
    b = synth.generate(spec, 0, ncell)
    t2g = synth.tid_to_gid(spec)
    for res in resolutions:
        if res == "trivial" and spec.usa_mode: continue
        for st in (100, 0):
            o = QuantOpts(resolution=res, usa_mode=spec.usa_mode, num_gene_ids=spec.num_gene_ids, num_rows=spec.num_rows, small_thresh=st)
            with Quantifier(o, t2g) as q:
                got = q.quantify_batch(b)
            want = oracle_lib.oracle_quant(o, t2g, b)
            allok &= cmp(got, want, f"{cfgname}/{res}/st{st}", exact=not res.endswith("-em"))
spec = synth.SynthSpec(reads_mean=3000.0, lognorm_sigma=1.6, reads_per_umi=1.3)
b = synth.generate(spec, 0, 400); t2g = synth.tid_to_gid(spec)
for res in resolutions:
    o = QuantOpts(resolution=res, num_gene_ids=spec.num_gene_ids, num_rows=spec.num_rows)
    with Quantifier(o, t2g) as q:
        got = q.quantify_batch(b)
    allok &= cmp(got, oracle_lib.oracle_quant(o, t2g, b), f"skew/{res}", exact=not res.endswith("-em"))
print("ALL OK" if allok else "SOME FAILED", flush=True)

# rough device-resident timing on a C2 slice
spec = synth.config_spec("C2")
nc = int(os.environ.get("NC", 20000))
t0 = time.time(); b = synth.generate(spec, 0, nc); print(f"gen {nc} cells {b.n_records} recs in {time.time()-t0:.1f}s", flush=True)
t2g = synth.tid_to_gid(spec)
dev = torch.device("cuda:0")
def T(a, dt): return torch.from_numpy(a.view(dt)).to(dev)
db = dict(cell_rec_offsets=T(b.cell_rec_offsets, np.int64), rec_umi32=T(b.rec_umi32, np.int32), rec_ref_offsets=T(b.rec_ref_offsets, np.int32), refs=T(b.refs, np.int32))
do = dict(row_ptr=torch.empty(nc + 1, dtype=torch.int64, device=dev), col=torch.empty(b.n_refs_total, dtype=torch.int32, device=dev), val=torch.empty(b.n_refs_total, dtype=torch.float32, device=dev),
          sum_umi=torch.empty(nc, dtype=torch.float32, device=dev), max_umi=torch.empty(nc, dtype=torch.float32, device=dev), num_expr=torch.empty(nc, dtype=torch.int32, device=dev),
          num_over_mean=torch.empty(nc, dtype=torch.int32, device=dev), flags=torch.empty(nc, dtype=torch.uint8, device=dev))
for res in resolutions:
    o = QuantOpts(resolution=res, num_gene_ids=spec.num_gene_ids, num_rows=spec.num_rows)
    with Quantifier(o, t2g) as q:
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3): q.quant_device(db, do, st)
        q.device_finish(st)
        q.set_profiling(True); q.profile_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): q.quant_device(db, do, st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        nnz = q.device_finish(st, do["row_ptr"])
        inb = 8 * b.n_records + 4 * b.n_refs_total
        print(f"{res}: {ms:.3f} ms/batch  {nc/ms*1e3:.3e} cells/s  {b.n_records/ms*1e3:.3e} rec/s  in-bytes {inb/1e6:.1f}MB -> {inb/ms/1e6:.1f} GB/s  nnz={nnz}")
        for k, (m, n) in q.profile().items(): print(f"    {k:24s} {m/5:.3f} ms/batch  ({n} launches)")
        t0 = time.time(); r = q.quantify_batch(b); dt = time.time() - t0
        print(f"    host API e2e: {dt*1e3:.1f} ms -> {nc/dt:.3e} cells/s")
