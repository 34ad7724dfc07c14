"""A fixed list of cells run through the emulated kernels and compared with the oracle; prints one line per
case. Used by scripts/emu_asan.sh (sanitizers) and tests/test_emu_schedules.py (alternative fiber schedules,
AFQ_EMU_SCHED). `quick` as first argument runs a shorter list. Test infrastructure."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
QUICK = len(sys.argv) > 1 and sys.argv[1] == 'quick'
import numpy as np, cases, emu_lib, oracle_lib
from alevin_fry_b200 import CellBatch, QuantOpts
import synth
import test_emu_parity as T
def run(o,t2g,b,tag):
    got=emu_lib.emu_quant(o,t2g,b); want=oracle_lib.oracle_quant(o,t2g,b,n_threads=2)
    ok=all(np.array_equal(getattr(got,f),getattr(want,f)) for f in ('row_ptr','col','val','num_expr','flags'))
    print(tag,'OK' if ok else 'MISMATCH', flush=True)
spec=synth.SynthSpec(reads_mean=400.0); b=synth.generate(spec,0,4 if QUICK else 8); t2g=synth.tid_to_gid(spec)
for res in ['parsimony','parsimony-em','parsimony-gene','cr-like-em','cr-like']:
    run(T.opts_for(spec,res),t2g,b,'c2/'+res)
for regime in (T.RANDOM_REGIMES[1:3] if QUICK else T.RANDOM_REGIMES[:4]):
    seed, n_tx, per_gene, umi_bits, n_labels, max_label, lo, hi = regime
    rng=np.random.default_rng(seed); n_genes=(n_tx+per_gene-1)//per_gene
    t2=(np.arange(n_tx,dtype=np.uint32)//per_gene).astype(np.uint32)
    bb=CellBatch.from_cells(cases.random_pug_cells(rng,4,n_tx,umi_bits,n_labels,max_label,lo,hi))
    for res in ['parsimony','parsimony-em']:
        run(QuantOpts(resolution=res,num_gene_ids=n_genes,num_rows=n_genes,umi_len=(umi_bits+1)//2,small_thresh=0),t2,bb,f'rand{seed}/{res}')
if not QUICK:
    spec=synth.SynthSpec(fixed_reads=10000,n_genes=3000); b=synth.generate(spec,0,1)
    run(T.opts_for(spec,'parsimony'),synth.tid_to_gid(spec),b,'global-arena')
