import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import synth
from alevin_fry_b200 import QuantOpts, Quantifier
spec = synth.config_spec("C3"); t2g = synth.tid_to_gid(spec)
b = synth.generate(spec, 0, int(sys.argv[1]) if len(sys.argv) > 1 else 20000)
o = QuantOpts(resolution="parsimony", num_gene_ids=spec.num_gene_ids, num_rows=spec.num_rows)
with Quantifier(o, t2g) as q:
    q.quantify_batch(b)
    q.set_profiling(True)
    r = q.quantify_batch(b)
    print({k: (round(v[0], 2), v[1]) for k, v in q.profile().items()})
