"""`alevin-fry infer` (src/infer.rs): EM over a global gene-eq-class table. The kernel (k_em_subset) under emulation and on
the GPU through afq_infer vs the oracle's em_optimize_subset; plus the reference's own known-answer tests for that function."""
import numpy as np
import pytest

import oracle_lib
from alevin_fry_b200 import QuantOpts, Quantifier


def random_table(rng, n_genes, n_classes, usa):
    """a global class table: single-gene and multi-gene labels (ascending, distinct), in the S | U | A column space for USA"""
    width = 3 * n_genes if usa else n_genes
    labels, off = [], [0]
    for _ in range(n_classes):
        k = 1 if rng.random() < 0.45 else int(rng.integers(2, 7))
        if usa:
            g = rng.choice(n_genes, size=min(k, n_genes), replace=False)
            lab = sorted(set(int(x) + n_genes * int(rng.integers(0, 3)) for x in g))
        else:
            lab = sorted(int(x) for x in rng.choice(width, size=min(k, width), replace=False))
        labels.extend(lab)
        off.append(len(labels))
    return np.array(off, dtype=np.uint32), np.array(labels, dtype=np.uint32), width


def random_cells(rng, n_cells, n_classes, lo, hi):
    off, eq, cnt = [0], [], []
    for _ in range(n_cells):
        k = int(rng.integers(lo, hi))
        ids = np.sort(rng.choice(n_classes, size=min(k, n_classes), replace=False))   # CSR row: ascending class ids
        eq.extend(int(x) for x in ids)
        cnt.extend(int(x) for x in (1 + rng.geometric(0.3, size=len(ids))))
        off.append(len(eq))
    return np.array(off, dtype=np.uint64), np.array(eq, dtype=np.uint32), np.array(cnt, dtype=np.uint32)


def same(got, want, exact):
    assert np.array_equal(got.row_ptr, want.row_ptr)
    assert np.array_equal(got.col, want.col)
    if exact:
        assert np.array_equal(got.val, want.val)
    else:
        np.testing.assert_allclose(got.val, want.val, rtol=1e-5, atol=0)
    assert np.array_equal(got.num_expr, want.num_expr)


CASES = [(False, False), (False, True), (True, False), (True, True)]


@pytest.mark.parametrize("usa,uniform", CASES)
def test_emu_infer_matches_the_oracle(usa, uniform):
    import emu_lib
    rng = np.random.default_rng(11 + 2 * usa + uniform)
    lo, lb, width = random_table(rng, 60, 400, usa)
    co, ce, cc = random_cells(rng, 14, 400, 0, 90)
    want = oracle_lib.infer_cells(width, usa, uniform, lo, lb, co, ce, cc)
    same(emu_lib.emu_infer(width, usa, uniform, lo, lb, co, ce, cc), want, exact=True)
    same(emu_lib.emu_infer(width, usa, uniform, lo, lb, co, ce, cc, force_global=True), want, exact=True)   # per-CTA global arena


def test_emu_infer_reference_known_answers():
    # src/em.rs:1175-1215: classes {0},{1},{0,1},{1,2},{2,3,4} over 8 genes; cell data [], [(0,7)], [(0,20),(1,4),(2,8),(3,1),(4,2)];
    # the clamp case [(0,10000),(1,1),(2,1),(3,1)] over {0},{0,1},{1,2},{2,3,4} must contain an exact 0.0
    import emu_lib
    lo = np.array([0, 1, 2, 4, 6, 9], dtype=np.uint32)
    lb = np.array([0, 1, 0, 1, 1, 2, 2, 3, 4], dtype=np.uint32)
    co = np.array([0, 0, 1, 6], dtype=np.uint64)
    ce = np.array([0, 0, 1, 2, 3, 4], dtype=np.uint32)
    cc = np.array([7, 20, 4, 8, 1, 2], dtype=np.uint32)
    for uniform in (False, True):
        got = emu_lib.emu_infer(8, False, uniform, lo, lb, co, ce, cc)
        same(got, oracle_lib.infer_cells(8, False, uniform, lo, lb, co, ce, cc), exact=True)
        assert got.num_expr[0] == 0 and got.flags[0] == 4
        assert got.row(1)[0].tolist() == [0] and got.row(1)[1].tolist() == [7.0]          # no ambiguity: the tallies themselves
        assert abs(float(got.row(2)[1].sum()) - 35.0) < 1e-2                               # mass is conserved
    lo2 = np.array([0, 1, 3, 5, 8], dtype=np.uint32)
    lb2 = np.array([0, 0, 1, 1, 2, 2, 3, 4], dtype=np.uint32)
    got = emu_lib.emu_infer(8, False, False, lo2, lb2, np.array([0, 4], dtype=np.uint64), np.array([0, 1, 2, 3], dtype=np.uint32),
                            np.array([10000, 1, 1, 1], dtype=np.uint32))
    want = oracle_lib.infer_cells(8, False, False, lo2, lb2, [0, 4], np.array([0, 1, 2, 3], dtype=np.uint32), np.array([10000, 1, 1, 1], dtype=np.uint32))
    same(got, want, exact=True)
    assert got.num_expr[0] < 5        # some gene was clamped to an exact 0.0 and left the row


@pytest.mark.gpu
@pytest.mark.parametrize("usa,uniform", CASES)
def test_gpu_infer_matches_the_oracle(usa, uniform):
    rng = np.random.default_rng(5 + 2 * usa + uniform)
    lo, lb, width = random_table(rng, 3000, 20000, usa)
    co, ce, cc = random_cells(rng, 700, 20000, 0, 900)
    # one very large cell: beyond the shared-memory arena, runs on the global arena
    big = np.sort(rng.choice(20000, size=9000, replace=False)).astype(np.uint32)
    co = np.concatenate([co, [co[-1] + len(big)]]).astype(np.uint64)
    ce = np.concatenate([ce, big]); cc = np.concatenate([cc, np.ones(len(big), dtype=np.uint32)])
    want = oracle_lib.infer_cells(width, usa, uniform, lo, lb, co, ce, cc)
    o = QuantOpts(resolution="cr-like-em", usa_mode=usa, num_gene_ids=(2 * width // 3) if usa else width, num_rows=width, init_uniform=uniform)
    with Quantifier(o, np.zeros(4, dtype=np.uint32)) as q:
        got = q.infer(lo, lb, co, ce, cc)
        same(got, want, exact=False)
        assert np.array_equal(got.val, want.val)          # in practice bit-identical (same operation order and roundings)
        with pytest.raises(Exception):
            q.infer(lo, lb, co, ce + 50000, cc)           # class id out of range


# ---- --dump-eqclasses: the per-cell gene eq-classes (src/quant.rs:1282-1307) ------------------------------------------
DUMP_RES = ["cr-like", "cr-like-em", "parsimony", "parsimony-em", "parsimony-gene", "parsimony-gene-em"]


def same_dump(got, want):
    assert np.array_equal(got.cell_cls_ptr, want.cell_cls_ptr)
    assert np.array_equal(got.cls_lab_ptr, want.cls_lab_ptr)
    assert np.array_equal(got.labels, want.labels)
    assert np.array_equal(got.counts, want.counts)


@pytest.mark.parametrize("res", DUMP_RES)
@pytest.mark.parametrize("usa", [False, True])
def test_emu_dump_eqclasses_matches_the_oracle(res, usa):
    import emu_lib
    import synth
    spec = synth.SynthSpec(reads_mean=350.0, usa_mode=usa, n_genes=1500)
    b = synth.generate(spec, 0, 10)
    t2g = synth.tid_to_gid(spec)
    o = QuantOpts(resolution=res, usa_mode=usa, num_gene_ids=spec.num_gene_ids, num_rows=spec.num_rows, umi_len=spec.umi_len, dump_eq=True)
    got, gd = emu_lib.emu_quant_with_classes(o, t2g, b)
    want, wd = oracle_lib.oracle_quant_with_classes(o, t2g, b, n_threads=2)
    same(got, want, exact=True)
    same_dump(gd, wd)
    assert wd.cell_cls_ptr[-1] > 0
    # the counts are the same with or without the dump
    o2 = QuantOpts(resolution=res, usa_mode=usa, num_gene_ids=spec.num_gene_ids, num_rows=spec.num_rows, umi_len=spec.umi_len)
    same(emu_lib.emu_quant(o2, t2g, b), want, exact=True)


@pytest.mark.gpu
@pytest.mark.parametrize("res", DUMP_RES)
def test_gpu_dump_eqclasses_matches_the_oracle(res):
    import synth
    for usa in (False, True):
        spec = synth.config_spec("C4" if usa else "C2")
        b = synth.generate(spec, 0, 300)
        t2g = synth.tid_to_gid(spec)
        o = QuantOpts(resolution=res, usa_mode=usa, num_gene_ids=spec.num_gene_ids, num_rows=spec.num_rows, dump_eq=True)
        with Quantifier(o, t2g) as q:
            got, gd = q.quantify_batch_with_classes(b, use_na8=True, use_pack24=True)
            got2, gd2 = q.quantify_batch_with_classes(b.slice_cells(10, 200))
        want, wd = oracle_lib.oracle_quant_with_classes(o, t2g, b)
        same(got, want, exact=not res.endswith("-em"))
        same_dump(gd, wd)
        assert gd2.cell(0) == wd.cell(10) and gd2.cell(189) == wd.cell(199)
