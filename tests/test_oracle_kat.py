"""Known-answer tests pinning the CPU oracle against the vectors the reference's own test
suite holds for the quant hot path:

  * src/em.rs:1175-1215  (sparse-support EM == dense EM bit-for-bit; threshold clamp)
  * src/em.rs:1217-1241  (support set; zero mass outside it)
  * tests/multi_barcode_integration.rs:1404-1556  (tiny-cell path vs requested resolution)
  * tests/multi_barcode_integration.rs:721-863   (`-r trivial`, one ref per read)
  * src/utils.rs:389-393 (2-bit Hamming distance)
plus hand-derived cases for cr-like / USA / PUG semantics.
"""
import numpy as np
import pytest

import oracle_lib
from alevin_fry_b200 import CellBatch, QuantOpts, FLAG_TINY, FLAG_EMPTY, FLAG_ALT
import synth

F = np.float32


def dense_reference(classes, cell_data, init_uniform, num_alphas, usa_offsets=None):
    """Independent restatement of the in-test `dense_reference` of src/em.rs:1050-1133,
    in numpy float32 scalar arithmetic (one rounding per operation, like Rust f32)."""
    a_in = np.zeros(num_alphas, dtype=F)
    a_out = np.zeros(num_alphas, dtype=F)
    needs_em = False
    for eq, ct in cell_data:
        lab = classes[eq]
        if len(lab) == 1:
            a_in[lab[0]] = F(a_in[lab[0]] + F(ct))
        else:
            needs_em = True
    if not needs_em:
        return a_in
    prior = F(1.0) / F(num_alphas)
    for i in range(num_alphas):
        a_in[i] = prior if init_uniform else F(F(a_in[i] + F(0.5)) * F(1e-3))

    def abund(idx):
        if usa_offsets is None:
            return a_in[idx]
        uo, ao = usa_offsets
        if idx >= ao:
            return F(F(a_in[idx - uo] + a_in[idx - ao]) + a_in[idx])
        if idx >= uo:
            return F(a_in[idx + uo] + a_in[idx])
        return F(a_in[idx + ao] + a_in[idx])

    it, converged, last_round = 0, True, False
    while it < 2 or (it < 100 and not converged) or last_round:
        for eq, ct in cell_data:
            lab = classes[eq]
            if len(lab) > 1:
                den = F(0)
                for l in lab:
                    den = F(den + abund(l))
                if den > 0:
                    inv = F(F(ct) / den)
                    for l in lab:
                        a_out[l] = F(a_out[l] + F(abund(l) * inv))
            else:
                a_out[lab[0]] = F(a_out[lab[0]] + F(ct))
        converged = True
        for i in range(num_alphas):
            if a_out[i] > F(1e-2) and abs(F(a_in[i] - a_out[i])) > F(1e-2):
                converged = False
            a_in[i] = a_out[i]
            a_out[i] = F(0)
        it += 1
        if last_round:
            break
        if it >= 2 and converged:
            a_in[a_in < F(0.01)] = F(0)
            last_round = True
    a_in[a_in < F(0.01)] = F(0)
    return a_in


CLASSES_A = [[0], [1], [0, 1], [1, 2], [2, 3, 4]]


@pytest.mark.parametrize("init_uniform", [True, False])
@pytest.mark.parametrize("cell_data", [[], [(0, 7)], [(0, 20), (1, 4), (2, 8), (3, 1), (4, 2)]])
def test_em_subset_matches_dense_gene_mode(init_uniform, cell_data):
    # src/em.rs:1175-1184
    got = oracle_lib.em_subset(CLASSES_A, cell_data, init_uniform, 8)
    want = dense_reference(CLASSES_A, cell_data, init_uniform, 8)
    assert np.array_equal(got, want), (got, want)


CLASSES_USA = [[0, 1], [3, 4], [6, 7], [0, 4, 8], [2]]


@pytest.mark.parametrize("init_uniform", [True, False])
@pytest.mark.parametrize("cell_data", [[(0, 5)], [(1, 5)], [(2, 5)], [(0, 3), (1, 4), (2, 5), (3, 7), (4, 2)]])
def test_em_subset_matches_dense_usa(init_uniform, cell_data):
    # src/em.rs:1186-1202: 3 genes, S [0,3), U [3,6), A [6,9), offsets (3, 6)
    got = oracle_lib.em_subset(CLASSES_USA, cell_data, init_uniform, 9, usa_offsets=(3, 6))
    want = dense_reference(CLASSES_USA, cell_data, init_uniform, 9, usa_offsets=(3, 6))
    assert np.array_equal(got, want), (got, want)
    assert got.sum() > 0


def test_em_output_threshold_contains_exact_zero():
    # src/em.rs:1204-1215
    classes = [[0], [0, 1], [1, 2], [2, 3, 4]]
    cd = [(0, 10_000), (1, 1), (2, 1), (3, 1)]
    got = oracle_lib.em_subset(classes, cd, False, 6)
    want = dense_reference(classes, cd, False, 6)
    assert np.array_equal(got, want)
    assert (got == 0.0).any()


def test_em_support_only_mass():
    # src/em.rs:1217-1241 (deterministic part): classes {1,2},{2,3}; counts (4, 0)
    got = oracle_lib.em_subset([[1, 2], [2, 3]], [(0, 4), (1, 0)], True, 8)
    assert got[1] + got[2] > 0
    assert got[0] == 0            # outside the support {1,2,3}
    assert np.all(got[4:] == 0)


def test_em_only_unique_returns_singleton_tally():
    got = oracle_lib.em_subset(CLASSES_A, [(0, 20), (1, 4), (2, 8)], False, 8, only_unique=True)
    assert got.tolist() == [20, 4, 0, 0, 0, 0, 0, 0]
    got = oracle_lib.em_dense(CLASSES_A, [20, 4, 8, 1, 2], False, 8, only_unique=True)
    assert got.tolist() == [20, 4, 0, 0, 0, 0, 0, 0]


def test_em_dense_conserves_mass_when_nothing_is_clamped():
    got = oracle_lib.em_dense([[0], [1], [0, 1]], [10, 10, 6], False, 4)
    assert abs(float(got.sum()) - 26.0) < 1e-3
    assert abs(got[0] - 13.0) < 1e-2 and abs(got[1] - 13.0) < 1e-2


def test_hamming_2bit():
    # src/utils.rs:389-393
    h = oracle_lib.lib().afq_oracle_hamming
    assert h(0, 0) == 0
    assert h(0b00, 0b01) == 1 and h(0b00, 0b10) == 1 and h(0b00, 0b11) == 1
    assert h(0b0100, 0b0001) == 2
    assert h(0xFFFFFF, 0x000000) == 12


def test_single_base_substitutions_match_the_reference_snp_vector():
    # src/utils.rs:1207-1213 (test_get_all_snps): the SNP neighbours of barcode 7 at length 3. The PUG
    # kernels enumerate a UMI's 1-Hamming neighbourhood as u ^ (d << 2*pos), d = 1..3 (afq_pug.cuh
    # visit_neighbours, afq_pugs.cuh phase 4): same 2-bit packing, same set.
    golden = [3, 4, 5, 6, 11, 15, 23, 39, 55]
    mine = sorted(7 ^ (d << (2 * pos)) for pos in range(3) for d in (1, 2, 3))
    assert mine == golden
    h = oracle_lib.lib().afq_oracle_hamming
    assert all(h(7, x) == 1 for x in golden)
    assert sorted(x for x in range(64) if h(7, x) == 1) == golden
    # neighbours of barcode 7 at length 3 (src/utils.rs:1207-1256 conventions): all 9 SNPs at distance 1
    base = 7
    for pos in range(3):
        for d in (1, 2, 3):
            nb = base ^ (d << (2 * pos))
            assert h(base, nb) == 1


# ---- synthetic cells from the reference's integration tests ------------------------------
def make_packed(idx, length):
    # tests/multi_barcode_integration.rs:36-41
    return (idx * 2654435761) & ((1 << (2 * length)) - 1)


def ten_gene_t2g():
    return np.arange(10, dtype=np.uint32)  # write_tg_map: 10 single-transcript genes


def ambiguous_tiny_cells(num_cells=4, reads_per_cell=8):
    # create_ambiguous_tiny_cell_rad, tests/multi_barcode_integration.rs:1350-1390
    cells = []
    for ci in range(num_cells):
        recs = []
        for ri in range(reads_per_cell):
            umi = make_packed(ci * 100 + ri, 12)
            recs.append((umi, [ri % 10, (ri + 1) % 10]))
        cells.append(recs)
    return CellBatch.from_cells(cells)


def run(res, batch, t2g, **kw):
    o = QuantOpts(resolution=res, num_gene_ids=kw.pop("num_gene_ids", len(set(t2g.tolist()))),
                  num_rows=kw.pop("num_rows", len(set(t2g.tolist()))), **kw)
    return oracle_lib.oracle_quant(o, t2g, batch, n_threads=2)


def test_tiny_cell_fast_path_does_not_override_requested_resolution():
    # tests/multi_barcode_integration.rs:1404-1556 (large_graph_thresh(0) as in the test)
    b, t2g = ambiguous_tiny_cells(), ten_gene_t2g()
    cr = run("cr-like", b, t2g, small_thresh=100, large_graph_thresh=0)
    em_fast = run("parsimony-em", b, t2g, small_thresh=100, large_graph_thresh=0)
    em_full = run("parsimony-em", b, t2g, small_thresh=0, large_graph_thresh=0)
    assert int((cr.flags & FLAG_TINY != 0).sum()) == 4
    assert int((em_fast.flags & FLAG_TINY != 0).sum()) == 4
    assert int((em_full.flags & FLAG_TINY != 0).sum()) == 0
    assert float(cr.val.sum()) == 0.0
    assert int((cr.flags & FLAG_EMPTY != 0).sum()) == 4
    assert float(em_full.val.sum()) > 0.0
    assert float(em_full.val.sum()) > float(em_fast.val.sum())


def test_trivial_single_ref_reads():
    # tests/multi_barcode_integration.rs:721-863: 8 reads/cell, read i -> ref i % 10, distinct UMIs
    cells = [[(make_packed(ci * 100 + ri, 12), [ri % 10]) for ri in range(8)] for ci in range(3)]
    b, t2g = CellBatch.from_cells(cells), ten_gene_t2g()
    for res, st in (("trivial", 0), ("trivial", 100), ("cr-like", 100), ("parsimony", 0)):
        r = run(res, b, t2g, small_thresh=st)
        for c in range(3):
            col, val = r.row(c)
            assert col.tolist() == list(range(8)) and val.tolist() == [1.0] * 8, (res, st)
        assert r.num_expr.tolist() == [8, 8, 8]
        assert r.sum_umi.tolist() == [8.0, 8.0, 8.0]
        assert r.num_over_mean.tolist() == [0, 0, 0]


# ---- hand-derived semantics ----------------------------------------------------------------
def test_crlike_argmax_with_ties():
    # genes = refs (identity t2g). UMI 1: gene 0 x3, gene 1 x1 -> gene 0. UMI 2: g0 x1, g1 x1 -> tie, dropped.
    # UMI 3: read {0,1} x2 and read {1} x1 -> W(1)=3 > W(0)=2 -> gene 1.
    recs = [(1, [0]), (1, [0]), (1, [0]), (1, [1]), (2, [0]), (2, [1]), (3, [0, 1]), (3, [0, 1]), (3, [1])]
    t2g = ten_gene_t2g()
    for st in (100, 0):  # tiny path and the <=250 path must agree
        r = run("cr-like", CellBatch.from_cells([recs]), t2g, small_thresh=st)
        col, val = r.row(0)
        assert col.tolist() == [0, 1] and val.tolist() == [1.0, 1.0]
    # > 250 records exercises the eq-class path (src/quant.rs:853): replicate every read 40x
    big = [r_ for r_ in recs for _ in range(40)]
    r = run("cr-like", CellBatch.from_cells([big]), t2g, small_thresh=0)
    col, val = r.row(0)
    assert col.tolist() == [0, 1] and val.tolist() == [1.0, 1.0]


def test_crlike_usa_rules():
    # 2 genes, USA ids: g0 S=0 U=1, g1 S=2 U=3; tid i -> id i. num_rows = 6: S [0,2) U [2,4) A [4,6)
    t2g = np.arange(4, dtype=np.uint32)
    cells = [[(1, [0]),            # S g0 -> slot 0
              (2, [1]),            # U g0 -> slot 2
              (3, [0, 1]),         # S+U g0 -> A slot 4
              (4, [0, 3]),         # S g0 + U g1 -> prefer spliced: slot 0
              (5, [0, 2]),         # S g0 + S g1 -> dropped
              (6, [1, 3]),         # U g0 + U g1 -> dropped
              (7, [0, 1, 3])]]     # S g0, U g0, U g1 -> exactly one spliced with partner -> A slot 4
    for st in (100, 0):
        r = run("cr-like", CellBatch.from_cells(cells), t2g, small_thresh=st, usa_mode=True, num_gene_ids=4, num_rows=6)
        col, val = r.row(0)
        assert col.tolist() == [0, 2, 4] and val.tolist() == [2.0, 1.0, 2.0], (st, col, val)


def test_parsimony_collapses_one_edit_neighbours_directionally():
    t2g = ten_gene_t2g()
    u = 0b0110_1100_0011  # arbitrary 6-base UMI
    v = u ^ 0b01           # 1 substitution
    w = u ^ 0b0101         # 2 substitutions from u, 1 from v
    # counts: u x5, v x1 (u -> v directed), all on gene 3: one molecule
    recs = [(u, [3])] * 5 + [(v, [3])]
    r = run("parsimony", CellBatch.from_cells([recs]), t2g, small_thresh=0)
    assert r.row(0)[0].tolist() == [3] and r.row(0)[1].tolist() == [1.0]
    # cr-like keeps them apart
    r = run("cr-like", CellBatch.from_cells([recs]), t2g, small_thresh=0)
    assert r.row(0)[1].tolist() == [2.0]
    # chain u(5) -> v(2) <-> w(2): u reaches v and w: one molecule
    recs = [(u, [3])] * 5 + [(v, [3])] * 2 + [(w, [3])] * 2
    r = run("parsimony", CellBatch.from_cells([recs]), t2g, small_thresh=0)
    assert r.row(0)[1].tolist() == [1.0]
    # exact-UMI mode: no 1-edit edges -> 3 molecules
    r = run("parsimony", CellBatch.from_cells([recs]), t2g, small_thresh=0, pug_exact_umi=True)
    assert r.row(0)[1].tolist() == [3.0]
    # different genes never share a transcript: no edge even at distance 1
    recs = [(u, [3])] * 5 + [(v, [4])]
    r = run("parsimony", CellBatch.from_cells([recs]), t2g, small_thresh=0)
    assert r.row(0)[0].tolist() == [3, 4] and r.row(0)[1].tolist() == [1.0, 1.0]


def test_parsimony_monochromatic_cover_and_label_intersection():
    # transcripts 0,1 -> gene 0; 2 -> gene 1. Same UMI in classes {0,1}, {1,2}: they share t1,
    # distance 0 -> bidirected; the MCC through t1 covers both -> label intersection {1} -> gene 0.
    t2g = np.array([0, 0, 1], dtype=np.uint32)
    recs = [(9, [0, 1]), (9, [1, 2])]
    r = run("parsimony", CellBatch.from_cells([recs]), t2g, small_thresh=0, num_gene_ids=2, num_rows=2)
    assert r.row(0)[0].tolist() == [0] and r.row(0)[1].tolist() == [1.0]
    # large_graph_thresh = 1 forces the cr-like fallback on that 2-vertex component and flags the cell:
    # W(g0) = 2, W(g1) = 1 -> gene 0
    r = run("parsimony", CellBatch.from_cells([recs]), t2g, small_thresh=0, large_graph_thresh=1, num_gene_ids=2, num_rows=2)
    assert r.flags[0] & FLAG_ALT
    assert r.row(0)[0].tolist() == [0] and r.row(0)[1].tolist() == [1.0]


def test_empty_and_ragged_cells():
    t2g = ten_gene_t2g()
    cells = [[], [(5, [1])], [(5, [1, 2])], [(7, [])]]
    r = run("cr-like", CellBatch.from_cells(cells), t2g, small_thresh=100)
    assert r.num_expr.tolist() == [0, 1, 0, 0]
    assert (r.flags & FLAG_EMPTY != 0).tolist() == [True, False, True, True]


# ---- tie census of the parsimony cover ------------------------------------------
def _census(cells, t2g, n_genes, res="parsimony"):
    o = QuantOpts(resolution=res, num_gene_ids=n_genes, num_rows=n_genes, small_thresh=0)
    return oracle_lib.tie_census(o, np.asarray(t2g, dtype=np.uint32), CellBatch.from_cells(cells))


def test_tie_census_hand_built_components():
    u0 = 0b0000; u1 = u0 ^ 1; u2 = u1 ^ (1 << 2)          # u0 - u1 - u2: a 1-edit path, u0 / u2 two edits apart
    t2g = [0, 1, 2]
    # (a) no multi-vertex component: nothing to break
    d = _census([[(5, [0]), (900, [1]), (77777, [2])]], t2g, 3)
    assert d["components_multi"] == 0 and d["molecules"] == 3 and d["components_tie_on_path"] == 0
    # (b) a tie between two maximal MCCs that ends in the same gene-label multiset either way
    d = _census([[(u0, [0]), (u1, [0, 1]), (u2, [1])]], t2g, 3)
    assert d["components_multi"] == 1 and d["components_tie_on_path"] == 1
    assert d["components_label_sensitive"] == 0 and d["molecules_changed_worst_case"] == 0 and d["molecules"] == 2
    # (c) the same tie, but the two choices leave different labels behind: {[0], [1,2]} vs {[0], [1]}
    d = _census([[(u0, [0]), (u1, [0, 1]), (u2, [1, 2])]], t2g, 3)
    assert d["components_tie_on_path"] == 1 and d["components_label_sensitive"] == 1 and d["components_count_sensitive"] == 1
    assert d["molecules_changed_worst_case"] == 1 and d["cells_count_sensitive"] == 1 and d["components_capped"] == 0
    # (d) directed edges remove the tie: u1 has >= 2x the reads of its neighbours, so only u1 reaches both
    d = _census([[(u0, [0]), (u1, [0, 1]), (u1, [0, 1]), (u2, [1, 2])]], t2g, 3)
    assert d["components_multi"] == 1 and d["components_tie_on_path"] == 0


def test_tie_census_bounds_the_parsimony_deviation_on_the_bench_shapes():
    # DESIGN.md §2 quotes these: the share of molecules whose gene label could differ from a real reference run
    # (which visits start vertices in hash order) is well below 1 % on C3 and C5
    for cfg, res, bound in (("C3", "parsimony", 0.005), ("C5", "parsimony-em", 0.012)):
        spec = synth.config_spec(cfg)
        b = synth.generate(spec, 0, 600)
        o = QuantOpts(resolution=res, usa_mode=spec.usa_mode, num_gene_ids=spec.num_gene_ids, num_rows=spec.num_rows)
        d = oracle_lib.tie_census(o, synth.tid_to_gid(spec), b)
        assert d["components_capped"] == 0
        assert 0 < d["frac_molecules_changed_worst_case"] < bound, d
        assert d["molecules_changed_worst_case"] <= d["molecules_label_sensitive"] <= d["molecules_tie_on_path"]
        # the canonical-order molecule count equals what oracle_quant reports
        if res == "parsimony":
            r = oracle_lib.oracle_quant(QuantOpts(resolution="parsimony-em", usa_mode=spec.usa_mode, num_gene_ids=spec.num_gene_ids,
                                                  num_rows=spec.num_rows), synth.tid_to_gid(spec), b)
            tiny = np.diff(b.cell_rec_offsets.astype(np.int64)) < 100
            assert abs(float(r.sum_umi[~tiny].sum()) - d["molecules"]) <= 1e-3 * d["molecules"]
