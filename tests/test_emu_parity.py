"""CPU-tier check of the CUDA kernels' LOGIC: the product's device code, compiled for the
host against the test-only emulator (tests/emu), must reproduce the oracle bit-for-bit.
(The real-GPU parity tests are in test_gpu_parity.py; this tier exists because the build
container has no GPU, and it is also the debugging harness for the kernels.)"""
import numpy as np
import pytest

import cases
import emu_lib
import oracle_lib
from alevin_fry_b200 import CellBatch, QuantOpts, FLAG_ALT
import synth

ALL_RES = ["cr-like", "trivial", "parsimony", "parsimony-gene", "cr-like-em", "parsimony-em", "parsimony-gene-em"]


def check(opts, t2g, b, tag=""):
    got = emu_lib.emu_quant(opts, t2g, b)
    want = oracle_lib.oracle_quant(opts, t2g, b, n_threads=2)
    assert np.array_equal(got.row_ptr, want.row_ptr), tag
    assert np.array_equal(got.col, want.col), tag
    # same operation order and roundings => bit-identical even for EM
    assert np.array_equal(got.val, want.val), tag
    assert np.array_equal(got.sum_umi, want.sum_umi), tag
    assert np.array_equal(got.max_umi, want.max_umi), tag
    assert np.array_equal(got.num_expr, want.num_expr), tag
    assert np.array_equal(got.num_over_mean, want.num_over_mean), tag
    assert np.array_equal(got.flags, want.flags), tag
    return got


def opts_for(spec, res, **kw):
    return QuantOpts(resolution=res, usa_mode=spec.usa_mode, num_gene_ids=spec.num_gene_ids, num_rows=spec.num_rows,
                     umi_len=spec.umi_len, **kw)


@pytest.mark.parametrize("res", ALL_RES)
def test_emu_mini_c2(res):
    spec = synth.SynthSpec(reads_mean=400.0)
    b = synth.generate(spec, 0, 12)
    check(opts_for(spec, res), synth.tid_to_gid(spec), b, res)


@pytest.mark.parametrize("res", ["cr-like", "cr-like-em", "parsimony", "parsimony-em"])
def test_emu_mini_c4_usa(res):
    spec = synth.SynthSpec(usa_mode=True, reads_mean=400.0, n_genes=2000)
    b = synth.generate(spec, 0, 10)
    check(opts_for(spec, res), synth.tid_to_gid(spec), b, res)


@pytest.mark.parametrize("res", ["parsimony", "parsimony-em", "parsimony-gene", "cr-like-em"])
@pytest.mark.parametrize("thresh", [1000, 40, 3, 0])
def test_emu_dense_umi_space_components(res, thresh):
    # 4-base UMIs (256 values) and few genes: large 1-Hamming components. thresh selects the
    # mid-size cooperative cover (33..thresh), the per-thread cover (<= 32) or the cr-like
    # fallback (> thresh, AFQ_FLAG_ALT)
    spec = synth.SynthSpec(n_genes=40, umi_len=4, reads_mean=500.0, reads_per_umi=1.5, umi_err=0.05, zipf_s=0.7)
    b = synth.generate(spec, 0, 6)
    got = check(opts_for(spec, res, large_graph_thresh=thresh, small_thresh=0), synth.tid_to_gid(spec), b, f"{res}/{thresh}")
    if thresh <= 3 and res.startswith("parsimony"):
        assert (got.flags & FLAG_ALT).any()


@pytest.mark.parametrize("res", ["parsimony", "parsimony-em"])
def test_emu_exact_umi(res):
    spec = synth.SynthSpec(n_genes=200, umi_len=5, reads_mean=300.0)
    b = synth.generate(spec, 0, 6)
    check(opts_for(spec, res, pug_exact_umi=True, small_thresh=0), synth.tid_to_gid(spec), b, res)


@pytest.mark.parametrize("res", ALL_RES)
def test_emu_edge_cases(res):
    t2g = np.arange(10, dtype=np.uint32)
    cells = [[], [(5, [1])], [(5, [1, 2])], [(7, [])], [(1, [0])] * 120 + [(2, [0, 3])] * 2, [],
             [(9, [0, 1, 2, 3, 4]), (9, [0, 1, 2, 3, 5]), (8, [0, 1, 2]), (8, [0, 1, 2, 3])]]
    b = CellBatch.from_cells(cells)
    for st in (100, 0):
        check(QuantOpts(resolution=res, num_gene_ids=10, num_rows=10, small_thresh=st), t2g, b, f"{res}/{st}")


@pytest.mark.parametrize("res", ALL_RES)
@pytest.mark.parametrize("usa", [False, True])
def test_emu_tiny_cells_take_the_crlike_path_whatever_the_resolution(res, usa):
    # src/quant.rs:780-846: a cell below --small-thresh is resolved cr-like regardless of -r. These cells
    # separate the two semantics for `trivial` (a UMI seen on two genes: trivial counts both, cr-like drops the tie)
    t2g = np.arange(10, dtype=np.uint32)
    cells = [[(5, [1]), (5, [2])], [(5, [1]), (5, [1]), (5, [2]), (6, [3])], [(5, [2, 3]), (5, [2]), (7, [4, 5])],
             [(1, [0])] * 99 + [(1, [1])], [(1, [0])] * 100 + [(1, [1])]]
    b = CellBatch.from_cells(cells)
    n = 4 if usa else 10
    for st in (100, 3, 0):
        check(QuantOpts(resolution=res, usa_mode=usa, num_gene_ids=10, num_rows=15 if usa else 10, small_thresh=st), t2g, b, f"{res}/{st}/{usa}")


def test_emu_uniform_init_and_many_gene_labels():
    spec = synth.SynthSpec(n_genes=60, reads_mean=300.0, p_multi2=0.3, p_multi3=0.3, reads_per_umi=2.0)
    b = synth.generate(spec, 0, 6)
    t2g = synth.tid_to_gid(spec)
    for res in ("cr-like-em", "parsimony-em"):
        check(opts_for(spec, res, init_uniform=True, small_thresh=0), t2g, b, res)
        check(opts_for(spec, res, small_thresh=0), t2g, b, res)


@pytest.mark.parametrize("res", ["cr-like", "trivial"])
def test_emu_resolve_hot_buckets_and_large_gene_axis(res, monkeypatch):
    # few genes, many molecules: winner buckets hold hundreds of equal slots (cooperative hot-bucket path)
    spec = synth.SynthSpec(n_genes=40, reads_mean=3000.0, reads_per_umi=1.5, zipf_s=1.3)
    b = synth.generate(spec, 0, 4)
    check(opts_for(spec, res), synth.tid_to_gid(spec), b, res + "/hot")
    # 300k genes: slot >> shift needs > 8 low bits -> winners are ordered directly (rank / bitonic path)
    spec = synth.SynthSpec(n_genes=300_000, reads_mean=2500.0, reads_per_umi=1.2, zipf_s=0.3)
    b = synth.generate(spec, 0, 3)
    check(opts_for(spec, res), synth.tid_to_gid(spec), b, res + "/wide")


@pytest.mark.parametrize("force_bin", ["2", "3", "4", "6"])
def test_emu_forced_arena(monkeypatch, force_bin):
    monkeypatch.setenv("AFQ_FORCE_BIN", force_bin)
    spec = synth.SynthSpec(reads_mean=600.0)
    b = synth.generate(spec, 50, 6)
    for res in ("cr-like", "trivial"):
        check(opts_for(spec, res), synth.tid_to_gid(spec), b, f"{res}/bin{force_bin}")


def test_emu_overflow_requeue():
    # no duplication, several genes per read: distinct pairs exceed 75 % of the arena chosen from
    # the record count, so cells are re-queued on the next arena size
    spec = synth.SynthSpec(reads_mean=900.0, reads_per_umi=1.0, p_multi2=0.45, p_multi3=0.45, lognorm_sigma=0.3)
    b = synth.generate(spec, 0, 6)
    check(opts_for(spec, "cr-like"), synth.tid_to_gid(spec), b, "overflow")


def test_emu_na8_compact_offsets():
    # afq_batch.rec_na8: per-record alignment counts instead of CSR offsets; several scan tiles
    spec = synth.SynthSpec(reads_mean=1500.0)
    b = synth.generate(spec, 0, 9)
    assert b.n_records > 2 * 4096
    t2g = synth.tid_to_gid(spec)
    for res in ("cr-like", "parsimony"):
        o = opts_for(spec, res)
        got = emu_lib.emu_quant(o, t2g, b, use_na8=True)
        want = emu_lib.emu_quant(o, t2g, b)
        assert np.array_equal(got.row_ptr, want.row_ptr) and np.array_equal(got.col, want.col) and np.array_equal(got.val, want.val)
    with pytest.raises(ValueError):
        CellBatch.from_cells([[(1, list(range(300)))]]).na8()


def test_emu_pack24_wire_arrays():
    # afq_batch.rec_umi24 / refs24: 3-byte UMIs and transcript ids widened on the device (k_unpack24);
    # record / ref counts that are not multiples of 4 exercise the tail path
    spec = synth.SynthSpec(reads_mean=700.0)
    t2g = synth.tid_to_gid(spec)
    for ncell in (1, 7):
        b = synth.generate(spec, 5, ncell)
        for res in ("cr-like", "parsimony-em"):
            o = opts_for(spec, res)
            got = emu_lib.emu_quant(o, t2g, b, use_na8=True, use_pack24=True)
            want = emu_lib.emu_quant(o, t2g, b)
            assert np.array_equal(got.row_ptr, want.row_ptr) and np.array_equal(got.col, want.col) and np.array_equal(got.val, want.val)
    tiny = CellBatch.from_cells([[(0xABCDEF, [5, 0xFFFFFF % len(t2g)]), (3, [1]), (0xFFFFFF, [2, 3, 4])]])
    u24, r24 = tiny.pack24()
    assert bytes(u24[:9]) == bytes([0xEF, 0xCD, 0xAB, 3, 0, 0, 0xFF, 0xFF, 0xFF])
    o = opts_for(spec, "cr-like")
    got, want = emu_lib.emu_quant(o, t2g, tiny, use_pack24=True), emu_lib.emu_quant(o, t2g, tiny)
    assert np.array_equal(got.col, want.col) and np.array_equal(got.val, want.val)
    with pytest.raises(ValueError):
        CellBatch.from_cells([[(1 << 24, [1])]]).pack24()


def _wide_cells(rng, n_cells, n_rec, na_lo, na_hi, n_tx, n_umi, p_empty=0.0):
    cells = []
    for _ in range(n_cells):
        recs = []
        for _ in range(n_rec):
            if rng.random() < p_empty:
                recs.append((int(rng.integers(n_umi)), []))
                continue
            na = int(rng.integers(na_lo, na_hi + 1))
            refs = np.sort(rng.choice(n_tx, size=min(na, n_tx), replace=False))
            recs.append((int(rng.integers(n_umi)), [int(x) for x in refs]))
        cells.append(recs)
    return cells


@pytest.mark.parametrize("res", ["cr-like", "trivial", "cr-like-em"])
def test_emu_flat_alignment_loop_shapes(res):
    # phase 1 of the resolve runs one lane per alignment: records wider than a bitmap word (head
    # search walks back), a NON-monotone tid->gid map (duplicate genes far apart inside a record),
    # records without alignments and cells whose refs overflow the bitmap (lane-per-record form)
    rng = np.random.default_rng(7)
    n_tx, n_genes = 400, 37
    t2g = rng.integers(0, n_genes, size=n_tx).astype(np.uint32)
    cells = (_wide_cells(rng, 2, 150, 30, 80, n_tx, 40) +            # wide records, flat form
             _wide_cells(rng, 2, 300, 1, 6, n_tx, 60) +               # narrow records
             _wide_cells(rng, 1, 120, 60, 90, n_tx, 30) +             # refs overflow the smallest arena's bitmap
             _wide_cells(rng, 2, 200, 1, 5, n_tx, 50, p_empty=0.1))   # alignment-free records
    b = CellBatch.from_cells(cells)
    check(QuantOpts(resolution=res, num_gene_ids=n_genes, num_rows=n_genes), t2g, b, res)


@pytest.mark.parametrize("res", ["cr-like", "cr-like-em"])
def test_emu_all_ones_umi_and_window_holes(res):
    # a 16-base UMI of all T's is 0xFFFFFFFF, the high half of the table's EMPTY marker: the windowed
    # probe layout leaves empty slots between a UMI's entries, which must not be mistaken for it
    t2g = np.arange(64, dtype=np.uint32)
    U = 0xFFFFFFFF
    recs = [(U, [1])] * 3 + [(U, [7])] * 5 + [(U, [9])] * 5 + [(U - 1, [7])] * 2 + [(5, [3, 4, 12])] * 4 + [(6, [3])]
    recs = recs * 6 + [(i * 2654435761 % (1 << 32), [i % 64]) for i in range(150)]
    b = CellBatch.from_cells([recs, recs[::-1]])
    check(QuantOpts(resolution=res, num_gene_ids=64, num_rows=64, umi_len=16), t2g, b, res)


# ---- k_pug_smem (shared-memory parsimony kernel) -------------------------------------------------
@pytest.mark.parametrize("res", ["parsimony", "parsimony-em", "parsimony-gene", "parsimony-gene-em", "cr-like-em"])
def test_emu_pug_smem_takes_the_cells(res, monkeypatch):
    spec = synth.SynthSpec(reads_mean=400.0)
    b = synth.generate(spec, 0, 12)
    t2g = synth.tid_to_gid(spec)
    check(opts_for(spec, res), t2g, b, res)
    cnt = emu_lib.last_counts()
    n_ps = sum(cnt[emu_lib.LIST_PS0:emu_lib.LIST_PS0 + 4])
    n_tiny = sum(cnt[:7])
    assert n_ps > 0 and n_ps + n_tiny == b.n_cells, cnt      # every non-tiny cell binned to k_pug_smem
    assert cnt[emu_lib.LIST_GE_NORMAL] == 0, cnt                # ... and none handed back
    # same answer from the global-arena kernel alone
    monkeypatch.setenv("AFQ_NO_PS", "1")
    check(opts_for(spec, res), t2g, b, res + "/no-ps")
    cnt = emu_lib.last_counts()
    assert sum(cnt[emu_lib.LIST_PS0:emu_lib.LIST_PS0 + 4]) == 0 and cnt[emu_lib.LIST_GE_NORMAL] > 0, cnt


@pytest.mark.parametrize("res", ["parsimony", "parsimony-em"])
@pytest.mark.parametrize("limit", [600, 2500, 4000])
def test_emu_pug_smem_hands_back_what_does_not_fit(res, limit, monkeypatch):
    # a smaller arena than the binning assumed: cells fail at different points (layout, dense
    # vertices, EM back end) and must come out of k_gene_eqc with identical results
    monkeypatch.setenv("AFQ_PS_LIMIT_WORDS", str(limit))
    spec = synth.SynthSpec(reads_mean=400.0)
    b = synth.generate(spec, 0, 12)
    check(opts_for(spec, res), synth.tid_to_gid(spec), b, f"{res}/{limit}")
    cnt = emu_lib.last_counts()
    assert cnt[emu_lib.LIST_GE_NORMAL] > 0, cnt


@pytest.mark.parametrize("res", ["parsimony", "parsimony-em", "parsimony-gene"])
def test_emu_pug_smem_big_components_fall_back(res):
    # 4-base UMIs: components beyond 32 vertices are handed to the global-arena kernel (mid-size
    # cooperative cover there); smaller ones are covered by k_pug_smem's per-thread cover
    spec = synth.SynthSpec(n_genes=40, umi_len=4, reads_mean=500.0, reads_per_umi=1.5, umi_err=0.05, zipf_s=0.7)
    b = synth.generate(spec, 0, 6)
    check(opts_for(spec, res, small_thresh=0), synth.tid_to_gid(spec), b, res)
    cnt = emu_lib.last_counts()
    assert sum(cnt[emu_lib.LIST_PS0:emu_lib.LIST_PS0 + 4]) == b.n_cells, cnt
    spec2 = synth.SynthSpec(n_genes=300, umi_len=6, reads_mean=600.0, reads_per_umi=1.5, umi_err=0.05)
    b2 = synth.generate(spec2, 0, 6)   # 4096 UMI values: small multi-vertex components, covered in shared memory
    check(opts_for(spec2, res, small_thresh=0), synth.tid_to_gid(spec2), b2, res + "/umi6")
    cnt2 = emu_lib.last_counts()
    assert cnt2[emu_lib.LIST_GE_NORMAL] < b2.n_cells, cnt2


@pytest.mark.parametrize("res", ["parsimony", "parsimony-em", "cr-like-em"])
def test_emu_pug_smem_usa_and_record_order(res):
    spec = synth.SynthSpec(usa_mode=True, reads_mean=500.0, n_genes=2000)
    b = synth.generate(spec, 0, 8)
    t2g = synth.tid_to_gid(spec)
    got = check(opts_for(spec, res), t2g, b, res)
    if res != "cr-like-em":   # (USA cr-like-em stays on k_gene_eqc, afq_pipeline.cuh)
        assert sum(emu_lib.last_counts()[emu_lib.LIST_PS0:emu_lib.LIST_PS0 + 4]) > 0
    # record order inside a cell must not matter (canonical orders)
    rng = np.random.default_rng(5)
    cells = []
    for c in range(b.n_cells):
        r0, r1 = int(b.cell_rec_offsets[c]), int(b.cell_rec_offsets[c + 1])
        recs = [(int(b.rec_umi32[r]), b.refs[b.rec_ref_offsets[r]:b.rec_ref_offsets[r + 1]].tolist()) for r in range(r0, r1)]
        rng.shuffle(recs)
        cells.append(recs)
    got2 = emu_lib.emu_quant(opts_for(spec, res), t2g, CellBatch.from_cells(cells))
    assert np.array_equal(got.col, got2.col) and np.array_equal(got.val, got2.val)


@pytest.mark.parametrize("res", ["parsimony", "parsimony-em", "parsimony-gene"])
def test_emu_pug_smem_warp_cover_on_star_components(res):
    rng = np.random.default_rng(11)
    n_genes = 50
    t2g = np.repeat(np.arange(n_genes, dtype=np.uint32), 3)
    b = CellBatch.from_cells(cases.star_cells(rng, 5, n_genes))
    o = QuantOpts(resolution=res, num_gene_ids=n_genes, num_rows=n_genes, umi_len=12)
    check(o, t2g, b, res)
    cnt = emu_lib.last_counts()
    assert sum(cnt[emu_lib.LIST_PS0:emu_lib.LIST_PS0 + 4]) == b.n_cells and cnt[emu_lib.LIST_GE_NORMAL] == 0, cnt


@pytest.mark.parametrize("res", ["parsimony", "parsimony-em"])
def test_emu_pug_global_arena_variant(res, monkeypatch):
    # cells beyond the shared-memory arenas run the same code on a global-memory arena (variant 3);
    # a tiny arena limit pushes ordinary cells there
    spec = synth.SynthSpec(fixed_reads=10000, n_genes=3000)
    b = synth.generate(spec, 0, 2)
    check(opts_for(spec, res), synth.tid_to_gid(spec), b, res)
    cnt = emu_lib.last_counts()
    assert cnt[emu_lib.LIST_PS0 + 3] > 0, cnt
    assert cnt[emu_lib.LIST_GE_NORMAL] == 0, cnt


@pytest.mark.parametrize("res", ["cr-like", "trivial"])
def test_emu_crlike_record_shapes(res):
    n_genes, t2g, cells = cases.record_shape_cells(np.random.default_rng(5))
    b = CellBatch.from_cells(cells)
    check(QuantOpts(resolution=res, num_gene_ids=n_genes, num_rows=n_genes, umi_len=12), t2g, b, res)


@pytest.mark.parametrize("res", ["parsimony", "parsimony-em", "cr-like-em"])
def test_emu_arena_pools_planned_on_the_device(res, monkeypatch):
    # the afq_submit form of the pipeline: the control block is not read back after the binning, the global-arena kernels
    # take their stride and the CTAs that have an arena from k_plan_arenas' plan in the control block, on fixed pools
    spec = synth.SynthSpec(fixed_reads=10000, n_genes=3000)
    b = synth.generate(spec, 0, 2)
    monkeypatch.setenv("AFQ_EMU_ASYNC", "1")
    check(opts_for(spec, res), synth.tid_to_gid(spec), b, res)            # roomy pools: both CTAs of a grid have an arena
    monkeypatch.setenv("AFQ_PS_LIMIT_WORDS", "3000")                        # ... and with cells handed back to k_gene_eqc
    check(opts_for(spec, res), synth.tid_to_gid(spec), b.slice_cells(0, 1), res)
    monkeypatch.delenv("AFQ_PS_LIMIT_WORDS")
    if res != "parsimony":      # the back end's other shapes plan their arenas with the same arithmetic
        for k, v in (("AFQ_NO_EM_SPLIT", "1"), ("AFQ_BACK_MAX_TIER", "2")):
            monkeypatch.setenv(k, v)
            check(opts_for(spec, res), synth.tid_to_gid(spec), b.slice_cells(0, 1), res + "/" + k)
            monkeypatch.delenv(k)
    # a pool that holds no arena for the largest cell: the batch is flagged, never silently wrong (the host API re-runs it)
    monkeypatch.setenv("AFQ_EMU_POOL_WORDS", "2000")
    with pytest.raises(RuntimeError, match="pool of a global-arena kernel"):
        emu_lib.emu_quant(opts_for(spec, res), synth.tid_to_gid(spec), b)


@pytest.mark.parametrize("res", ["parsimony", "parsimony-em"])
def test_emu_pug_smem_long_labels_take_the_warp_cover(res):
    # labels of more than 32 transcripts do not fit the group cover's position masks: such components
    # are re-routed to the warp-cooperative cover inside the same kernel
    n_genes, t2g, cells = cases.long_label_cells()
    b = CellBatch.from_cells(cells)
    check(QuantOpts(resolution=res, num_gene_ids=n_genes, num_rows=n_genes, umi_len=12), t2g, b, res)
    cnt = emu_lib.last_counts()
    assert sum(cnt[emu_lib.LIST_PS0:emu_lib.LIST_PS0 + 4]) == b.n_cells and cnt[emu_lib.LIST_GE_NORMAL] == 0, cnt


RANDOM_REGIMES = [
    # (seed, n_tx, tx per gene, umi_bits, n_labels, max_label, records lo, hi)
    (101, 24, 3, 6, 6, 3, 120, 260),      # 64 UMIs: very dense, components often beyond 32 (hand-backs)
    (102, 60, 3, 8, 12, 4, 120, 300),     # 256 UMIs
    (103, 90, 3, 10, 20, 6, 150, 400),    # 1024 UMIs: mostly 2..16 vertex components
    (104, 200, 5, 12, 30, 40, 150, 350),  # long labels (up to 40 transcripts, beyond the 32-bit position masks)
    (105, 12, 1, 7, 5, 5, 110, 200),      # every transcript its own gene: multi-gene labels everywhere
]


@pytest.mark.parametrize("regime", RANDOM_REGIMES, ids=lambda r: f"seed{r[0]}")
@pytest.mark.parametrize("res", ["parsimony", "parsimony-em", "parsimony-gene", "cr-like-em"])
def test_emu_random_adversarial_cells(regime, res):
    seed, n_tx, per_gene, umi_bits, n_labels, max_label, lo, hi = regime
    rng = np.random.default_rng(seed)
    n_genes = (n_tx + per_gene - 1) // per_gene
    t2g = (np.arange(n_tx, dtype=np.uint32) // per_gene).astype(np.uint32)
    b = CellBatch.from_cells(cases.random_pug_cells(rng, 8, n_tx, umi_bits, n_labels, max_label, lo, hi))
    umi_len = (umi_bits + 1) // 2
    for thresh in (1000, 6):
        o = QuantOpts(resolution=res, num_gene_ids=n_genes, num_rows=n_genes, umi_len=umi_len, large_graph_thresh=thresh, small_thresh=0)
        check(o, t2g, b, f"{res}/seed{seed}/thresh{thresh}")


@pytest.mark.parametrize("res", ["parsimony", "parsimony-em", "cr-like-em", "cr-like"])
def test_emu_random_adversarial_cells_usa(res):
    rng = np.random.default_rng(207)
    n_genes, per = 20, 4                                   # 3 spliced + 1 unspliced transcript per gene
    t2g = np.array([2 * (t // per) + (1 if t % per == 3 else 0) for t in range(n_genes * per)], dtype=np.uint32)
    b = CellBatch.from_cells(cases.random_pug_cells(rng, 8, n_genes * per, 9, 16, 5, 130, 320))
    o = QuantOpts(resolution=res, usa_mode=True, num_gene_ids=2 * n_genes, num_rows=3 * n_genes, umi_len=5, small_thresh=0)
    check(o, t2g, b, res)


# ---- --sa-model prefer-ambig (src/pugutils.rs:505-641) ----------------------
@pytest.mark.parametrize("res", ["cr-like", "cr-like-em"])
def test_emu_prefer_ambig_usa(res):
    spec = synth.SynthSpec(usa_mode=True, reads_mean=400.0, n_genes=300)
    b = synth.generate(spec, 0, 10)
    t2g = synth.tid_to_gid(spec)
    wta = check(opts_for(spec, res), t2g, b, res + "/wta")
    pa = check(opts_for(spec, res, sa_model="prefer-ambig"), t2g, b, res + "/prefer-ambig")
    # the option must actually change something on USA data with spliced + unspliced evidence
    assert not (np.array_equal(wta.col, pa.col) and np.array_equal(wta.val, pa.val))
    # the tiny-cell fast path is switched off by it (src/quant.rs:794)
    tiny = synth.generate(synth.SynthSpec(usa_mode=True, fixed_reads=40, n_genes=300), 0, 12)
    got = check(opts_for(spec, res, sa_model="prefer-ambig"), t2g, tiny, res + "/prefer-ambig/tiny")
    assert not (got.flags & 1).any()


@pytest.mark.parametrize("res", ["cr-like", "cr-like-em", "parsimony"])
def test_emu_prefer_ambig_adversarial_and_gene_mode(res):
    rng = np.random.default_rng(311)
    n_genes, per = 20, 4
    t2g = np.array([2 * (t // per) + (1 if t % per == 3 else 0) for t in range(n_genes * per)], dtype=np.uint32)
    b = CellBatch.from_cells(cases.random_pug_cells(rng, 8, n_genes * per, 9, 16, 5, 130, 320))
    o = QuantOpts(resolution=res, usa_mode=True, num_gene_ids=2 * n_genes, num_rows=3 * n_genes, umi_len=5, small_thresh=0,
                  sa_model="prefer-ambig")
    check(o, t2g, b, res + "/usa")
    # the library applies the model as given in gene mode too (the CLI host resets it there, as the reference does)
    spec = synth.SynthSpec(reads_mean=300.0, n_genes=200)
    bg = synth.generate(spec, 0, 8)
    check(opts_for(spec, res, sa_model="prefer-ambig"), synth.tid_to_gid(spec), bg, res + "/gene")
