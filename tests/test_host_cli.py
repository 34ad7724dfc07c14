"""Host side (C++ `quantify` drop-in + CLI): CPU-tier argument/contract checks, and GPU-tier
end-to-end runs from a synthetic collated RAD directory compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
from alevin_fry_b200 import QuantOpts, host
import synth


def make_input(tmp_path, spec, n_cells, first=0):
    b = synth.generate(spec, first, n_cells)
    d = tmp_path / "in"
    names = host.write_synth_t2g(str(tmp_path / "t2g.tsv"), spec)
    bcs = host.make_barcodes(first, n_cells)
    host.write_collated_rad(str(d), b, bcs, names, 16, spec.umi_len)
    return b, bcs, str(d), str(tmp_path / "t2g.tsv")


def test_rad_writer_layout(tmp_path):
    spec = synth.SynthSpec(n_genes=5, fixed_reads=3)
    b, bcs, d, _ = make_input(tmp_path, spec, 2)
    raw = open(os.path.join(d, "map.collated.rad"), "rb").read()
    assert raw[0] == 0 and int.from_bytes(raw[1:9], "little") == 15           # is_paired, ref_count
    assert raw[9:11] == (2).to_bytes(2, "little") and raw[11:13] == b"t0"     # first name: u16 len + bytes
    assert b"compressed_ori_refid" in raw and b"cblen" in raw
    import json
    assert json.load(open(os.path.join(d, "collate.json")))["compressed_output"] is False


def test_cli_argument_validation(tmp_path):
    cli = host.CLI_PATH
    r = subprocess.run([cli, "quant", "-i", str(tmp_path), "-m", "x", "-o", "y", "-r", "cr-like", "--umi-edit-dist", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "doesn't currently support 1-edit UMI resolution" in r.stderr   # src/main.rs:672-682
    r = subprocess.run([cli, "quant", "-i", str(tmp_path), "-m", "x", "-o", "y", "-r", "cr-like", "-b", "5"], capture_output=True, text=True)
    assert r.returncode == 1 and "bootstrapping can only be used" in r.stderr                     # src/main.rs:713-728
    r = subprocess.run([cli, "quant", "-i", str(tmp_path), "-m", "x", "-o", "y", "-r", "trivial", "-d"], capture_output=True, text=True)
    assert r.returncode == 1 and "not meaningful in case of Trivial" in r.stderr                  # src/main.rs:705-711
    r = subprocess.run([cli, "quant", "-i", str(tmp_path), "-m", "x", "-o", "y", "-r", "cr-like", "--use-eds"], capture_output=True, text=True)
    assert r.returncode == 1 and "--use-eds is no longer supported" in r.stderr                   # src/main.rs:639-643
    r = subprocess.run([cli, "quant", "-i", str(tmp_path), "-m", "x", "-o", str(tmp_path / "o"), "-r", "CR-LIKE"], capture_output=True, text=True)
    assert r.returncode == 1 and "generate_permit_list.json" in r.stderr                           # src/main.rs:812-820
    r = subprocess.run([cli, "quant", "-i", str(tmp_path), "-m", "x", "-o", "y", "-r", "full"], capture_output=True, text=True)
    assert r.returncode == 1 and "invalid value" in r.stderr


def test_quantify_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    spec = synth.SynthSpec(n_genes=20, fixed_reads=10)
    _, _, d, t2g = make_input(tmp_path, spec, 3)
    with pytest.raises(RuntimeError) as ei:
        host.quantify(d, t2g, str(tmp_path / "out"), "cr-like")
    assert "no CPU fallback" in str(ei.value)
    # t2g problems are reported before any device work
    open(tmp_path / "bad.tsv", "w").write("t0\tg0\n")
    with pytest.raises(RuntimeError) as ei:
        host.quantify(d, str(tmp_path / "bad.tsv"), str(tmp_path / "out"), "cr-like")
    assert "tg-map must contain a gene mapping for all transcripts" in str(ei.value)


def fmt_f32(v):
    """Rust `{}` for f32 as the host writes it (shortest round-trip, fixed notation)."""
    v = np.float32(v)
    if np.isnan(v):
        return "NaN"
    s = np.format_float_positional(v, unique=True, trim="-")
    return s


@pytest.mark.gpu
@pytest.mark.parametrize("res,usa", [("cr-like", False), ("parsimony", False), ("cr-like-em", True), ("trivial", False)])
def test_cli_end_to_end_matches_oracle(tmp_path, res, usa):
    spec = synth.SynthSpec(n_genes=300, reads_mean=300.0, usa_mode=usa)
    b, bcs, d, t2g_path = make_input(tmp_path, spec, 120)
    out = str(tmp_path / "out")
    r = subprocess.run([host.CLI_PATH, "quant", "-i", d, "-m", t2g_path, "-o", out, "-r", res, "-t", "4"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    q = host.load_quant_dir(out)
    t2g = synth.tid_to_gid(spec)
    thresh = 1000 if res.startswith("parsimony") else 0
    o = QuantOpts(resolution=res, usa_mode=usa, num_gene_ids=spec.num_gene_ids, num_rows=spec.num_rows, large_graph_thresh=thresh)
    want = oracle_lib.oracle_quant(o, t2g, b)
    # MTX layout pinned by the reference's tests (tests/infer_matrix_market.rs:54-73; multi_barcode_integration.rs:1491-1498)
    assert q["header"][0] == "%%MatrixMarket matrix coordinate real general"
    assert q["dims"] == (120, spec.num_rows, want.nnz)
    assert q["rows"] == [host.decode_barcode(x, 16) for x in bcs]
    G = spec.n_genes
    names = [f"g{i}" for i in range(G)]
    assert q["cols"] == (names + [n + "-U" for n in names] + [n + "-A" for n in names] if usa else names)
    got_rows = np.array([t[0] for t in q["triplets"]], dtype=np.int64)
    got_cols = np.array([t[1] for t in q["triplets"]], dtype=np.uint32)
    exp_rows = np.repeat(np.arange(120), np.diff(want.row_ptr.astype(np.int64)))
    assert np.array_equal(got_rows, exp_rows) and np.array_equal(got_cols, want.col)
    if res.endswith("-em"):
        np.testing.assert_allclose(np.array([float(t[2]) for t in q["triplets"]], dtype=np.float32), want.val, rtol=1e-5)
    else:
        assert [t[2] for t in q["triplets"]] == [fmt_f32(v) for v in want.val]   # byte-identical value text
    # featureDump: 9 columns, header, per-cell integer fields
    fd = q["feature_dump"]
    assert fd[0] == ["CB", "CorrectedReads", "MappedReads", "DeduplicatedReads", "MappingRate", "DedupRate", "MeanByMax", "NumGenesExpressed", "NumGenesOverMean"]
    nrec = np.diff(b.cell_rec_offsets.astype(np.int64))
    for c in range(120):
        row = fd[1 + c]
        assert row[0] == q["rows"][c] and int(row[1]) == nrec[c] and int(row[2]) == nrec[c]
        assert int(row[7]) == want.num_expr[c] and int(row[8]) == want.num_over_mean[c]
        if not res.endswith("-em"):
            assert row[3] == fmt_f32(want.sum_umi[c]) and row[4] == "1"
            assert row[5] == fmt_f32(np.float32(want.sum_umi[c]) / np.float32(nrec[c]))
    m = q["meta"]
    assert m["num_quantified_cells"] == 120 and m["num_genes"] == spec.num_rows and m["usa_mode"] == usa
    assert m["quant_options"]["small_thresh"] == 100 and m["num_tiny_cell_resolved"] == len(m["tiny_cell_resolved_cell_numbers"])
    assert m["resolution_strategy"] == {"cr-like": "CellRangerLike", "parsimony": "Parsimony", "cr-like-em": "CellRangerLikeEm", "trivial": "Trivial"}[res]


@pytest.mark.gpu
def test_quant_subset_and_multi_batch(tmp_path):
    spec = synth.SynthSpec(n_genes=300, reads_mean=200.0)
    b, bcs, d, t2g_path = make_input(tmp_path, spec, 90)
    keep = [0, 7, 8, 55, 89]
    open(tmp_path / "subset.txt", "w").write("".join(host.decode_barcode(bcs[i], 16) + "\n" for i in keep))
    out = str(tmp_path / "out")
    host.quantify(d, t2g_path, out, "cr-like", filter_list=str(tmp_path / "subset.txt"), batch_records=2000)
    q = host.load_quant_dir(out)
    assert q["rows"] == [host.decode_barcode(bcs[i], 16) for i in keep] and q["dims"][0] == 5
    out2 = str(tmp_path / "out2")
    host.quantify(d, t2g_path, out2, "cr-like", batch_records=2000)   # many small device batches
    out3 = str(tmp_path / "out3")
    host.quantify(d, t2g_path, out3, "cr-like")                       # one batch
    for fn in ("alevin/quants_mat.mtx", "alevin/quants_mat_rows.txt", "featureDump.txt"):
        assert open(os.path.join(out2, fn), "rb").read() == open(os.path.join(out3, fn), "rb").read()


# ---- snappy-framed input (map.collated.rad.sz, `collate --compress`; src/quant.rs:373-395) ----------
def test_snappy_framed_decoder_roundtrip_and_errors():
    import snappy_ref
    rng = np.random.default_rng(7)
    assert snappy_ref.crc32c(b"123456789") == 0xE3069283                      # CRC-32C check value
    payloads = [b"", b"a", b"abcd" * 5000, bytes(rng.integers(0, 256, 70000, dtype=np.uint8)),
                bytes(rng.integers(0, 4, 200000, dtype=np.uint8)),              # long matches, overlapping copies
                b"".join(int(x).to_bytes(4, "little") for x in rng.integers(0, 50, 40000))]
    for i, data in enumerate(payloads):
        for kw in ({}, {"block": 4096, "store_every": 3, "padding": True}, {"force4": True, "block": 30000}):
            z = snappy_ref.frame(data, **kw)
            assert host.snappy_framed_decompress(z, n_threads=1 + i % 4) == data, (i, kw)
    z = bytearray(snappy_ref.frame(b"hello world, hello world, hello world" * 100))
    z[-1] ^= 0x40
    with pytest.raises(RuntimeError):
        host.snappy_framed_decompress(bytes(z))
    good = snappy_ref.frame(b"x" * 1000)
    for bad in (good[:-3], good[10:], b"\x02\x01\x00\x00z" + good):
        with pytest.raises(RuntimeError):
            host.snappy_framed_decompress(bad)


@pytest.mark.gpu
@pytest.mark.parametrize("res", ["cr-like", "parsimony"])
def test_cli_reads_snappy_compressed_collated_rad(tmp_path, res):
    import json
    import shutil
    import snappy_ref
    spec = synth.SynthSpec(reads_mean=300.0, n_genes=500)
    b, bcs, d, t2g = make_input(tmp_path, spec, 40)
    host.quantify(d, t2g, str(tmp_path / "plain"), res)
    dz = tmp_path / "inz"
    shutil.copytree(d, dz)
    raw = open(dz / "map.collated.rad", "rb").read()
    os.remove(dz / "map.collated.rad")
    open(dz / "map.collated.rad.sz", "wb").write(snappy_ref.frame(raw, store_every=5, padding=True))
    json.dump({"compressed_output": True}, open(dz / "collate.json", "w"))
    host.quantify(str(dz), t2g, str(tmp_path / "fromz"), res)
    for f in ("alevin/quants_mat.mtx", "alevin/quants_mat_rows.txt", "alevin/quants_mat_cols.txt", "featureDump.txt"):
        assert open(tmp_path / "plain" / f, "rb").read() == open(tmp_path / "fromz" / f, "rb").read(), f


# ---- the RAD reader against files written from the REFERENCE's byte-level statements ---------
def _fixture_cells(rng, n_cells, n_refs, bc_len, umi_len, min_recs=3, max_recs=40):
    cells = []
    for c in range(n_cells):
        bc = int(rng.integers(0, 1 << min(2 * bc_len, 62)))
        recs = []
        for _ in range(int(rng.integers(min_recs, max_recs))):
            na = int(rng.integers(1, 5))
            refs = sorted(set(int(x) for x in rng.integers(0, n_refs, size=na)))
            recs.append((int(rng.integers(0, 1 << min(2 * umi_len, 62))), refs, [bool(x) for x in rng.integers(0, 2, size=len(refs))]))
        cells.append((bc, recs))
    return cells


def _expect(cells):
    import rad_fixture
    return (rad_fixture.fnv([bc for bc, recs in cells for _ in recs]), rad_fixture.fnv([r[0] for _, recs in cells for r in recs]),
            rad_fixture.fnv([x for _, recs in cells for r in recs for x in r[1]]),
            sum(len(recs) for _, recs in cells), sum(len(r[1]) for _, recs in cells for r in recs))


@pytest.mark.parametrize("bc_len,umi_len", [(16, 12), (16, 10), (8, 4), (24, 12), (14, 16), (3, 7)])
def test_rad_reader_on_reference_layout_fixture(tmp_path, bc_len, umi_len):
    # the plain layout `alevin-fry convert` writes (src/convert.rs:280-383): widths follow the barcode / UMI lengths
    import rad_fixture
    rng = np.random.default_rng(bc_len * 100 + umi_len)
    names = [f"tx{i}" for i in range(37)]
    cells = _fixture_cells(rng, 9, len(names), bc_len, umi_len)
    p = str(tmp_path / "map.collated.rad")
    rad_fixture.write_collated_rad(p, names, cells, bc_len, umi_len)
    info = host.rad_summary(p)
    hb, hu, hr, nrec, naln = _expect(cells)
    size = {"u8": 1, "u16": 2, "u32": 4, "u64": 8}
    assert (info.n_refs, info.num_chunks, info.n_records, info.n_alignments) == (37, 9, nrec, naln)
    assert (info.bc_len, info.umi_len) == (bc_len, umi_len)
    assert info.bc_size == size[rad_fixture.width_type(bc_len)] and info.umi_size == size[rad_fixture.width_type(umi_len)]
    assert (info.bc_off, info.umi_off, info.read_bytes, info.aln_bytes, info.refid_off) == (0, info.bc_size, info.bc_size + info.umi_size, 4, 0)
    assert (info.sum_bc, info.sum_umi, info.sum_refs) == (hb, hu, hr)


@pytest.mark.parametrize("layout", ["plain-pack24", "plain-u32", "extra-tags", "wide-umi"])
def test_product_parser_against_the_reference_walk(tmp_path, layout, monkeypatch):
    # the quantifier's own parallel parser (parse_batch: the 24-bit fast path with 4-byte stores and unconditional short-record
    # copies, and the general path) run without a GPU by afqh_host_stage_bench, against the independent record walk of
    # afqh_rad_summary and the values the fixture was written from: records of 1..4, of 5..9 and of 300 alignments (beyond the
    # 1-byte alignment counts), chunks that end in short and in long records
    import rad_fixture
    rng = np.random.default_rng(17)
    n_refs = 5000
    names = [f"t{i}" for i in range(n_refs)]
    umi_len = 16 if layout == "wide-umi" else 12
    cells = _fixture_cells(rng, 40, n_refs, 16, umi_len, min_recs=1, max_recs=60)
    for c in (3, 11, 39):       # long records, one of them the last of its chunk, one in the last chunk of the file
        bc, recs = cells[c]
        for na in (7, 300, 9):
            refs = sorted(set(int(x) for x in rng.integers(0, n_refs, size=na)))
            recs.insert(len(recs) if na == 9 else 1, (int(rng.integers(0, 1 << 20)), refs, [True] * len(refs)))
    p = str(tmp_path / "map.collated.rad")
    kw = {}
    if layout == "extra-tags":
        kw = dict(extra_read_tags=[(("frag_q", "u8"), 7)], extra_aln_tags=[(("pos", "u32"), 5)], extra_first=True)
    rad_fixture.write_collated_rad(p, names, cells, 16, umi_len, **kw)
    if layout == "plain-u32":
        monkeypatch.setenv("AFQ_NO_PACK24", "1")
    hb, hu, hr, nrec, naln = _expect(cells)
    ref = host.rad_summary(p)
    got = host.host_stage_bench(p, 4, 3)
    assert got.pack24 == (1 if layout in ("plain-pack24", "extra-tags") else 0)
    assert (got.n_cells, got.n_records, got.n_alignments) == (40, nrec, naln) == (ref.num_chunks, ref.n_records, ref.n_alignments)
    assert (got.sum_umi, got.sum_refs) == (hu, hr) == (ref.sum_umi, ref.sum_refs)
    assert got.sum_na == rad_fixture.fnv([len(r[1]) for _, recs in cells for r in recs])
    assert got.nnz > 0 and got.mtx_bytes > 0


def test_matrix_text_formatter_matches_printf(tmp_path):
    # the fast "row col val" formatter of the matrix body (two-digit table, integer path for whole numbers, shortest
    # round-trip digits otherwise) against Python's own formatting of the same synthetic result
    spec = synth.SynthSpec(n_genes=50, fixed_reads=40)
    b, bcs, d, _ = make_input(tmp_path, spec, 30)
    a = host.host_stage_bench(os.path.join(d, "map.collated.rad"), 2, 0)
    c = host.host_stage_bench(os.path.join(d, "map.collated.rad"), 4, 0)
    assert (a.nnz, a.mtx_bytes, a.mtx_sum) == (c.nnz, c.mtx_bytes, c.mtx_sum)      # independent of the thread count
    # re-create the synthetic result of afqh_host_stage_bench and format it here (every 3rd value a non-integer: Rust's `{}`
    # prints the shortest digits that round-trip the f32, which is what numpy's repr of a float32 gives)
    n_cols = spec.num_refs
    for frac_every in (0, 3):
        got = host.host_stage_bench(os.path.join(d, "map.collated.rad"), 2, frac_every)
        lines, e = [], 0
        for cell in range(30):
            nrec = int(b.cell_rec_offsets[cell + 1] - b.cell_rec_offsets[cell])
            k = min(nrec // 8 + 1, n_cols)
            stride = max(1, n_cols // k)
            for j in range(k):
                v = np.float32(1 + (((e * 2654435761) & 0xFFFFFFFFFFFFFFFF) >> 7) % 7)
                if frac_every and e % frac_every == 0:
                    v = np.float32(v + np.float32(1.0) / np.float32(2 + e % 5))
                text = str(int(v)) if float(v) == int(v) else np.format_float_positional(v, unique=True, trim="-")
                lines.append(f"{cell + 1} {j * stride + 1} {text}\n")
                e += 1
        txt = "".join(lines).encode()
        h = 0xCBF29CE484222325
        for ch in txt:
            h = ((h ^ ch) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
        assert (got.nnz, got.mtx_bytes, got.mtx_sum) == (len(lines), len(txt), h), frac_every


def test_rad_reader_skips_extra_tags_of_every_type(tmp_path):
    # tags a mapper may add around the standard ones: string / array / float file tags (e.g. `known_rad_type`,
    # tests/multi_barcode_integration.rs:72-75), extra read-level and alignment-level tags before or after b / u / refid
    import rad_fixture
    rng = np.random.default_rng(99)
    names = [f"gene_{i}" for i in range(12)]
    cells = _fixture_cells(rng, 5, len(names), 16, 12)
    hb, hu, hr, nrec, naln = _expect(cells)
    extra_file = [(("known_rad_type", "string"), "sc_rna_basic"), (("ref_lengths", "array", "u32", "u32"), list(range(12))),
                  (("frac", "f32"), 0.25), (("scale", "f64"), 2.5), (("flag", "bool"), 1), (("big", "u64"), 1 << 40)]
    for first in (False, True):
        p = str(tmp_path / f"extra_{first}.rad")
        rad_fixture.write_collated_rad(p, names, cells, 16, 12, extra_file_tags=extra_file,
                                       extra_read_tags=[(("frag_q", "u8"), 7), (("score", "f32"), 1.5)],
                                       extra_aln_tags=[(("pos", "u32"), 123456), (("mapq", "u16"), 60)], extra_first=first)
        info = host.rad_summary(p)
        assert (info.n_file_tags, info.n_read_tags, info.n_aln_tags) == (8, 4, 3)
        assert (info.n_records, info.n_alignments, info.bc_len, info.umi_len) == (nrec, naln, 16, 12)
        assert (info.sum_bc, info.sum_umi, info.sum_refs) == (hb, hu, hr)
        assert (info.bc_off, info.umi_off, info.refid_off) == ((5, 9, 6) if first else (0, 4, 0))
        assert (info.read_bytes, info.aln_bytes) == (13, 10)
    # this repo's own writer and the reference-derived writer agree byte for byte on the plain layout
    spec = synth.SynthSpec(n_genes=4, fixed_reads=5)
    b, bcs, d, _ = make_input(tmp_path, spec, 3)
    cells2 = []
    for c in range(3):
        r0, r1 = int(b.cell_rec_offsets[c]), int(b.cell_rec_offsets[c + 1])
        cells2.append((int(bcs[c]), [(int(b.rec_umi32[r]), [int(x) for x in b.refs[b.rec_ref_offsets[r]:b.rec_ref_offsets[r + 1]]]) for r in range(r0, r1)]))
    p2 = str(tmp_path / "ref_layout.rad")
    rad_fixture.write_collated_rad(p2, [f"t{i}" for i in range(spec.num_refs)], cells2, 16, spec.umi_len)
    assert open(p2, "rb").read() == open(os.path.join(d, "map.collated.rad"), "rb").read()


def test_rad_reader_rejects_corrupt_files(tmp_path):
    import rad_fixture
    rng = np.random.default_rng(5)
    names = [f"t{i}" for i in range(6)]
    cells = _fixture_cells(rng, 4, 6, 16, 12)
    p = str(tmp_path / "ok.rad")
    n = rad_fixture.write_collated_rad(p, names, cells)
    raw = open(p, "rb").read()
    for cut in (n - 1, n - 9, 40, 3):
        open(tmp_path / "cut.rad", "wb").write(raw[:cut])
        with pytest.raises(RuntimeError):
            host.rad_summary(str(tmp_path / "cut.rad"))


@pytest.mark.gpu
@pytest.mark.parametrize("res", ["cr-like", "parsimony"])
def test_cli_on_reference_layout_fixture_with_extra_tags(tmp_path, res):
    # end to end from a file written by tests/rad_fixture.py (NOT by this repo's writer), with extra tags of every kind
    # and no `ulen` tag in the second pass (the reference's quant never reads it)
    import json
    import rad_fixture
    spec = synth.SynthSpec(n_genes=200, reads_mean=250.0)
    b = synth.generate(spec, 0, 60)
    bcs = host.make_barcodes(0, 60)
    names = host.write_synth_t2g(str(tmp_path / "t2g.tsv"), spec)
    cells = []
    for c in range(60):
        r0, r1 = int(b.cell_rec_offsets[c]), int(b.cell_rec_offsets[c + 1])
        cells.append((int(bcs[c]), [(int(b.rec_umi32[r]), [int(x) for x in b.refs[b.rec_ref_offsets[r]:b.rec_ref_offsets[r + 1]]]) for r in range(r0, r1)]))
    t2g = synth.tid_to_gid(spec)
    o = QuantOpts(resolution=res, num_gene_ids=spec.num_gene_ids, num_rows=spec.num_rows, large_graph_thresh=1000 if res == "parsimony" else 0)
    want = oracle_lib.oracle_quant(o, t2g, b)
    for k, with_ulen in enumerate((True, False)):
        d = tmp_path / f"in{k}"
        os.makedirs(d)
        rad_fixture.write_collated_rad(str(d / "map.collated.rad"), names, cells, 16, 12, with_ulen=with_ulen, extra_first=bool(k),
                                       extra_file_tags=[(("known_rad_type", "string"), "sc_rna_basic"), (("lens", "array", "u16", "u32"), [1, 2, 3])],
                                       extra_read_tags=[(("q", "u8"), 3)], extra_aln_tags=[(("pos", "u32"), 77)])
        json.dump({"compressed_output": False}, open(d / "collate.json", "w"))
        json.dump({"velo_mode": False}, open(d / "generate_permit_list.json", "w"))
        out = str(tmp_path / f"out{k}")
        host.quantify(str(d), str(tmp_path / "t2g.tsv"), out, res)
        q = host.load_quant_dir(out)
        assert q["dims"] == (60, spec.num_rows, want.nnz)
        assert np.array_equal(np.array([t[1] for t in q["triplets"]], dtype=np.uint32), want.col)
        assert [t[2] for t in q["triplets"]] == [fmt_f32(v) for v in want.val]


@pytest.mark.gpu
def test_unmapped_counts_file_feeds_feature_dump(tmp_path):
    # unmapped_bc_count_collated.bin (src/quant.rs:1484-1494, 1181-1196): CorrectedReads = mapped + unmapped, MappingRate
    import struct
    spec = synth.SynthSpec(n_genes=100, reads_mean=150.0)
    b, bcs, d, t2g_path = make_input(tmp_path, spec, 30)
    unm = {int(bcs[i]): 10 * (i + 1) for i in range(0, 30, 3)}
    with open(os.path.join(d, "unmapped_bc_count_collated.bin"), "wb") as f:      # bincode HashMap<u64, u32> (src/atac/collate.rs:270-283)
        f.write(struct.pack("<Q", len(unm)))
        for k, v in unm.items():
            f.write(struct.pack("<QI", k, v))
    out = str(tmp_path / "out")
    host.quantify(d, t2g_path, out, "cr-like")
    fd = host.load_quant_dir(out)["feature_dump"]
    nrec = np.diff(b.cell_rec_offsets.astype(np.int64))
    for c in range(30):
        u = unm.get(int(bcs[c]), 0)
        assert int(fd[1 + c][1]) == nrec[c] + u and int(fd[1 + c][2]) == nrec[c]
        assert fd[1 + c][4] == fmt_f32(np.float32(nrec[c]) / np.float32(nrec[c] + u))


@pytest.mark.gpu
@pytest.mark.parametrize("res", ["cr-like", "parsimony-em"])
def test_device_list_one_reader_many_contexts_one_matrix(tmp_path, res):
    # src/quant.rs:1567-1575, 1811-1847: one reader feeds N workers and ONE matrix comes out. Batches go round-robin
    # to one context per listed GPU; the files must be byte-identical to the single-context run whatever the device
    # count (an ordinal may repeat, so a single-GPU box runs three contexts on GPU 0; `all` = every visible GPU)
    spec = synth.SynthSpec(n_genes=300, reads_mean=220.0)
    b, bcs, d, t2g_path = make_input(tmp_path, spec, 150)
    host.quantify(d, t2g_path, str(tmp_path / "one"), res, batch_records=3000)
    host.quantify(d, t2g_path, str(tmp_path / "three"), res, batch_records=3000, devices="0,0,0")
    host.quantify(d, t2g_path, str(tmp_path / "all"), res, batch_records=3000, devices="all")
    for fn in ("alevin/quants_mat.mtx", "alevin/quants_mat_rows.txt", "alevin/quants_mat_cols.txt", "featureDump.txt"):
        ref = open(os.path.join(tmp_path / "one", fn), "rb").read()
        assert open(os.path.join(tmp_path / "three", fn), "rb").read() == ref, fn
        assert open(os.path.join(tmp_path / "all", fn), "rb").read() == ref, fn
    with pytest.raises(RuntimeError):
        host.quantify(d, t2g_path, str(tmp_path / "bad"), res, devices="0,x")


# ---- --dump-eqclasses + `infer` (src/quant.rs:218-355, 1282-1307; src/infer.rs) -----------------------
def _read_dump(out):
    import gzip
    lines = gzip.open(os.path.join(out, "alevin", "gene_eqclass.txt.gz"), "rt").read().split("\n")[:-1]
    num_genes, num_eqc = int(lines[0]), int(lines[1])
    classes = {}
    for l in lines[2:]:
        v = [int(x) for x in l.split()]
        classes[v[-1]] = tuple(v[:-1])
    m = open(os.path.join(out, "alevin", "geqc_counts.mtx")).read().split("\n")
    assert m[0] == "%%MatrixMarket matrix coordinate real general"       # tests/infer_matrix_market.rs:52-63
    body = [l for l in m if l and not l.startswith("%")]
    dims = tuple(int(x) for x in body[0].split())
    trips = [(int(a) - 1, int(b) - 1, int(c)) for a, b, c in (l.split() for l in body[1:])]
    return num_genes, num_eqc, classes, dims, trips


@pytest.mark.gpu
@pytest.mark.parametrize("res,usa", [("cr-like", False), ("parsimony-em", False), ("cr-like-em", True), ("parsimony", True)])
def test_cli_dump_eqclasses_and_infer(tmp_path, res, usa):
    spec = synth.SynthSpec(n_genes=400, reads_mean=260.0, usa_mode=usa)
    b, bcs, d, t2g_path = make_input(tmp_path, spec, 80)
    out = str(tmp_path / "out")
    r = subprocess.run([host.CLI_PATH, "quant", "-i", d, "-m", t2g_path, "-o", out, "-r", res, "-t", "4", "-d"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    t2g = synth.tid_to_gid(spec)
    o = QuantOpts(resolution=res, usa_mode=usa, num_gene_ids=spec.num_gene_ids, num_rows=spec.num_rows, dump_eq=True,
                  large_graph_thresh=1000 if res.startswith("parsimony") else 0)
    want, wd = oracle_lib.oracle_quant_with_classes(o, t2g, b)
    num_genes, num_eqc, classes, dims, trips = _read_dump(out)
    assert num_genes == spec.num_rows and len(classes) == num_eqc and dims == (80, num_eqc, len(trips))
    assert host.load_quant_dir(out)["meta"]["dump_eq"] is True
    # every cell's classes: the file's (label -> count) multiset equals the oracle's gene_eqc, after the USA id remap
    G = spec.n_genes

    def remap(lab):   # S -> k, U -> G + k, adjacent S, U of one gene -> 2G + k (src/quant.rs:288-335)
        if not usa:
            return tuple(lab)
        res_, i = [], 0
        while i < len(lab):
            if i + 1 < len(lab) and (lab[i] | 1) == (lab[i + 1] | 1):
                res_.append((lab[i] >> 1) + 2 * G); i += 2
            else:
                res_.append((lab[i] >> 1) if lab[i] % 2 == 0 else (lab[i] >> 1) + G); i += 1
        return tuple(res_)
    per_cell = {}
    for r_, c_, v_ in trips:
        per_cell.setdefault(r_, []).append((classes[c_], v_))
    for c in range(80):
        exp = sorted((remap(lab), cnt) for lab, cnt in wd.cell(c))
        assert sorted(per_cell.get(c, [])) == exp, c
    # `infer` on the dump: the same EM the oracle runs on those rows (informative init), written as a matrix
    out2 = str(tmp_path / "inferred")
    r = subprocess.run([host.CLI_PATH, "infer", "-c", os.path.join(out, "alevin", "geqc_counts.mtx"),
                        "-e", os.path.join(out, "alevin", "gene_eqclass.txt.gz"), "-o", out2] + (["--usa"] if usa else []), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(os.path.join(out2, "quants_mat_rows.txt")).read() == open(os.path.join(out, "alevin", "quants_mat_rows.txt")).read()
    assert open(os.path.join(out2, "quants_mat_cols.txt")).read() == open(os.path.join(out, "alevin", "quants_mat_cols.txt")).read()
    lo, lb = [0], []
    for k in range(num_eqc):
        lb.extend(classes[k]); lo.append(len(lb))
    co, ce, cc = [0], [], []
    for c in range(80):
        row = sorted((c_, v_) for r_, c_, v_ in trips if r_ == c)
        ce.extend(x[0] for x in row); cc.extend(x[1] for x in row); co.append(len(ce))
    exp = oracle_lib.infer_cells(spec.num_rows, usa, False, np.array(lo, dtype=np.uint32), np.array(lb, dtype=np.uint32),
                                 np.array(co, dtype=np.uint64), np.array(ce, dtype=np.uint32), np.array(cc, dtype=np.uint32))
    m = open(os.path.join(out2, "quants_mat.mtx")).read().split("\n")
    body = [l for l in m if l and not l.startswith("%")]
    assert tuple(int(x) for x in body[0].split()) == (80, spec.num_rows, exp.nnz)
    got_cols = np.array([int(l.split()[1]) - 1 for l in body[1:]], dtype=np.uint32)
    got_vals = np.array([float(l.split()[2]) for l in body[1:]], dtype=np.float32)
    assert np.array_equal(got_cols, exp.col)
    np.testing.assert_allclose(got_vals, exp.val, rtol=1e-5)
    # --quant-subset keeps the listed barcodes only, in matrix order
    keep = [3, 17, 42]
    open(tmp_path / "subset.txt", "w").write("".join(host.decode_barcode(bcs[i], 16) + "\n" for i in keep))
    out3 = str(tmp_path / "inferred_subset")
    host.infer(os.path.join(out, "alevin", "geqc_counts.mtx"), os.path.join(out, "alevin", "gene_eqclass.txt.gz"), out3, usa_mode=usa,
               filter_list=str(tmp_path / "subset.txt"))
    assert open(os.path.join(out3, "quants_mat_rows.txt")).read().split("\n")[:-1] == [host.decode_barcode(bcs[i], 16) for i in keep]


def test_infer_cli_argument_validation(tmp_path):
    r = subprocess.run([host.CLI_PATH, "infer", "-c", "x"], capture_output=True, text=True)
    assert r.returncode == 2 and "required arguments" in r.stderr
    r = subprocess.run([host.CLI_PATH, "infer", "-c", "x", "-e", "y", "-o", "z", "--use-eds"], capture_output=True, text=True)
    assert r.returncode == 1 and "--use-eds is no longer supported" in r.stderr
    with pytest.raises(RuntimeError):
        host.infer(str(tmp_path / "missing.mtx"), str(tmp_path / "missing.gz"), str(tmp_path / "o"))
