"""Python access to the C++ host side (libafq_host.so): the `quantify(QuantOpts)` drop-in,
a collated-RAD writer for synthetic inputs, and a loader for quant output directories."""
import ctypes as C
import json
import os

import numpy as np

from ._abi import REPO_ROOT

HOST_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libafq_host.so")
CLI_PATH = os.path.join(REPO_ROOT, "bin", "alevin-fry")


class _Opts(C.Structure):
    _fields_ = [("input_dir", C.c_char_p), ("tg_map", C.c_char_p), ("output_dir", C.c_char_p),
                ("num_threads", C.c_uint32), ("num_bootstraps", C.c_uint32),
                ("init_uniform", C.c_int32), ("summary_stat", C.c_int32), ("dump_eq", C.c_int32),
                ("resolution", C.c_char_p), ("pug_exact_umi", C.c_int32), ("sa_model", C.c_char_p),
                ("small_thresh", C.c_uint64), ("large_graph_thresh", C.c_uint64), ("filter_list", C.c_char_p),
                ("cmdline", C.c_char_p), ("version", C.c_char_p), ("device", C.c_int32), ("batch_records", C.c_uint64),
                ("devices", C.c_char_p), ("process_exits", C.c_int32)]


class _InferOpts(C.Structure):
    _fields_ = [("count_mat", C.c_char_p), ("eq_labels", C.c_char_p), ("output_dir", C.c_char_p), ("usa_mode", C.c_int32),
                ("filter_list", C.c_char_p), ("num_threads", C.c_uint32), ("device", C.c_int32)]


class RadInfo(C.Structure):
    _fields_ = [("n_refs", C.c_uint64), ("num_chunks", C.c_uint64), ("n_records", C.c_uint64), ("n_alignments", C.c_uint64),
                ("sum_bc", C.c_uint64), ("sum_umi", C.c_uint64), ("sum_refs", C.c_uint64),
                ("bc_len", C.c_uint32), ("umi_len", C.c_uint32),
                ("read_bytes", C.c_uint32), ("aln_bytes", C.c_uint32), ("bc_size", C.c_uint32), ("umi_size", C.c_uint32),
                ("bc_off", C.c_uint32), ("umi_off", C.c_uint32), ("refid_off", C.c_uint32),
                ("n_file_tags", C.c_uint32), ("n_read_tags", C.c_uint32), ("n_aln_tags", C.c_uint32)]


class StageInfo(C.Structure):
    _fields_ = [("n_cells", C.c_uint64), ("n_records", C.c_uint64), ("n_alignments", C.c_uint64), ("nnz", C.c_uint64),
                ("mtx_bytes", C.c_uint64), ("mtx_sum", C.c_uint64), ("sum_umi", C.c_uint64), ("sum_refs", C.c_uint64), ("sum_na", C.c_uint64),
                ("walk_s", C.c_double), ("parse_s", C.c_double), ("format_s", C.c_double), ("parse_warm_s", C.c_double), ("threads", C.c_uint32), ("pack24", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise FileNotFoundError(f"{HOST_LIB_PATH} missing: run `make` or __graft_entry__.build()")
        l = C.CDLL(HOST_LIB_PATH)
        l.afqh_quantify.restype = C.c_int
        l.afqh_quantify.argtypes = [C.POINTER(_Opts), C.c_char_p, C.c_size_t]
        l.afqh_snappy_framed_decompress.restype = C.c_int
        l.afqh_snappy_framed_decompress.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_uint32,
                                                    C.c_char_p, C.c_size_t]
        l.afqh_infer.restype = C.c_int
        l.afqh_infer.argtypes = [C.POINTER(_InferOpts), C.c_char_p, C.c_size_t]
        l.afqh_rad_summary.restype = C.c_int
        l.afqh_rad_summary.argtypes = [C.c_char_p, C.POINTER(RadInfo), C.c_char_p, C.c_size_t]
        l.afqh_host_stage_bench.restype = C.c_int
        l.afqh_host_stage_bench.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.POINTER(StageInfo), C.c_char_p, C.c_size_t]
        l.afqh_free.restype = None
        l.afqh_free.argtypes = [C.c_void_p]
        l.afqh_write_collated_rad.restype = C.c_int
        l.afqh_write_collated_rad.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.POINTER(C.c_char_p), C.c_uint64, C.c_uint16, C.c_uint16,
                                              C.c_char_p, C.c_size_t]
        _lib = l
    return _lib


def quantify(input_dir, tg_map, output_dir, resolution, num_threads=2, small_thresh=100, large_graph_thresh=None,
             pug_exact_umi=False, init_uniform=False, filter_list=None, device=0, batch_records=0, cmdline="python",
             version="0.18.0-afq-b200", num_bootstraps=0, dump_eq=False, sa_model="winner-take-all", devices=None):
    """alevin_fry::quant::quantify(QuantOpts) (src/quant.rs:359). Raises RuntimeError on failure."""
    if large_graph_thresh is None:
        large_graph_thresh = 1000 if resolution.lower().startswith("parsimony") else 0
    o = _Opts(os.fsencode(input_dir), os.fsencode(tg_map), os.fsencode(output_dir), num_threads, num_bootstraps,
              int(init_uniform), 0, int(dump_eq), resolution.encode(), int(pug_exact_umi), sa_model.encode(),
              small_thresh, large_graph_thresh, os.fsencode(filter_list) if filter_list else None,
              cmdline.encode(), version.encode(), device, batch_records, devices.encode() if devices else None)
    err = C.create_string_buffer(2048)
    rc = lib().afqh_quantify(C.byref(o), err, 2048)
    if rc != 0:
        raise RuntimeError(err.value.decode(errors="replace"))


def infer(count_mat, eq_labels, output_dir, usa_mode=False, filter_list=None, num_threads=2, device=0):
    """alevin_fry::infer::infer (src/infer.rs:31). Raises RuntimeError on failure."""
    o = _InferOpts(os.fsencode(count_mat), os.fsencode(eq_labels), os.fsencode(output_dir), int(usa_mode),
                   os.fsencode(filter_list) if filter_list else None, num_threads, device)
    err = C.create_string_buffer(2048)
    if lib().afqh_infer(C.byref(o), err, 2048) != 0:
        raise RuntimeError(err.value.decode(errors="replace"))


def host_stage_bench(path, n_threads=0, frac_every=0) -> StageInfo:
    """The two host stages of `quant` without a GPU: chunk index + the product's parallel parser, and the parallel text
    formatting on a synthetic result (wall seconds per stage + checksums of the parsed arrays)."""
    info = StageInfo()
    err = C.create_string_buffer(1024)
    if lib().afqh_host_stage_bench(os.fsencode(path), n_threads or (os.cpu_count() or 2), frac_every, C.byref(info), err, 1024) != 0:
        raise RuntimeError(err.value.decode(errors="replace"))
    return info


def rad_summary(path) -> RadInfo:
    """CPU-only probe of a collated RAD file through the quantifier's own prelude / layout / chunk-walk code."""
    info = RadInfo()
    err = C.create_string_buffer(1024)
    if lib().afqh_rad_summary(os.fsencode(path), C.byref(info), err, 1024) != 0:
        raise RuntimeError(err.value.decode(errors="replace"))
    return info


def make_barcodes(first_cell, n_cells, bc_len=16):
    """Deterministic distinct barcodes (tests/multi_barcode_integration.rs:36-41 style mix)."""
    idx = np.arange(first_cell, first_cell + n_cells, dtype=np.uint64)
    mask = np.uint64((1 << (2 * bc_len)) - 1) if bc_len < 32 else np.uint64(0xFFFFFFFFFFFFFFFF)
    return (idx * np.uint64(2654435761)) & mask


def write_collated_rad(dirname, batch, barcodes, ref_names, bc_len=16, umi_len=12):
    names = (C.c_char_p * len(ref_names))(*[n.encode() for n in ref_names])
    bcs = np.ascontiguousarray(barcodes, dtype=np.uint64)
    err = C.create_string_buffer(1024)
    rc = lib().afqh_write_collated_rad(os.fsencode(dirname), batch.n_cells, batch.cell_rec_offsets.ctypes.data,
                                       bcs.ctypes.data, batch.rec_umi32.ctypes.data, batch.rec_ref_offsets.ctypes.data,
                                       batch.refs.ctypes.data, names, len(ref_names), bc_len, umi_len, err, 1024)
    if rc != 0:
        raise RuntimeError(err.value.decode())


def write_synth_t2g(path, spec):
    """t2g TSV for a SynthSpec: tx names t<i>, gene names g<k>; 3 columns in USA mode."""
    with open(path, "w") as f:
        if not spec.usa_mode:
            for g in range(spec.n_genes):
                for k in range(3):
                    f.write(f"t{3 * g + k}\tg{g}\n")
        else:
            for g in range(spec.n_genes):
                for k in range(3):
                    f.write(f"t{4 * g + k}\tg{g}\tS\n")
                f.write(f"t{4 * g + 3}\tg{g}\tU\n")
    return [f"t{i}" for i in range(spec.num_refs)]


def decode_barcode(bc, length):
    return "".join("ACGT"[(int(bc) >> (2 * (length - 1 - i))) & 3] for i in range(length))


def load_quant_dir(path):
    """Parse a quant output directory -> dict(rows, cols, triplets (r, c, v) 0-based, feature_dump, meta)."""
    rows = open(os.path.join(path, "alevin", "quants_mat_rows.txt")).read().split("\n")[:-1]
    cols = open(os.path.join(path, "alevin", "quants_mat_cols.txt")).read().split("\n")[:-1]
    lines = open(os.path.join(path, "alevin", "quants_mat.mtx")).read().split("\n")
    header = [l for l in lines if l.startswith("%")]
    body = [l for l in lines if l and not l.startswith("%")]
    dims = tuple(int(x) for x in body[0].split())
    trip = [(int(a) - 1, int(b) - 1, c) for a, b, c in (l.split() for l in body[1:])]
    fd = [l.split("\t") for l in open(os.path.join(path, "featureDump.txt")).read().split("\n")[:-1]]
    meta = json.load(open(os.path.join(path, "quant.json")))
    return dict(rows=rows, cols=cols, header=header, dims=dims, triplets=trip, feature_dump=fd, meta=meta)


def snappy_framed_decompress(data: bytes, n_threads: int = 2) -> bytes:
    """decode a Snappy framing-format stream (map.collated.rad.sz) with the host library's decoder"""
    out = C.c_void_p()
    n = C.c_size_t()
    err = C.create_string_buffer(512)
    rc = lib().afqh_snappy_framed_decompress(data, len(data), C.byref(out), C.byref(n), n_threads, err, 512)
    if rc != 0:
        raise RuntimeError("snappy: " + err.value.decode())
    try:
        return C.string_at(out.value, n.value)
    finally:
        lib().afqh_free(out)
