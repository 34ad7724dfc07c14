"""ctypes access to the CPU oracle (oracle/libafq_oracle.so). TEST INFRASTRUCTURE ONLY:
nothing under alevin_fry_b200/ imports this module."""
import ctypes as C
import os

import numpy as np

from alevin_fry_b200._abi import AfqBatch, AfqConfig, AfqResult, REPO_ROOT
from alevin_fry_b200.quant import CellBatch, QuantOpts, QuantResult

ORACLE_LIB_PATH = os.path.join(REPO_ROOT, "oracle", "libafq_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(ORACLE_LIB_PATH)
        l.afq_oracle_quant.restype = C.c_int
        l.afq_oracle_quant.argtypes = [C.POINTER(AfqConfig), C.c_void_p, C.c_uint64, C.POINTER(AfqBatch), C.c_int,
                                       C.POINTER(AfqResult), C.POINTER(C.c_void_p)]
        l.afq_oracle_quant_dump.restype = C.c_int
        l.afq_oracle_quant_dump.argtypes = [C.POINTER(AfqConfig), C.c_void_p, C.c_uint64, C.POINTER(AfqBatch), C.c_int,
                                            C.POINTER(AfqResult), C.POINTER(C.c_void_p), C.c_void_p]
        l.afq_oracle_release.restype = None
        l.afq_oracle_release.argtypes = [C.c_void_p]
        l.afq_oracle_em_subset.restype = C.c_int
        l.afq_oracle_em_subset.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                           C.c_int, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]
        l.afq_oracle_em_dense.restype = C.c_int
        l.afq_oracle_em_dense.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_uint32,
                                          C.c_int, C.c_void_p]
        l.afq_oracle_tie_census.restype = C.c_int
        l.afq_oracle_tie_census.argtypes = [C.POINTER(AfqConfig), C.c_void_p, C.c_uint64, C.POINTER(AfqBatch), C.c_int, C.c_void_p]
        l.afq_oracle_hamming.restype = C.c_int
        l.afq_oracle_hamming.argtypes = [C.c_uint64, C.c_uint64]
        _lib = l
    return _lib


def oracle_quant(opts: QuantOpts, tid_to_gid: np.ndarray, batch: CellBatch, n_threads: int = 0) -> QuantResult:
    if n_threads <= 0:
        n_threads = os.cpu_count() or 1
    t2g = np.ascontiguousarray(tid_to_gid, dtype=np.uint32)
    cfg = opts.to_c()
    cb = batch.to_c()
    r = AfqResult()
    h = C.c_void_p()
    rc = lib().afq_oracle_quant(C.byref(cfg), t2g.ctypes.data_as(C.c_void_p), len(t2g), C.byref(cb), n_threads,
                                C.byref(r), C.byref(h))
    assert rc == 0, rc
    out = QuantResult.from_c(r)
    lib().afq_oracle_release(h)
    return out


TIE_CENSUS_FIELDS = ("cells", "molecules", "components_multi", "components_tie_on_path", "components_label_sensitive",
                     "components_capped", "molecules_label_sensitive", "cells_label_sensitive", "components_count_sensitive",
                     "molecules_count_sensitive", "cells_count_sensitive", "molecules_tie_on_path", "molecules_changed_worst_case")


def tie_census(opts: QuantOpts, tid_to_gid: np.ndarray, batch: CellBatch, n_threads: int = 0) -> dict:
    """How much of a parsimony result can depend on the cover loop's start-vertex order (the reference follows
    hash-iteration order there, src/pugutils.rs:1090-1110): see tie_census_cell in oracle/afq_oracle.cpp."""
    if n_threads <= 0:
        n_threads = os.cpu_count() or 1
    t2g = np.ascontiguousarray(tid_to_gid, dtype=np.uint32)
    cfg = opts.to_c()
    cb = batch.to_c()
    out = np.zeros(len(TIE_CENSUS_FIELDS), dtype=np.uint64)
    rc = lib().afq_oracle_tie_census(C.byref(cfg), t2g.ctypes.data_as(C.c_void_p), len(t2g), C.byref(cb), n_threads, out.ctypes.data)
    assert rc == 0, rc
    d = {k: int(v) for k, v in zip(TIE_CENSUS_FIELDS, out)}
    mol = max(d["molecules"], 1)
    d["frac_molecules_label_sensitive"] = d["molecules_label_sensitive"] / mol
    d["frac_molecules_count_sensitive"] = d["molecules_count_sensitive"] / mol
    d["frac_molecules_changed_worst_case"] = d["molecules_changed_worst_case"] / mol
    d["frac_cells_count_sensitive"] = d["cells_count_sensitive"] / max(d["cells"], 1)
    return d


def oracle_quant_with_classes(opts: QuantOpts, tid_to_gid: np.ndarray, batch: CellBatch, n_threads: int = 0):
    """(QuantResult, EqcDump): the reference's per-cell gene_eqc maps (what --dump-eqclasses records), canonical order."""
    from alevin_fry_b200._abi import AfqEqcDump
    from alevin_fry_b200.quant import EqcDump
    if n_threads <= 0:
        n_threads = os.cpu_count() or 1
    t2g = np.ascontiguousarray(tid_to_gid, dtype=np.uint32)
    cfg = opts.to_c()
    cb = batch.to_c()
    r, d, h = AfqResult(), AfqEqcDump(), C.c_void_p()
    rc = lib().afq_oracle_quant_dump(C.byref(cfg), t2g.ctypes.data_as(C.c_void_p), len(t2g), C.byref(cb), n_threads, C.byref(r), C.byref(h), C.byref(d))
    assert rc == 0, rc
    out = QuantResult.from_c(r), EqcDump.from_c(d)
    lib().afq_oracle_release(h)
    return out


def _csr(classes):
    labels, starts = [], [0]
    for c in classes:
        labels.extend(c)
        starts.append(len(labels))
    return np.array(labels, dtype=np.uint32), np.array(starts, dtype=np.uint32)


def em_subset(classes, cell_data, init_uniform, num_alphas, only_unique=False, usa_offsets=None):
    """em_optimize_subset (src/em.rs:251-456) on explicit classes; cell_data = [(eq_id, count)]."""
    labels, starts = _csr(classes)
    eq = np.array([c[0] for c in cell_data], dtype=np.uint32)
    ct = np.array([c[1] for c in cell_data], dtype=np.uint32)
    out = np.zeros(num_alphas, dtype=np.float32)
    uo, ao = usa_offsets if usa_offsets else (0, 0)
    lib().afq_oracle_em_subset(labels.ctypes.data, starts.ctypes.data, len(classes), eq.ctypes.data, ct.ctypes.data,
                               len(cell_data), int(init_uniform), num_alphas, int(only_unique), uo, ao, out.ctypes.data)
    return out


def em_dense(classes, counts, init_uniform, num_alphas, only_unique=False):
    """em_optimize (src/em.rs:487-582) on explicit classes."""
    labels, starts = _csr(classes)
    ct = np.array(counts, dtype=np.uint32)
    out = np.zeros(num_alphas, dtype=np.float32)
    lib().afq_oracle_em_dense(labels.ctypes.data, starts.ctypes.data, len(classes), ct.ctypes.data, int(init_uniform),
                              num_alphas, int(only_unique), out.ctypes.data)
    return out


def infer_cells(num_alphas, usa, init_uniform, label_offsets, labels, cell_offsets, cell_eq, cell_cnt) -> QuantResult:
    """The oracle's em_optimize_subset (src/em.rs:251-456) over the rows of a count matrix: what `alevin-fry infer` computes."""
    lo = np.ascontiguousarray(label_offsets, dtype=np.uint32); lb = np.ascontiguousarray(labels, dtype=np.uint32)
    uo, ao = (num_alphas // 3, 2 * (num_alphas // 3)) if usa else (0, 0)
    rp, col, val, sm, mx, ne, nom, fl = [0], [], [], [], [], [], [], []
    for c in range(len(cell_offsets) - 1):
        a, b = int(cell_offsets[c]), int(cell_offsets[c + 1])
        eq = np.ascontiguousarray(cell_eq[a:b], dtype=np.uint32); ct = np.ascontiguousarray(cell_cnt[a:b], dtype=np.uint32)
        out = np.zeros(num_alphas, dtype=np.float32)
        if b > a:
            lib().afq_oracle_em_subset(lb.ctypes.data, lo.ctypes.data, len(lo) - 1, eq.ctypes.data, ct.ctypes.data, b - a,
                                       int(init_uniform), num_alphas, 0, uo, ao, out.ctypes.data)
        nz = np.nonzero(out > 0)[0]
        col.extend(nz.tolist()); val.extend(out[nz].tolist()); rp.append(len(col))
        sm.append(out.sum(dtype=np.float32)); mx.append(out.max() if len(out) else 0); ne.append(len(nz))
        mean = np.float32(sm[-1]) / np.float32(len(nz)) if len(nz) else np.float32("nan")
        nom.append(int((out[nz] > mean).sum())); fl.append(4 if len(nz) == 0 else 0)
    return QuantResult(np.array(rp, dtype=np.uint64), np.array(col, dtype=np.uint32), np.array(val, dtype=np.float32),
                       np.array(sm, dtype=np.float32), np.array(mx, dtype=np.float32), np.array(ne, dtype=np.uint32),
                       np.array(nom, dtype=np.uint32), np.array(fl, dtype=np.uint8))
