"""Deterministic synthetic collated-cell generator (bench/test infrastructure — deliberately OUTSIDE the
product package alevin_fry_b200/).

ctypes wrapper over synth/libafq_synth.so; see synth/afq_synth.cpp for the model and
SURVEY.md §8(d) for the named configurations C1-C5.
"""
import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from alevin_fry_b200._abi import REPO_ROOT
from alevin_fry_b200.quant import CellBatch

SYNTH_LIB_PATH = os.path.join(REPO_ROOT, "synth", "libafq_synth.so")
GLOBAL_SEED = 20260925


class _Spec(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_genes", C.c_uint32), ("usa_mode", C.c_int32), ("umi_len", C.c_int32),
                ("fixed_reads", C.c_int32), ("reads_mean", C.c_double), ("lognorm_sigma", C.c_double),
                ("reads_per_umi", C.c_double), ("zipf_s", C.c_double), ("p_multi2", C.c_double),
                ("p_multi3", C.c_double), ("umi_err", C.c_double)]


@dataclass
class SynthSpec:
    seed: int = GLOBAL_SEED
    n_genes: int = 30000
    usa_mode: bool = False
    umi_len: int = 12
    fixed_reads: int = 0
    reads_mean: float = 2000.0
    lognorm_sigma: float = 0.8
    reads_per_umi: float = 4.0
    zipf_s: float = 1.05
    p_multi2: float = 0.10
    p_multi3: float = 0.05
    umi_err: float = 0.01

    def to_c(self):
        return _Spec(self.seed, self.n_genes, int(self.usa_mode), self.umi_len, self.fixed_reads, self.reads_mean,
                     self.lognorm_sigma, self.reads_per_umi, self.zipf_s, self.p_multi2, self.p_multi3, self.umi_err)

    @property
    def num_refs(self): return self.n_genes * (4 if self.usa_mode else 3)
    @property
    def num_gene_ids(self): return self.n_genes * (2 if self.usa_mode else 1)
    @property
    def num_rows(self): return self.n_genes * (3 if self.usa_mode else 1)


# SURVEY.md §8(d) configurations (cells, resolution) -> spec
def config_spec(name: str) -> SynthSpec:
    name = name.upper()
    if name == "C1":
        return SynthSpec(fixed_reads=50, reads_per_umi=2.0)
    if name in ("C2", "C3"):
        return SynthSpec()
    if name == "C4":
        return SynthSpec(usa_mode=True)
    if name == "C5":
        return SynthSpec(n_genes=5000, reads_per_umi=40.0, p_multi2=0.20, p_multi3=0.10, umi_err=0.02)
    raise ValueError(name)


_lib = None


def _synth_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SYNTH_LIB_PATH):
            raise FileNotFoundError(f"{SYNTH_LIB_PATH} missing: run `make` or __graft_entry__.build()")
        l = C.CDLL(SYNTH_LIB_PATH)
        l.afq_synth_t2g.argtypes = [C.POINTER(_Spec), C.c_void_p]
        l.afq_synth_t2g.restype = None
        l.afq_synth_sizes.argtypes = [C.POINTER(_Spec), C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
        l.afq_synth_sizes.restype = None
        l.afq_synth_fill.argtypes = [C.POINTER(_Spec), C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int]
        l.afq_synth_fill.restype = None
        _lib = l
    return _lib


def tid_to_gid(spec: SynthSpec) -> np.ndarray:
    t = np.empty(spec.num_refs, dtype=np.uint32)
    cs = spec.to_c()
    _synth_lib().afq_synth_t2g(C.byref(cs), t.ctypes.data_as(C.c_void_p))
    return t


def sizes(spec: SynthSpec, first_cell: int, n_cells: int, n_threads: int = 0):
    """(n_records, n_refs_total) of cells [first_cell, first_cell+n_cells) without generating them."""
    if n_threads <= 0:
        n_threads = os.cpu_count() or 1
    cs = spec.to_c()
    nrec = np.zeros(n_cells, dtype=np.uint64)
    nref = np.zeros(n_cells, dtype=np.uint64)
    _synth_lib().afq_synth_sizes(C.byref(cs), first_cell, n_cells, nrec.ctypes.data_as(C.c_void_p),
                                 nref.ctypes.data_as(C.c_void_p), n_threads)
    return int(nrec.sum()), int(nref.sum())


def generate(spec: SynthSpec, first_cell: int, n_cells: int, n_threads: int = 0, alloc=None) -> CellBatch:
    """Generate cells [first_cell, first_cell+n_cells). `alloc(nbytes_dtype_tuple)` may supply
    pinned arrays: alloc(n, dtype) -> np.ndarray."""
    if n_threads <= 0:
        n_threads = os.cpu_count() or 1
    l = _synth_lib()
    cs = spec.to_c()
    nrec = np.zeros(n_cells, dtype=np.uint64)
    nref = np.zeros(n_cells, dtype=np.uint64)
    l.afq_synth_sizes(C.byref(cs), first_cell, n_cells, nrec.ctypes.data_as(C.c_void_p),
                      nref.ctypes.data_as(C.c_void_p), n_threads)
    cro = np.zeros(n_cells + 1, dtype=np.uint64)
    np.cumsum(nrec, out=cro[1:])
    cfo = np.zeros(n_cells + 1, dtype=np.uint64)
    np.cumsum(nref, out=cfo[1:])
    n_rec, n_ref = int(cro[-1]), int(cfo[-1])
    if n_ref >= 2 ** 32 - 16:
        raise ValueError("batch too large for u32 ref offsets; generate fewer cells per batch")
    mk = alloc if alloc is not None else (lambda n, dt: np.empty(n, dtype=dt))
    umi = mk(n_rec, np.uint32)
    roff = mk(n_rec + 1, np.uint32)
    refs = mk(max(n_ref, 1), np.uint32)[:n_ref]
    l.afq_synth_fill(C.byref(cs), first_cell, n_cells, cro.ctypes.data_as(C.c_void_p), cfo.ctypes.data_as(C.c_void_p),
                     umi.ctypes.data_as(C.c_void_p), roff.ctypes.data_as(C.c_void_p),
                     refs.ctypes.data_as(C.c_void_p), n_threads)
    b = CellBatch.__new__(CellBatch)
    b.cell_rec_offsets, b.rec_umi32, b.rec_ref_offsets, b.refs, b.first_cell_index = cro, umi, roff, refs, first_cell
    return b
