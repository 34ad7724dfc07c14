"""Python host side of the quant hot path: option struct, batch container, context.

`QuantOpts` mirrors the reference's QuantOpts (src/prog_opts.rs:24-43) for the fields that
reach the worker (WorkerConfig, src/quant.rs:398-416); names and defaults follow the
reference CLI (src/main.rs:294-348).
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _abi
from ._abi import AfqBatch, AfqConfig, AfqDeviceOut, AfqError, AfqResult


@dataclass
class QuantOpts:
    resolution: str = "cr-like"          # -r/--resolution (case-insensitive, src/main.rs:320)
    usa_mode: bool = False               # decided by the t2g file having 3 columns
    init_uniform: bool = False           # --init-uniform
    pug_exact_umi: bool = False          # --umi-edit-dist 0
    sa_model: str = "winner-take-all"    # --sa-model (hidden)
    small_thresh: int = 100              # --small-thresh
    large_graph_thresh: int = 1000       # --large-graph-thresh (hidden)
    num_gene_ids: int = 0                # G, or 2G in USA mode
    num_rows: int = 0                    # G, or 3G in USA mode
    barcode_len: int = 16
    umi_len: int = 12
    device: int = 0
    dump_eq: bool = False                # -d/--dump-eqclasses: results carry every cell's gene eq-classes

    def to_c(self) -> AfqConfig:
        res = self.resolution.lower()
        if res not in _abi.RESOLUTIONS:
            raise ValueError(f"unknown resolution {self.resolution!r}; expected one of {sorted(_abi.RESOLUTIONS)}")
        if self.sa_model not in ("winner-take-all", "prefer-ambig"):
            raise ValueError("sa_model must be winner-take-all or prefer-ambig")
        c = AfqConfig()
        c.resolution = _abi.RESOLUTIONS[res]
        c.usa_mode = int(self.usa_mode)
        c.em_init_uniform = int(self.init_uniform)
        c.pug_exact_umi = int(self.pug_exact_umi)
        c.sa_model = 0 if self.sa_model == "winner-take-all" else 1
        c.num_gene_ids = self.num_gene_ids
        c.num_rows = self.num_rows
        c.small_thresh = self.small_thresh
        c.large_graph_thresh = self.large_graph_thresh
        c.barcode_len = self.barcode_len
        c.umi_len = self.umi_len
        c.device = self.device
        c.dump_eq = int(self.dump_eq)
        return c


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


@dataclass
class CellBatch:
    """SoA batch of consecutive collated cells (host numpy arrays), see afq_batch."""
    cell_rec_offsets: np.ndarray   # u64 [n_cells+1]
    rec_umi32: np.ndarray          # u32 [n_records]
    rec_ref_offsets: np.ndarray    # u32 [n_records+1]
    refs: np.ndarray               # u32 [n_refs_total]
    first_cell_index: int = 0

    def __post_init__(self):
        self.cell_rec_offsets = np.ascontiguousarray(self.cell_rec_offsets, dtype=np.uint64)
        self.rec_umi32 = np.ascontiguousarray(self.rec_umi32, dtype=np.uint32)
        self.rec_ref_offsets = np.ascontiguousarray(self.rec_ref_offsets, dtype=np.uint32)
        self.refs = np.ascontiguousarray(self.refs, dtype=np.uint32)
        if len(self.cell_rec_offsets) < 1 or len(self.rec_ref_offsets) != len(self.rec_umi32) + 1:
            raise ValueError("inconsistent batch arrays")
        if int(self.cell_rec_offsets[-1]) != len(self.rec_umi32):
            raise ValueError("cell_rec_offsets[-1] must equal n_records")
        if int(self.rec_ref_offsets[-1]) != len(self.refs):
            raise ValueError("rec_ref_offsets[-1] must equal n_refs_total")

    @property
    def n_cells(self): return len(self.cell_rec_offsets) - 1
    @property
    def n_records(self): return len(self.rec_umi32)
    @property
    def n_refs_total(self): return len(self.refs)

    def na8(self) -> np.ndarray:
        """Per-record alignment counts as u8 (the compact alternative to rec_ref_offsets)."""
        cached = getattr(self, "_na8", None)
        if cached is None:
            na = np.diff(self.rec_ref_offsets.astype(np.int64))
            if len(na) and na.max() > 255:
                raise ValueError("a record has more than 255 alignments: rec_na8 cannot be used")
            cached = self._na8 = na.astype(np.uint8)
        return cached

    def pack24(self, alloc=None):
        """(rec_umi24, refs24): little-endian 24-bit packed copies of rec_umi32 / refs (the compact
        wire arrays of afq_batch); `alloc(n_bytes, dtype)` may hand out pinned memory."""
        cached = getattr(self, "_p24", None)
        if cached is None:
            alloc = alloc or (lambda n, dt: np.empty(n, dtype=dt))
            out = []
            for name, src in (("rec_umi32", self.rec_umi32), ("refs", self.refs)):
                if len(src) and int(src.max()) >= 1 << 24:
                    raise ValueError(f"{name} holds a value >= 2^24: the 24-bit wire array cannot be used")
                dst = alloc(3 * len(src) + 16, np.uint8)
                dst[:3 * len(src)].reshape(-1, 3)[:] = src.view(np.uint8).reshape(-1, 4)[:, :3]
                out.append(dst)
            cached = self._p24 = tuple(out)
        return cached

    def to_c(self, use_na8: bool = False, use_pack24: bool = False) -> AfqBatch:
        b = AfqBatch()
        b.first_cell_index = self.first_cell_index
        b.n_cells, b.n_records, b.n_refs_total = self.n_cells, self.n_records, self.n_refs_total
        b.cell_rec_offsets = _ptr(self.cell_rec_offsets)
        if use_pack24:
            u24, r24 = self.pack24()
            b.rec_umi32, b.refs = None, None
            b.rec_umi24, b.refs24 = _ptr(u24), _ptr(r24)
        else:
            b.rec_umi32, b.refs = _ptr(self.rec_umi32), _ptr(self.refs)
            b.rec_umi24, b.refs24 = None, None
        if use_na8:
            b.rec_ref_offsets = None
            b.rec_na8 = _ptr(self.na8())
        else:
            b.rec_ref_offsets = _ptr(self.rec_ref_offsets)
            b.rec_na8 = None
        return b

    def slice_cells(self, c0: int, c1: int) -> "CellBatch":
        r0, r1 = int(self.cell_rec_offsets[c0]), int(self.cell_rec_offsets[c1])
        f0, f1 = int(self.rec_ref_offsets[r0]), int(self.rec_ref_offsets[r1])
        return CellBatch(self.cell_rec_offsets[c0:c1 + 1] - np.uint64(r0), self.rec_umi32[r0:r1],
                         self.rec_ref_offsets[r0:r1 + 1] - np.uint32(f0), self.refs[f0:f1],
                         self.first_cell_index + c0)

    @staticmethod
    def from_cells(cells, first_cell_index=0) -> "CellBatch":
        """cells: list of cells, each a list of (umi, [refs...]) records."""
        cro, umi, ro, refs = [0], [], [0], []
        for cell in cells:
            for (u, rs) in cell:
                umi.append(u)
                refs.extend(rs)
                ro.append(len(refs))
            cro.append(len(umi))
        return CellBatch(np.array(cro, dtype=np.uint64), np.array(umi, dtype=np.uint32),
                         np.array(ro, dtype=np.uint32), np.array(refs, dtype=np.uint32), first_cell_index)


@dataclass
class EqcDump:
    """Per-cell gene eq-classes (afq_eqc_dump): cell c owns classes [cell_cls_ptr[c], cell_cls_ptr[c+1])."""
    cell_cls_ptr: np.ndarray
    cls_lab_ptr: np.ndarray
    labels: np.ndarray
    counts: np.ndarray

    def cell(self, c):
        """[(label tuple, count)] of cell c, in canonical (lexicographic) order"""
        a, b = int(self.cell_cls_ptr[c]), int(self.cell_cls_ptr[c + 1])
        return [(tuple(int(x) for x in self.labels[int(self.cls_lab_ptr[k]):int(self.cls_lab_ptr[k + 1])]), int(self.counts[k])) for k in range(a, b)]

    @staticmethod
    def from_c(d) -> "EqcDump":
        def arr(p, n, dt):
            if n == 0 or not p:
                return np.zeros(0, dtype=dt)
            buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(p)
            return np.frombuffer(buf, dtype=dt, count=n).copy()
        nc, ncl, nl = int(d.n_cells), int(d.n_classes), int(d.n_labels)
        return EqcDump(arr(d.cell_cls_ptr, nc + 1, np.uint64), arr(d.cls_lab_ptr, ncl + 1, np.uint64), arr(d.labels, nl, np.uint32),
                       arr(d.counts, ncl, np.uint32))


@dataclass
class QuantResult:
    """Per-cell sparse counts in input cell order (CSR, ascending columns) + featureDump stats."""
    row_ptr: np.ndarray
    col: np.ndarray
    val: np.ndarray
    sum_umi: np.ndarray
    max_umi: np.ndarray
    num_expr: np.ndarray
    num_over_mean: np.ndarray
    flags: np.ndarray

    @property
    def n_cells(self): return len(self.row_ptr) - 1
    @property
    def nnz(self): return int(self.row_ptr[-1])

    def row(self, c):
        a, b = int(self.row_ptr[c]), int(self.row_ptr[c + 1])
        return self.col[a:b], self.val[a:b]

    @staticmethod
    def from_c(r: AfqResult) -> "QuantResult":
        nc, nnz = int(r.n_cells), int(r.nnz)
        def arr(p, n, dt):
            if n == 0 or not p:
                return np.zeros(0, dtype=dt)
            buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(p)
            return np.frombuffer(buf, dtype=dt, count=n).copy()
        return QuantResult(arr(r.row_ptr, nc + 1, np.uint64), arr(r.col, nnz, np.uint32), arr(r.val, nnz, np.float32),
                           arr(r.sum_umi, nc, np.float32), arr(r.max_umi, nc, np.float32),
                           arr(r.num_expr, nc, np.uint32), arr(r.num_over_mean, nc, np.uint32),
                           arr(r.flags, nc, np.uint8))


class Quantifier:
    """One afq_ctx (one GPU). `quantify_batch` is the host-buffer (e2e) call; `submit`/`wait`
    expose the asynchronous pipeline; `quant_device` runs on device-resident torch tensors."""

    def __init__(self, opts: QuantOpts, tid_to_gid: np.ndarray):
        self._lib = _abi.lib()
        self.opts = opts
        self._t2g = np.ascontiguousarray(tid_to_gid, dtype=np.uint32)
        self._ctx = C.c_void_p()
        cfg = opts.to_c()
        rc = self._lib.afq_create(C.byref(cfg), _ptr(self._t2g), len(self._t2g), C.byref(self._ctx))
        if rc != 0:
            msg = self._lib.afq_last_error(None)
            raise AfqError(rc, msg.decode() if msg else "afq_create failed")
        self._keep = {}

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._lib.afq_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self): return self
    def __exit__(self, *a): self.close()

    def _check(self, rc):
        if rc != 0:
            msg = self._lib.afq_last_error(self._ctx)
            raise AfqError(rc, msg.decode() if msg else "")

    # ---- host API ---------------------------------------------------------------
    def submit(self, batch: CellBatch, use_na8: bool = False, use_pack24: bool = False) -> int:
        cb = batch.to_c(use_na8, use_pack24)
        t = C.c_uint64()
        self._check(self._lib.afq_submit(self._ctx, C.byref(cb), C.byref(t)))
        self._keep[t.value] = batch  # keep host arrays alive until the H2D copies are done
        return t.value

    def wait(self, ticket: int, copy: bool = True):
        r = AfqResult()
        self._check(self._lib.afq_wait(self._ctx, ticket, C.byref(r)))
        self._keep.pop(ticket, None)
        out = QuantResult.from_c(r) if copy else (int(r.n_cells), int(r.nnz))
        self._lib.afq_result_release(self._ctx, C.byref(r))
        return out

    def quantify_batch(self, batch: CellBatch, use_na8: bool = False, use_pack24: bool = False) -> QuantResult:
        return self.wait(self.submit(batch, use_na8, use_pack24))

    def quantify_batch_with_classes(self, batch: CellBatch, use_na8: bool = False, use_pack24: bool = False):
        """(QuantResult, EqcDump) — needs QuantOpts.dump_eq (afq_result_eqclasses)."""
        t = self.submit(batch, use_na8, use_pack24)
        r = AfqResult()
        self._check(self._lib.afq_wait(self._ctx, t, C.byref(r)))
        self._keep.pop(t, None)
        try:
            d = _abi.AfqEqcDump()
            self._check(self._lib.afq_result_eqclasses(self._ctx, C.byref(r), C.byref(d)))
            return QuantResult.from_c(r), EqcDump.from_c(d)
        finally:
            self._lib.afq_result_release(self._ctx, C.byref(r))

    # ---- device API (torch tensors on this ctx's GPU) ----------------------------
    def quant_device(self, dev_batch: dict, dev_out: dict, stream_ptr: int = 0):
        """dev_batch: dict of torch cuda tensors cell_rec_offsets(int64), rec_umi32(int32),
        rec_ref_offsets(int32), refs(int32); dev_out: row_ptr(int64), col(int32), val(float32),
        sum_umi, max_umi (float32), num_expr, num_over_mean (int32), flags (uint8)."""
        b = AfqBatch()
        b.first_cell_index = 0
        b.n_cells = dev_batch["cell_rec_offsets"].numel() - 1
        b.n_records = dev_batch["rec_umi32"].numel()
        b.n_refs_total = dev_batch["refs"].numel()
        b.cell_rec_offsets = dev_batch["cell_rec_offsets"].data_ptr()
        b.rec_umi32 = dev_batch["rec_umi32"].data_ptr()
        b.rec_ref_offsets = dev_batch["rec_ref_offsets"].data_ptr()
        b.refs = dev_batch["refs"].data_ptr()
        o = AfqDeviceOut()
        o.row_ptr = dev_out["row_ptr"].data_ptr(); o.cap_cells = dev_out["row_ptr"].numel()
        o.col = dev_out["col"].data_ptr(); o.val = dev_out["val"].data_ptr()
        o.cap_nnz = min(dev_out["col"].numel(), dev_out["val"].numel())
        o.sum_umi = dev_out["sum_umi"].data_ptr(); o.max_umi = dev_out["max_umi"].data_ptr()
        o.num_expr = dev_out["num_expr"].data_ptr(); o.num_over_mean = dev_out["num_over_mean"].data_ptr()
        o.flags = dev_out["flags"].data_ptr()
        self._check(self._lib.afq_quant_device(self._ctx, C.byref(b), C.byref(o), C.c_void_p(stream_ptr)))

    def device_finish(self, stream_ptr: int = 0, dev_row_ptr=None) -> int:
        nnz = C.c_uint64(0)
        if dev_row_ptr is not None:
            self._check(self._lib.afq_device_finish(self._ctx, C.c_void_p(stream_ptr), C.byref(nnz),
                                                    C.c_void_p(dev_row_ptr.data_ptr()), dev_row_ptr.numel() - 1))
        else:
            self._check(self._lib.afq_device_finish(self._ctx, C.c_void_p(stream_ptr), None, None, 0))
        return nnz.value

    # ---- infer: EM over a global gene-eq-class table (src/infer.rs) ------------------
    def infer(self, label_offsets, labels, cell_offsets, cell_eq, cell_cnt) -> QuantResult:
        """afq_infer: per-cell EM (em_optimize_subset, src/em.rs:251-456) from the (class id, count) rows of a
        geqc_counts matrix and the global class table of gene_eqclass.txt.gz. Host arrays in, QuantResult out."""
        lo = np.ascontiguousarray(label_offsets, dtype=np.uint32); lb = np.ascontiguousarray(labels, dtype=np.uint32)
        co = np.ascontiguousarray(cell_offsets, dtype=np.uint64); ce = np.ascontiguousarray(cell_eq, dtype=np.uint32)
        cc = np.ascontiguousarray(cell_cnt, dtype=np.uint32)
        t = _abi.AfqEqcTable(len(lo) - 1, _ptr(lo), _ptr(lb))
        r = AfqResult()
        self._check(self._lib.afq_infer(self._ctx, C.byref(t), len(co) - 1, _ptr(co), _ptr(ce), _ptr(cc), C.byref(r)))
        return QuantResult.from_c(r)

    # ---- introspection ----------------------------------------------------------
    @property
    def launch_count(self) -> int:
        return int(self._lib.afq_launch_count(self._ctx))

    @property
    def rerun_count(self) -> int:
        """batches that afq_wait ran again after growing a device arena"""
        return int(self._lib.afq_rerun_count(self._ctx))

    def set_profiling(self, on: bool):
        self._check(self._lib.afq_set_profiling(self._ctx, int(on)))

    def profile_reset(self):
        self._check(self._lib.afq_profile_reset(self._ctx))

    def profile(self) -> dict:
        """{kernel name: (ms, launches)} accumulated since the last reset (CUDA events)."""
        self._check(self._lib.afq_profile_collect(self._ctx))
        out = {}
        i = 0
        while True:
            name = C.c_char_p(); ms = C.c_double(); n = C.c_uint64()
            if self._lib.afq_profile_get(self._ctx, i, C.byref(name), C.byref(ms), C.byref(n)) != 0:
                break
            if n.value:
                out[name.value.decode()] = (ms.value, n.value)
            i += 1
        return out
