"""ctypes access to the TEST-ONLY CPU emulation of the CUDA kernels (tests/emu). It compiles
the product's device code with -DAFQ_EMU for the host; nothing under alevin_fry_b200/ uses it."""
import ctypes as C
import os
import subprocess

import numpy as np

from alevin_fry_b200._abi import AfqBatch, AfqConfig, AfqResult, REPO_ROOT
from alevin_fry_b200.quant import CellBatch, QuantOpts, QuantResult

EMU_DIR = os.path.join(REPO_ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "libafq_emu.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(EMU_DIR, "emu_pipeline.cpp"), os.path.join(EMU_DIR, "cuda_emu.h")]
    csrc = os.path.join(REPO_ROOT, "alevin_fry_b200", "csrc")
    srcs += [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not force and os.path.exists(EMU_LIB) and all(os.path.getmtime(s) <= os.path.getmtime(EMU_LIB) for s in srcs):
        return
    cmd = ["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-ffp-contract=off", "-DAFQ_EMU", "-I" + EMU_DIR, "-shared",
           "-o", EMU_LIB, os.path.join(EMU_DIR, "emu_pipeline.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emu build failed:\n" + r.stderr)


def lib():
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(EMU_LIB)
        l.afq_emu_quant.restype = C.c_int
        l.afq_emu_quant.argtypes = [C.POINTER(AfqConfig), C.c_void_p, C.c_uint64, C.POINTER(AfqBatch), C.POINTER(AfqResult),
                                    C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32)]
        l.afq_emu_quant_dump.restype = C.c_int
        l.afq_emu_quant_dump.argtypes = [C.POINTER(AfqConfig), C.c_void_p, C.c_uint64, C.POINTER(AfqBatch), C.POINTER(AfqResult),
                                         C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32), C.c_void_p]
        l.afq_emu_last_counts.restype = None
        l.afq_emu_last_counts.argtypes = [C.POINTER(C.c_uint32), C.c_int]
        l.afq_emu_infer.restype = C.c_int
        l.afq_emu_infer.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.POINTER(AfqResult), C.POINTER(C.c_void_p), C.c_uint32]
        l.afq_emu_release.restype = None
        l.afq_emu_release.argtypes = [C.c_void_p]
        _lib = l
    return _lib


def emu_quant(opts: QuantOpts, tid_to_gid, batch: CellBatch, use_na8: bool = False, use_pack24: bool = False) -> QuantResult:
    t2g = np.ascontiguousarray(tid_to_gid, dtype=np.uint32)
    cfg = opts.to_c()
    cb = batch.to_c(use_na8, use_pack24)
    r = AfqResult()
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    dev = C.c_uint32(0)
    rc = lib().afq_emu_quant(C.byref(cfg), t2g.ctypes.data_as(C.c_void_p), len(t2g), C.byref(cb), C.byref(r), C.byref(h),
                             err, 512, C.byref(dev))
    if rc != 0:
        raise RuntimeError(f"emu pipeline failed rc={rc} dev_error={dev.value}: {err.value.decode()}")
    out = QuantResult.from_c(r)
    lib().afq_emu_release(h)
    return out


# bin_list rows (alevin_fry_b200/csrc: NUM_BINS = 7)
LIST_GE_BIG, LIST_GE_NORMAL, LIST_OVF, LIST_PS0 = 7, 8, 9, 10


def last_counts():
    """cells per work list in the last emu_quant call (which kernel took how many cells)"""
    out = (C.c_uint32 * 14)()
    lib().afq_emu_last_counts(out, 14)
    return list(out)


def emu_quant_with_classes(opts: QuantOpts, tid_to_gid, batch: CellBatch):
    """(QuantResult, EqcDump) from the emulated pipeline with --dump-eqclasses."""
    from alevin_fry_b200._abi import AfqEqcDump
    from alevin_fry_b200.quant import EqcDump
    t2g = np.ascontiguousarray(tid_to_gid, dtype=np.uint32)
    cfg = opts.to_c()
    cb = batch.to_c()
    r, d, h = AfqResult(), AfqEqcDump(), C.c_void_p()
    err = C.create_string_buffer(512)
    dev = C.c_uint32(0)
    rc = lib().afq_emu_quant_dump(C.byref(cfg), t2g.ctypes.data_as(C.c_void_p), len(t2g), C.byref(cb), C.byref(r), C.byref(h),
                                  err, 512, C.byref(dev), C.byref(d))
    if rc != 0:
        raise RuntimeError(f"emu pipeline failed rc={rc} dev_error={dev.value}: {err.value.decode()}")
    out = QuantResult.from_c(r), EqcDump.from_c(d)
    lib().afq_emu_release(h)
    return out


def emu_infer(num_alphas, usa, init_uniform, lab_off, labels, cell_off, cell_eq, cell_cnt, force_global=False) -> QuantResult:
    """k_em_subset (the kernel behind afq_infer) under emulation."""
    lab_off = np.ascontiguousarray(lab_off, dtype=np.uint32); labels = np.ascontiguousarray(labels, dtype=np.uint32)
    cell_off = np.ascontiguousarray(cell_off, dtype=np.uint64); cell_eq = np.ascontiguousarray(cell_eq, dtype=np.uint32)
    cell_cnt = np.ascontiguousarray(cell_cnt, dtype=np.uint32)
    r = AfqResult()
    h = C.c_void_p()
    rc = lib().afq_emu_infer(num_alphas, int(usa), int(init_uniform), len(lab_off) - 1, lab_off.ctypes.data, labels.ctypes.data,
                             len(cell_off) - 1, cell_off.ctypes.data, cell_eq.ctypes.data, cell_cnt.ctypes.data, C.byref(r), C.byref(h),
                             int(force_global))
    assert rc == 0, rc
    out = QuantResult.from_c(r)
    lib().afq_emu_release(h)
    return out
