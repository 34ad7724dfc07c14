"""ctypes mirror of include/afq.h (the C-ABI boundary) and library loading.

The product library (libafq.so, CUDA) is loaded from this package directory. There is
no CPU fallback: if the library is missing or no sm_100 GPU is usable, calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)

AFQ_OK, AFQ_ERR_INVALID, AFQ_ERR_UNSUPPORTED, AFQ_ERR_CUDA, AFQ_ERR_NO_DEVICE, AFQ_ERR_INTERNAL = range(6)

RES_TRIVIAL, RES_CR_LIKE, RES_CR_LIKE_EM, RES_PARSIMONY_EM, RES_PARSIMONY, RES_PARSIMONY_GENE_EM, RES_PARSIMONY_GENE = range(7)
# `-r/--resolution` strings accepted by the reference CLI (src/quant.rs:100-113)
RESOLUTIONS = {
    "trivial": RES_TRIVIAL,
    "cr-like": RES_CR_LIKE,
    "cr-like-em": RES_CR_LIKE_EM,
    "parsimony-em": RES_PARSIMONY_EM,
    "parsimony": RES_PARSIMONY,
    "parsimony-gene-em": RES_PARSIMONY_GENE_EM,
    "parsimony-gene": RES_PARSIMONY_GENE,
}
SA_WINNER_TAKE_ALL, SA_PREFER_AMBIG = 0, 1
FLAG_TINY, FLAG_ALT, FLAG_EMPTY = 1, 2, 4


class AfqConfig(C.Structure):
    _fields_ = [
        ("resolution", C.c_int32), ("usa_mode", C.c_int32), ("em_init_uniform", C.c_int32),
        ("pug_exact_umi", C.c_int32), ("sa_model", C.c_int32), ("dump_eq", C.c_int32),
        ("num_gene_ids", C.c_uint32), ("num_rows", C.c_uint32),
        ("small_thresh", C.c_uint64), ("large_graph_thresh", C.c_uint64),
        ("barcode_len", C.c_uint16), ("umi_len", C.c_uint16), ("device", C.c_int32),
    ]


class AfqBatch(C.Structure):
    _fields_ = [
        ("first_cell_index", C.c_uint64), ("n_cells", C.c_uint64), ("n_records", C.c_uint64),
        ("n_refs_total", C.c_uint64),
        ("cell_rec_offsets", C.c_void_p), ("rec_umi32", C.c_void_p),
        ("rec_ref_offsets", C.c_void_p), ("refs", C.c_void_p), ("rec_na8", C.c_void_p),
        ("rec_umi24", C.c_void_p), ("refs24", C.c_void_p),
    ]


class AfqResult(C.Structure):
    _fields_ = [
        ("n_cells", C.c_uint64), ("nnz", C.c_uint64),
        ("row_ptr", C.c_void_p), ("col", C.c_void_p), ("val", C.c_void_p),
        ("sum_umi", C.c_void_p), ("max_umi", C.c_void_p),
        ("num_expr", C.c_void_p), ("num_over_mean", C.c_void_p), ("flags", C.c_void_p),
    ]


class AfqDeviceOut(C.Structure):
    _fields_ = [
        ("row_ptr", C.c_void_p), ("cap_cells", C.c_uint64),
        ("col", C.c_void_p), ("val", C.c_void_p), ("cap_nnz", C.c_uint64),
        ("sum_umi", C.c_void_p), ("max_umi", C.c_void_p),
        ("num_expr", C.c_void_p), ("num_over_mean", C.c_void_p), ("flags", C.c_void_p),
    ]


class AfqEqcDump(C.Structure):
    _fields_ = [("n_cells", C.c_uint64), ("n_classes", C.c_uint64), ("n_labels", C.c_uint64), ("cell_cls_ptr", C.c_void_p),
                ("cls_lab_ptr", C.c_void_p), ("labels", C.c_void_p), ("counts", C.c_void_p)]


class AfqEqcTable(C.Structure):
    _fields_ = [("n_classes", C.c_uint64), ("label_offsets", C.c_void_p), ("labels", C.c_void_p)]


# every symbol include/afq.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "afq_create": (C.c_int, [C.POINTER(AfqConfig), C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "afq_destroy": (None, [C.c_void_p]),
    "afq_last_error": (C.c_char_p, [C.c_void_p]),
    "afq_submit": (C.c_int, [C.c_void_p, C.POINTER(AfqBatch), C.POINTER(C.c_uint64)]),
    "afq_wait": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(AfqResult)]),
    "afq_result_release": (None, [C.c_void_p, C.POINTER(AfqResult)]),
    "afq_quant_device": (C.c_int, [C.c_void_p, C.POINTER(AfqBatch), C.POINTER(AfqDeviceOut), C.c_void_p]),
    "afq_device_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p, C.c_uint64]),
    "afq_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "afq_host_free": (None, [C.c_void_p]),
    "afq_device_count": (C.c_int, []),
    "afq_result_eqclasses": (C.c_int, [C.c_void_p, C.POINTER(AfqResult), C.POINTER(AfqEqcDump)]),
    "afq_infer": (C.c_int, [C.c_void_p, C.POINTER(AfqEqcTable), C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(AfqResult)]),
    "afq_abi_version": (C.c_int, []),
    "afq_launch_count": (C.c_uint64, [C.c_void_p]),
    "afq_rerun_count": (C.c_uint64, [C.c_void_p]),
    "afq_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "afq_profile_collect": (C.c_int, [C.c_void_p]),
    "afq_profile_reset": (C.c_int, [C.c_void_p]),
    "afq_profile_get": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
}

LIB_PATH = os.path.join(_HERE, "libafq.so")
_lib = None


class AfqError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"afq error {code}: {msg}")
        self.code = code


def lib():
    """Load libafq.so (the CUDA product). Raises if it has not been built — loudly, never
    substituting another implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()' "
                "or `make`). alevin_fry_b200 has no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib
