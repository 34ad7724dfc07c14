"""alevin_fry_b200 — B200-native implementation of alevin-fry's `quant` hot path.

Host-side mirror of the reference's quant interface (QuantOpts / quantify, reference
src/prog_opts.rs:24-43, src/quant.rs:359) over the C-ABI in include/afq.h. The compute
runs in hand-written sm_100a CUDA kernels (alevin_fry_b200/csrc); PyTorch is used only
for device memory, streams and torch.distributed plumbing.
"""
import os as _os

# libafq's pipeline keeps ~10 CUDA streams busy: more hardware work queues than the default 8 (read when the CUDA context is
# created; afq_cuda.cu sets the same default when the library is loaded — this covers a torch that initialises CUDA first)
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from ._abi import (AfqError, RESOLUTIONS, FLAG_TINY, FLAG_ALT, FLAG_EMPTY, lib, LIB_PATH)
from .quant import QuantOpts, CellBatch, QuantResult, Quantifier

__all__ = ["AfqError", "RESOLUTIONS", "FLAG_TINY", "FLAG_ALT", "FLAG_EMPTY", "lib", "LIB_PATH",
           "QuantOpts", "CellBatch", "QuantResult", "Quantifier"]
