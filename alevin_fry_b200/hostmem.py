"""Pinned host arrays (cudaHostAlloc through the C-ABI) as numpy views."""
import ctypes as C

import numpy as np

from . import _abi


class PinnedPool:
    """Owns pinned allocations; arrays are views and die with the pool."""

    def __init__(self):
        self._ptrs = []

    def empty(self, n, dtype):
        dt = np.dtype(dtype)
        nbytes = max(int(n) * dt.itemsize, 1)
        p = C.c_void_p()
        rc = _abi.lib().afq_host_alloc(C.byref(p), nbytes)
        if rc != 0:
            raise MemoryError(f"afq_host_alloc({nbytes}) failed with {rc}")
        self._ptrs.append(p)
        buf = (C.c_char * nbytes).from_address(p.value)
        return np.frombuffer(buf, dtype=dt, count=int(n))

    def close(self):
        for p in self._ptrs:
            _abi.lib().afq_host_free(p)
        self._ptrs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
