"""Pool of page-locked result buffers.

`AssemblyPlan.assemble` returns fresh numpy arrays on every call (value semantics, as the reference's assembly does).  A
fresh pageable array costs more than the assembly: 18 ms for the 75 MB of config-2 CSR values (page faults + the driver's
staged copy) against 2 ms into page-locked memory (tools/time_plugin.py).  The pool hands out numpy arrays that live in
page-locked memory obtained from the library (femb_host_alloc); when the last reference to such an array -- or to any view
of it, e.g. the `data` of a scipy matrix -- is dropped, its buffer goes back to the pool instead of to the OS.  A caller
that keeps results alive simply makes the pool allocate new buffers; beyond `maxBytes` of pinned memory (16 GB unless
FEMB200_PINNED_POOL_BYTES says otherwise) it returns ordinary pageable arrays.
"""
import ctypes as C
import os
import weakref

import numpy as np


class PinnedPool:
    def __init__(self, lib, maxBytes=None, maxFreePerSize=3):
        if maxBytes is None:  # default 16 GB of page-locked memory at most (FEMB200_PINNED_POOL_BYTES overrides; 0 switches the pool off)
            maxBytes = int(os.environ.get("FEMB200_PINNED_POOL_BYTES", 16 << 30))
        self._lib = lib
        self.maxBytes, self.maxFreePerSize = maxBytes, maxFreePerSize
        self._free = {}   # nbytes -> [pointers]
        self.bytesLive = 0  # handed out or parked

    def empty(self, n, dtype=np.float64):
        """uninitialised array of n items in page-locked memory (pageable when the pool is exhausted)"""
        nbytes = int(n) * np.dtype(dtype).itemsize
        if nbytes == 0:
            return np.empty(0, dtype=dtype)
        ptr = None
        if self._free.get(nbytes):
            ptr = self._free[nbytes].pop()
        elif self.bytesLive + nbytes <= self.maxBytes:
            out = C.c_void_p()
            if self._lib.femb_host_alloc(nbytes, C.byref(out)) == 0 and out.value:
                ptr = out.value
                self.bytesLive += nbytes
        if ptr is None:
            return np.empty(int(n), dtype=dtype)
        buf = (C.c_char * nbytes).from_address(ptr)
        weakref.finalize(buf, self._release, ptr, nbytes)
        return np.frombuffer(buf, dtype=dtype)

    def _release(self, ptr, nbytes):
        try:
            parked = self._free.setdefault(nbytes, [])
            if len(parked) < self.maxFreePerSize:
                parked.append(ptr)
            else:
                self._lib.femb_host_free(ptr)
                self.bytesLive -= nbytes
        except Exception:  # interpreter shutdown: the library may already be gone; the OS reclaims the memory
            pass

    def trim(self):
        """give every parked buffer back to the OS"""
        for nbytes, parked in self._free.items():
            while parked:
                self._lib.femb_host_free(parked.pop())
                self.bytesLive -= nbytes


_pool = None


def pool(lib):
    global _pool
    if _pool is None:
        _pool = PinnedPool(lib)
    return _pool
