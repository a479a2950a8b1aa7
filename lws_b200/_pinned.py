"""Result arrays in page-locked host memory.

Every reference call returns a NEW complex128 array.  Filling fresh pageable memory costs a page fault per 4 KB and a
staging pass (measured on BASELINE configs[1]: 330 MB of results, ~70 ms); the DMA engine writes a pinned buffer
directly at PCIe speed (~7 ms).  The arrays handed out here are ordinary ``numpy.ndarray`` objects whose memory comes
from ``lwsb_host_alloc``; when the last reference to one dies its block goes back to a small pool, so a loop that keeps
calling ``batch_lws`` on same-sized batches reuses the same few blocks.  If pinned memory cannot be had (no device,
allocation failure, pool budget exhausted) the caller falls back to ``numpy.empty``.
"""
from __future__ import annotations

import ctypes
import os
import threading

import numpy as np

from . import _native

_BUDGET = int(float(os.environ.get("LWSB_PINNED_POOL_MB", "4096")) * (1 << 20))  # pinned bytes this process may hold
_MIN = 1 << 20                                                                   # smaller results are not worth pinning
_lock = threading.Lock()
_free = {}      # nbytes -> [ptr, ...]
_held = 0       # bytes allocated (handed out + pooled)


class _Block(object):
    """owner of one pinned allocation; returns it to the pool when the array that wraps it is collected"""

    def __init__(self, ptr, nbytes):
        self.ptr, self.nbytes = ptr, nbytes

    def __del__(self):
        try:
            with _lock:
                _free.setdefault(self.nbytes, []).append(self.ptr)
        except Exception:  # interpreter shutdown
            pass


def _take(nbytes):
    global _held
    with _lock:
        lst = _free.get(nbytes)
        if lst:
            return lst.pop()
        if _held + nbytes > _BUDGET:
            # make room: release pooled blocks of other sizes
            for n in list(_free):
                while _free[n] and _held + nbytes > _BUDGET:
                    _native.lib().lwsb_host_free(ctypes.c_void_p(_free[n].pop()))
                    _held -= n
            if _held + nbytes > _BUDGET:
                return None
        _held += nbytes
    p = ctypes.c_void_p()
    if _native.lib().lwsb_host_alloc(nbytes, ctypes.byref(p)) != 0 or not p.value:
        with _lock:
            _held -= nbytes
        return None
    return p.value


def empty(shape, dtype=np.complex128):
    """np.empty(shape, dtype) in pinned memory when possible"""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    if n < _MIN or os.environ.get("LWSB_PINNED_RESULTS", "1") == "0":
        return np.empty(shape, dtype=dtype)
    try:
        ptr = _take(n)
    except Exception:
        ptr = None
    if ptr is None:
        return np.empty(shape, dtype=dtype)
    buf = (ctypes.c_char * n).from_address(ptr)
    buf._lwsb_owner = _Block(ptr, n)  # lives as long as any array (or view) built on this buffer
    return np.frombuffer(buf, dtype=dtype).reshape(shape)
