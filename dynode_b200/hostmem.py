"""Page-locked host buffers for the ensemble's host<->device stream (include/dynode_b200_host.h).

`Solution.ys` of an ensemble whose inputs live on the host is written by the GPU's copy engines straight into the
buffer returned here: page-locked, 2 MiB-aligned, backed by transparent huge pages where the kernel grants them
(512x fewer IOMMU / DMA page-table entries than cudaHostAlloc's 4 KiB pages -- what matters when several GPUs of
one host stream out at once), and faulted in by several host threads.
"""

from __future__ import annotations

import ctypes
import os
import weakref

import numpy as np

from . import _lib

HUGEPAGES, NO_PIN = 1, 2


class HostBuffer:
    """Owns one mapping; `.tensor` / `.array` are views that keep it alive."""

    def __init__(self, nbytes: int, hugepages: bool = True, pin: bool = True, threads: int = 0):
        L = _lib.load()
        self.nbytes = int(nbytes)
        self.flags = (HUGEPAGES if hugepages else 0) | (0 if pin else NO_PIN)
        if threads <= 0:
            try:
                threads = min(16, len(os.sched_getaffinity(0)))
            except AttributeError:
                threads = 4
        p = ctypes.c_void_p()
        _lib.check(L.dynode_host_alloc(self.nbytes, self.flags, threads, ctypes.byref(p)))
        self.ptr = p.value
        self._fin = weakref.finalize(self, L.dynode_host_free, ctypes.c_void_p(self.ptr), self.nbytes, self.flags)

    def huge_bytes(self) -> int:
        """Bytes of the mapping backed by huge pages right now (-1: cannot tell)."""
        return int(_lib.load().dynode_host_info(ctypes.c_void_p(self.ptr), self.nbytes))

    def array(self, shape, dtype=np.float64) -> np.ndarray:
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        assert n <= self.nbytes
        raw = (ctypes.c_char * n).from_address(self.ptr)
        a = np.frombuffer(raw, dtype=dtype).reshape(shape)
        _KEEP[id(raw)] = self  # the ctypes view does not own the mapping: tie the buffer's life to the view's
        weakref.finalize(raw, _KEEP.pop, id(raw), None)
        return a

    def tensor(self, shape, dtype=None):
        import torch
        dtype = dtype or torch.float64
        np_dtype = {torch.float64: np.float64, torch.int32: np.int32, torch.uint8: np.uint8,
                    torch.float32: np.float32, torch.int64: np.int64}[dtype]
        t = torch.from_numpy(self.array(shape, np_dtype))
        t._dynode_host_buffer = self
        return t

    def free(self):
        self._fin()


_KEEP: dict = {}


def pinned_empty(shape, dtype=None, hugepages: bool = True):
    """A page-locked (huge-page-backed when possible) host tensor: the `out=` of `simulate_ensemble`."""
    import torch
    dtype = dtype or torch.float64
    n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
    return HostBuffer(max(n, 1), hugepages=hugepages).tensor(tuple(shape), dtype)
