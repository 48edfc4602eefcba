"""tools/synthcorpus.py — ctypes access to the deterministic corpus generator (tools/synthcorpus.c)."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "libsynthcorpus.so")
ASCII, UTF8, BYTE = 0, 1, 2
# seeds of SURVEY.md §8(d)
SEED_C2, SEED_C3, SEED_C4, SEED_C5 = 0x5EED0002, 0x5EED0003, 0x5EED0004, 0x5EED0005
_lib = None


def _l():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError(f"{LIB} missing: run __graft_entry__.build()")
        _lib = ctypes.CDLL(LIB)
        _lib.synth_corpus_range.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_int]
    return _lib


def generate(n, seed, variant=BYTE, offset=0, nthreads=None, out=None) -> np.ndarray:
    """bytes [offset, offset+n) of corpus (seed, variant) as a uint8 array"""
    if out is None:
        out = np.empty(n, dtype=np.uint8)
    assert out.dtype == np.uint8 and out.size >= n
    nthreads = nthreads or max(1, min(32, (os.cpu_count() or 1)))
    rc = _l().synth_corpus_range(out.ctypes.data, offset, n, seed, variant, nthreads)
    if rc:
        raise MemoryError("synth_corpus_range failed")
    return out[:n]
