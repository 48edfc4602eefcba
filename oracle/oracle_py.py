"""oracle/oracle_py.py — ctypes access to the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this. The product path (zig-bpe_b200/) never does.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError(f"{LIB} missing: run `make -C oracle` (or __graft_entry__.build())")
        L = ctypes.CDLL(LIB)
        c = ctypes
        L.oracle_wyhash.restype = c.c_uint64
        L.oracle_wyhash.argtypes = [c.c_uint64, c.c_void_p, c.c_size_t]
        L.oracle_hash_pair.restype = c.c_uint64
        L.oracle_hash_pair.argtypes = [c.c_uint16, c.c_uint16]
        L.oracle_train.argtypes = [c.c_void_p, c.c_size_t, c.c_uint, c.c_long, c.c_int, c.c_void_p, c.c_void_p,
                                   c.POINTER(c.c_size_t), c.c_void_p]
        L.oracle_encode.argtypes = [c.c_void_p, c.c_size_t, c.c_void_p, c.c_size_t, c.c_int, c.c_void_p, c.POINTER(c.c_size_t)]
        L.oracle_decode.argtypes = [c.c_void_p, c.c_size_t, c.c_void_p, c.c_size_t, c.c_void_p, c.c_size_t, c.POINTER(c.c_size_t)]
        L.oracle_serialize.argtypes = [c.c_void_p, c.c_size_t, c.c_char_p]
        L.oracle_deserialize.restype = c.c_long
        L.oracle_deserialize.argtypes = [c.c_char_p, c.c_void_p, c.c_size_t]
        L.oracle_zigmap_slots.restype = c.c_uint32
        L.oracle_zigmap_slots.argtypes = [c.c_void_p, c.c_size_t, c.c_void_p]
        _lib = L
    return _lib


def _u8(data):
    if isinstance(data, (bytes, bytearray)):
        return np.frombuffer(bytes(data), dtype=np.uint8)
    return np.ascontiguousarray(data, dtype=np.uint8)


def _m(merges):
    return np.ascontiguousarray(np.array([tuple(int(x) for x in t) for t in merges], dtype=np.uint16).reshape(-1, 3))


def wyhash(seed, data: bytes) -> int:
    a = _u8(data)
    return lib().oracle_wyhash(seed, a.ctypes.data, a.size)


def hash_pair(first, second) -> int:
    return lib().oracle_hash_pair(first, second)


def train(data, vocab_size, max_steps=-1, fast=True, stats=False):
    """-> (merges (k,3) uint16, counts (k,) uint64[, stats8])"""
    a = _u8(data)
    cap = max(vocab_size - 256, 1)
    m = np.zeros((cap, 3), dtype=np.uint16)
    c = np.zeros(cap, dtype=np.uint64)
    n = ctypes.c_size_t(0)
    st = np.zeros(8, dtype=np.float64)
    rc = lib().oracle_train(a.ctypes.data, a.size, vocab_size, max_steps, 1 if fast else 0, m.ctypes.data, c.ctypes.data,
                            ctypes.byref(n), st.ctypes.data)
    if rc:
        raise ValueError(f"oracle_train rc={rc}")
    if stats:
        return m[: n.value].copy(), c[: n.value].copy(), st
    return m[: n.value].copy(), c[: n.value].copy()


def encode(data, merges, linear=True):
    a = _u8(data)
    m = _m(merges)
    out = np.zeros(max(a.size, 1), dtype=np.uint16)
    n = ctypes.c_size_t(0)
    rc = lib().oracle_encode(a.ctypes.data, a.size, m.ctypes.data, len(m), 1 if linear else 0, out.ctypes.data, ctypes.byref(n))
    if rc:
        raise ValueError(f"oracle_encode rc={rc}")
    return out[: n.value].copy()


def decode(tokens, merges, cap=None):
    """-> (rc, bytes)"""
    t = np.ascontiguousarray(tokens, dtype=np.uint16)
    m = _m(merges)
    cap = cap or max(64, t.size * 64)
    out = np.zeros(cap, dtype=np.uint8)
    n = ctypes.c_size_t(0)
    rc = lib().oracle_decode(t.ctypes.data, t.size, m.ctypes.data, len(m), out.ctypes.data, cap, ctypes.byref(n))
    return rc, out[: min(n.value, cap)].tobytes()


def serialize(merges, path):
    m = _m(merges)
    return lib().oracle_serialize(m.ctypes.data, len(m), str(path).encode())


def deserialize(path, cap=70000):
    """-> (rc_or_count, merges list)"""
    m = np.zeros((cap, 3), dtype=np.uint16)
    r = lib().oracle_deserialize(str(path).encode(), m.ctypes.data, cap)
    k = r if r >= 0 else None
    if k is None:
        # count parsed rows: rows are filled in order; trailing zeros are ambiguous, so re-read leniently
        k = 0
        for row in m:
            if not row.any():
                break
            k += 1
    return r, [tuple(int(x) for x in row) for row in m[:k]]


def zigmap_slots(pairs):
    p = np.ascontiguousarray(np.array(pairs, dtype=np.uint16).reshape(-1, 2))
    s = np.zeros(len(p), dtype=np.uint32)
    cap = lib().oracle_zigmap_slots(p.ctypes.data, len(p), s.ctypes.data)
    return cap, s
