"""zig-bpe_b200 — Python host binding of the B200-native BPE engine (libbpe_b200.so).

Import with ``importlib.import_module("zig-bpe_b200")`` (the directory name is the project
name and is not a Python identifier).

``BasicTokenizer`` mirrors the reference's public struct
(/root/reference/src/basic_tokenizer.zig:52-349): ``train`` / ``encode`` / ``decode`` /
``serializeMerges`` / ``deserializeMerges`` with the same argument meaning, the same errors
(``TrainError.InvalidVocabSize``, ``error.InvalidToken``, ``error.InvalidFormat`` ...) and the
same stderr text (verbose merge line :309, early-stop notice :189, TimeStats block
utils/time_statistics.zig:37-59). The hot loops run in CUDA behind the C ABI of
include/bpe_b200.h; there is no CPU fallback — constructing an ``Engine`` without a usable
GPU raises ``BpeError``.
"""
from __future__ import annotations

import ctypes
import os
import sys
import time
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_long, c_size_t, c_uint16, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libbpe_b200.so")

BPE_OK = 0
BPE_ERR_INVALID_VOCAB = 1
BPE_ERR_OOM = 2
BPE_ERR_INVALID_TOKEN = 3
BPE_ERR_CUDA = 4
BPE_ERR_INVALID_ARG = 5
BPE_ERR_INTERNAL = 7


class BpeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"bpe_b200 error {code}: {msg}")
        self.code = code


class InvalidVocabSize(BpeError):
    """TrainError.InvalidVocabSize (basic_tokenizer.zig:147-149)"""


class InvalidToken(BpeError):
    """error.InvalidToken (basic_tokenizer.zig:101,125,135)"""


class OutOfMemory(BpeError):
    """TrainError.OutOfMemory"""


class MergesFormatError(ValueError):
    """error.InvalidFormat / InvalidCharacter / Overflow / StreamTooLong from deserializeMerges (:332-348)"""

    def __init__(self, zig_error, line_no):
        super().__init__(f"error.{zig_error} at line {line_no}")
        self.zig_error = zig_error
        self.line_no = line_no


class bpe_stats_t(ctypes.Structure):
    _fields_ = [
        ("sort_pairs_ms", c_double),
        ("replace_pair_ms", c_double),
        ("generate_pairs_ms", c_double),
        ("just_count_pairs_ms", c_double),
        ("sort_pairs_calls", c_uint64),
        ("replace_pair_calls", c_uint64),
        ("generate_pairs_calls", c_uint64),
        ("just_count_pairs_calls", c_uint64),
        ("total_ms", c_double),
        ("device_ms", c_double),
        ("scanned_slots", c_uint64),
        ("kernel_launches", c_uint64),
        ("tie_steps", c_uint64),
        ("tie_slow_steps", c_uint64),
        ("compactions", c_uint64),
        ("kernel_ms", c_double * 12),
        ("kernel_calls", c_uint64 * 12),
        ("aeqb_steps", c_uint64),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["kernel_ms"] = list(self.kernel_ms)
        d["kernel_calls"] = list(self.kernel_calls)
        return d


KERNEL_CLASSES = ["init", "argmax", "tie", "replay", "halo", "merge", "apply", "compact", "table", "hostgap", "r10", "r11"]
MERGE_DTYPE = np.dtype([("first", "<u2"), ("second", "<u2"), ("new_token", "<u2")])  # = bpe_merge_t, 6 bytes

_lib = None


def load_library(path: str | None = None):
    """Load libbpe_b200.so (built by __graft_entry__.build()). Fails loudly if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("BPE_B200_LIB") or LIB_PATH
    if not os.path.exists(p):
        raise BpeError(BPE_ERR_CUDA, f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                     "(there is no CPU fallback)")
    lib = ctypes.CDLL(p)
    lib.bpe_version.restype = c_char_p
    lib.bpe_last_error.restype = c_char_p
    lib.bpe_last_error.argtypes = [c_void_p]
    lib.bpe_ctx_create.argtypes = [POINTER(c_void_p), c_int]
    lib.bpe_ctx_create_dist.argtypes = [POINTER(c_void_p), c_int, c_int, c_int, c_void_p]
    lib.bpe_nccl_unique_id.argtypes = [c_void_p]
    lib.bpe_ctx_destroy.argtypes = [c_void_p]
    lib.bpe_ctx_destroy.restype = None
    lib.bpe_ctx_set_option.argtypes = [c_void_p, c_char_p, c_long]
    for name in ("bpe_train", "bpe_train_device"):
        getattr(lib, name).argtypes = [c_void_p, c_void_p, c_size_t, c_uint16, c_void_p, c_void_p, POINTER(c_size_t),
                                       POINTER(bpe_stats_t)]
    for name in ("bpe_encode", "bpe_encode_device"):
        getattr(lib, name).argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, POINTER(c_size_t),
                                       POINTER(bpe_stats_t)]
    lib.bpe_decode_size.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, POINTER(c_size_t)]
    for name in ("bpe_decode", "bpe_decode_device"):
        getattr(lib, name).argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_size_t,
                                       POINTER(c_size_t), POINTER(bpe_stats_t)]
    if path is None:
        _lib = lib
    return lib


def _as_u8(data) -> np.ndarray:
    if isinstance(data, str):
        data = data.encode("utf-8")
    if isinstance(data, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(data), dtype=np.uint8)
    a = np.ascontiguousarray(data)
    if a.dtype != np.uint8:
        raise TypeError("text must be bytes/str/uint8 array")
    return a


def _as_merges(merges) -> np.ndarray:
    if isinstance(merges, np.ndarray) and merges.dtype == MERGE_DTYPE:
        return np.ascontiguousarray(merges)
    out = np.zeros(len(merges), dtype=MERGE_DTYPE)
    for i, m in enumerate(merges):
        out[i] = (m[0], m[1], m[2])
    return out


class Engine:
    """One GPU context (bpe_ctx). ``lib_path`` is only for the test-suite's emulation build."""

    def __init__(self, device: int = 0, rank: int = 0, world: int = 1, nccl_unique_id: bytes | None = None,
                 lib_path: str | None = None):
        self.lib = load_library(lib_path)
        self._ctx = c_void_p()
        if world > 1:
            assert nccl_unique_id is not None and len(nccl_unique_id) == 128
            uid = ctypes.create_string_buffer(nccl_unique_id, 128)
            rc = self.lib.bpe_ctx_create_dist(byref(self._ctx), device, rank, world, uid)
        else:
            rc = self.lib.bpe_ctx_create(byref(self._ctx), device)
        if rc != BPE_OK:
            raise BpeError(rc, self.lib.bpe_last_error(None).decode())
        self.rank, self.world = rank, world
        self.last_stats: dict = {}

    @staticmethod
    def nccl_unique_id(lib_path: str | None = None) -> bytes:
        lib = load_library(lib_path)
        buf = ctypes.create_string_buffer(128)
        rc = lib.bpe_nccl_unique_id(buf)
        if rc != BPE_OK:
            raise BpeError(rc, lib.bpe_last_error(None).decode())
        return buf.raw

    def close(self):
        if self._ctx:
            self.lib.bpe_ctx_destroy(self._ctx)
            self._ctx = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name: str, value: int):
        rc = self.lib.bpe_ctx_set_option(self._ctx, name.encode(), int(value))
        self._check(rc)

    def _check(self, rc):
        if rc == BPE_OK:
            return
        msg = self.lib.bpe_last_error(self._ctx).decode()
        cls = {BPE_ERR_INVALID_VOCAB: InvalidVocabSize, BPE_ERR_INVALID_TOKEN: InvalidToken, BPE_ERR_OOM: OutOfMemory}.get(rc, BpeError)
        raise cls(rc, msg)

    # ---- train ------------------------------------------------------------------------
    def train(self, text, vocab_size: int, device_ptr: int | None = None, n: int | None = None):
        """Returns (merges[MERGE_DTYPE], counts[u64]). ``device_ptr``/``n``: text already in HBM."""
        if not (0 <= vocab_size <= 0xFFFF):
            raise InvalidVocabSize(BPE_ERR_INVALID_VOCAB, "vocabSize is a u16 in the reference")
        cap = max(vocab_size - 256, 0)
        merges = np.zeros(max(cap, 1), dtype=MERGE_DTYPE)
        counts = np.zeros(max(cap, 1), dtype=np.uint64)
        out_n = c_size_t(0)
        st = bpe_stats_t()
        if device_ptr is not None:
            rc = self.lib.bpe_train_device(self._ctx, c_void_p(device_ptr), n, vocab_size, merges.ctypes.data,
                                           counts.ctypes.data, byref(out_n), byref(st))
        else:
            a = _as_u8(text)
            rc = self.lib.bpe_train(self._ctx, a.ctypes.data, a.size, vocab_size, merges.ctypes.data, counts.ctypes.data,
                                    byref(out_n), byref(st))
        self.last_stats = st.as_dict()
        self._check(rc)
        return merges[: out_n.value].copy(), counts[: out_n.value].copy()

    # ---- encode -----------------------------------------------------------------------
    def encode(self, text, merges, out=None) -> np.ndarray:
        """out: optional caller-owned uint16 buffer of at least len(text) ids (e.g. page-locked and reused between calls, as
        a C caller of bpe_encode would do); the result is then a view of it"""
        a = _as_u8(text)
        m = _as_merges(merges)
        own = out is None
        if own:
            out = np.empty(max(a.size, 1), dtype=np.uint16)  # the caller-allocated upper bound of the C ABI (n ids)
        elif out.dtype != np.uint16 or not out.flags["C_CONTIGUOUS"] or out.size < max(a.size, 1):
            raise ValueError("out must be a contiguous uint16 array of at least len(text) elements")
        out_n = c_size_t(0)
        st = bpe_stats_t()
        rc = self.lib.bpe_encode(self._ctx, a.ctypes.data, a.size, m.ctypes.data, len(m), out.ctypes.data, byref(out_n), byref(st))
        self.last_stats = st.as_dict()
        self._check(rc)
        if own and out_n.value * 4 < out.size:  # small result in a big buffer: shrink (the Zig shim does the same with its allocator)
            return out[: out_n.value].copy()
        return out[: out_n.value]

    def encode_device(self, d_text: int, n: int, merges, d_out: int) -> int:
        m = _as_merges(merges)
        out_n = c_size_t(0)
        st = bpe_stats_t()
        rc = self.lib.bpe_encode_device(self._ctx, c_void_p(d_text), n, m.ctypes.data, len(m), c_void_p(d_out), byref(out_n), byref(st))
        self.last_stats = st.as_dict()
        self._check(rc)
        return out_n.value

    # ---- decode -----------------------------------------------------------------------
    def decode(self, tokens, merges) -> bytes:
        t = np.ascontiguousarray(tokens, dtype=np.uint16)
        m = _as_merges(merges)
        need = c_size_t(0)
        rc = self.lib.bpe_decode_size(self._ctx, t.ctypes.data, t.size, m.ctypes.data, len(m), byref(need))
        self._check(rc)
        out = np.empty(max(need.value, 1), dtype=np.uint8)
        out_n = c_size_t(0)
        st = bpe_stats_t()
        rc = self.lib.bpe_decode(self._ctx, t.ctypes.data, t.size, m.ctypes.data, len(m), out.ctypes.data, need.value, byref(out_n), byref(st))
        self.last_stats = st.as_dict()
        self._check(rc)
        return out[: out_n.value].tobytes()

    def decode_device(self, d_toks: int, n: int, merges, d_out: int, cap: int) -> int:
        m = _as_merges(merges)
        out_n = c_size_t(0)
        st = bpe_stats_t()
        rc = self.lib.bpe_decode_device(self._ctx, c_void_p(d_toks), n, m.ctypes.data, len(m), c_void_p(d_out), cap, byref(out_n), byref(st))
        self.last_stats = st.as_dict()
        self._check(rc)
        return out_n.value


# ---------------------------------------------------------------------------------------
# merges.txt (serializeMerges :319-330 / deserializeMerges :332-348) — host side, as in the reference
# ---------------------------------------------------------------------------------------
def _zig_parse_u16(field: bytes) -> int:
    """std.fmt.parseInt(u16, field, 10): optional sign, '_' separators, overflow check."""
    if len(field) == 0:
        raise ValueError("InvalidCharacter")
    neg = False
    if field[:1] in (b"+", b"-"):
        neg = field[:1] == b"-"
        field = field[1:]
    if len(field) == 0 or field[:1] == b"_" or field[-1:] == b"_":
        raise ValueError("InvalidCharacter")
    acc = 0
    for c in field:
        if c == 0x5F:
            continue
        if not (0x30 <= c <= 0x39):
            raise ValueError("InvalidCharacter")
        d = c - 0x30
        if neg:
            if acc != 0 or d != 0:
                raise ValueError("Overflow")
        else:
            acc = acc * 10 + d
            if acc > 0xFFFF:
                raise ValueError("Overflow")
    return acc


def write_merges(path, merges) -> None:
    with open(path, "wb") as f:  # createFile truncates (:320)
        for m in merges:
            f.write(b"%d,%d,%d\n" % (int(m[0]), int(m[1]), int(m[2])))


def read_merges(path, into: list) -> None:
    """Appends to ``into`` (the reference never clears, :346). Raises MergesFormatError with the
    reference's error name; merges parsed before the bad line stay appended, as in the reference."""
    with open(path, "rb") as f:
        data = f.read()
    pos, line_no = 0, 0
    while pos < len(data):
        line_no += 1
        window = data[pos:pos + 100]  # readUntilDelimiterOrEof into a 100-byte buffer (:339-340)
        k = window.find(b"\n")
        if k < 0:
            if len(window) == 100:
                raise MergesFormatError("StreamTooLong", line_no)
            line, pos = window, len(data)
        else:
            line, pos = window[:k], pos + k + 1
        fields = line.split(b",")
        vals = []
        for i in range(3):
            if i >= len(fields):
                raise MergesFormatError("InvalidFormat", line_no)
            try:
                vals.append(_zig_parse_u16(fields[i]))
            except ValueError as e:
                raise MergesFormatError(str(e), line_no) from None
        into.append((vals[0], vals[1], vals[2]))


class TimeStats:
    """utils/time_statistics.zig:4-34 — ms buckets and call counts, filled from CUDA-event timings."""

    def __init__(self):
        self.sort_pairs_time = 0
        self.sort_pairs_calls = 0
        self.replace_pair_time = 0
        self.replace_pair_calls = 0
        self.generate_pairs_time = 0
        self.generate_pairs_calls = 0
        self.just_count_pairs_time = 0
        self.just_count_pairs_calls = 0


def print_time_stats(stats: TimeStats, total_time_ms: int, file=None) -> None:
    """printTimeStats (utils/time_statistics.zig:36-60); 0 calls prints nan like the reference."""

    file = file or sys.stderr

    def line(name, t, calls):
        avg = (t / (calls * 1000.0)) if calls else float("nan")
        print(f"{name}: {t / 1000.0:.3f}s total, {calls} calls, {avg:.3f}s avg", file=file)

    print("\nTime statistics:", file=file)
    line("sortCodePointPairs", stats.sort_pairs_time, stats.sort_pairs_calls)
    line("replaceTopPairWithIndex", stats.replace_pair_time, stats.replace_pair_calls)
    line("generateCodePointPairs", stats.generate_pairs_time, stats.generate_pairs_calls)
    line("countPointPairs", stats.just_count_pairs_time, stats.just_count_pairs_calls)
    other = total_time_ms - stats.sort_pairs_time - stats.replace_pair_time - stats.generate_pairs_time - stats.just_count_pairs_time
    print(f"Other operations: {other / 1000.0:.3f}s", file=file)


class BasicTokenizer:
    """Drop-in mirror of the reference's ``BasicTokenizer`` (basic_tokenizer.zig:52-349)."""

    def __init__(self, engine: Engine | None = None, quiet: bool = False, **engine_kwargs):
        self.engine = engine or Engine(**engine_kwargs)
        self.merges: list[tuple[int, int, int]] = []  # Merges.merges (first, second, new_token)
        self.merge_counts: list[int] = []
        self.timeStats = TimeStats()
        self.quiet = quiet  # suppress the reference's stderr chatter (tests)

    def put(self, pair, new_token):  # Merges.put (:32-37)
        self.merges.append((int(pair[0]), int(pair[1]), int(new_token)))

    def generateInitialTokens(self, text) -> np.ndarray:  # :155-170
        return _as_u8(text).astype(np.uint16)

    def train(self, text, vocabSize: int, verbose: bool = False) -> None:  # :140-153
        start = time.time()
        try:
            if vocabSize < 256:
                raise InvalidVocabSize(BPE_ERR_INVALID_VOCAB, "InvalidVocabSize")
            self.engine.set_option("time_phases", 0 if self.quiet else 1)
            merges, counts = self.engine.train(text, vocabSize)
            st = self.engine.last_stats
            if not self.quiet:  # generateInitialTokens' own line (:156-160); the widening is the load kernel here
                print(f"generateInitialTokens runtime: {st['kernel_ms'][0] / 1000.0:.3f} seconds", file=sys.stderr)
            ts = self.timeStats
            ts.sort_pairs_time += int(st["sort_pairs_ms"]); ts.sort_pairs_calls += st["sort_pairs_calls"]
            ts.replace_pair_time += int(st["replace_pair_ms"]); ts.replace_pair_calls += st["replace_pair_calls"]
            ts.just_count_pairs_time += int(st["just_count_pairs_ms"]); ts.just_count_pairs_calls += st["just_count_pairs_calls"]
            total = vocabSize - 256
            for i, m in enumerate(merges):
                if verbose and not self.quiet:  # printMergeInfo (:308-317)
                    print(f"merge {i + 1}/{total}: ({m['first']},{m['second']}) -> {m['new_token']} had {counts[i]} occurrences",
                          file=sys.stderr)
                self.merges.append((int(m["first"]), int(m["second"]), int(m["new_token"])))
                self.merge_counts.append(int(counts[i]))
            if len(merges) < total and not self.quiet:
                print("No more pairs to merge. Stopping early.", file=sys.stderr)  # :189
        finally:
            if not self.quiet:
                print_time_stats(self.timeStats, int((time.time() - start) * 1000))  # defer (:142-145)

    def encode(self, text) -> np.ndarray:  # :71-88
        ids = self.engine.encode(text, self.merges)
        if not self.quiet:  # encode calls generateInitialTokens too (:72), which prints its runtime (:156-160); fused away here
            print("generateInitialTokens runtime: 0.000 seconds", file=sys.stderr)
        return ids

    def decode(self, tokens) -> bytes:  # :90-138
        return self.engine.decode(tokens, self.merges)

    def serializeMerges(self, file_path) -> None:  # :319-330
        write_merges(file_path, self.merges)

    def deserializeMerges(self, file_path) -> None:  # :332-348
        read_merges(file_path, self.merges)
