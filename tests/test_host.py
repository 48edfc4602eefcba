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


def test_cpp_mirror_deserialize_edge_cases(ora, tmp_path):
    """The C++ host mirror's merges.txt reader (zig-bpe_b200/host/basic_tokenizer.hpp, no GPU involved) against the
    oracle's restatement of deserializeMerges (:332-348) on the malformed files of test_oracle.py"""
    import subprocess
    src = tmp_path / "rd.cpp"
    src.write_text('''#include "basic_tokenizer.hpp"
#include <cstdio>
int main(int argc, char** argv) {
    std::vector<zigbpe::Merge> ms;
    const char* err = "OK";
    std::string what;
    try { zigbpe::BasicTokenizer::readMergesFile(argv[1], ms); } catch (const std::exception& e) { err = "ERR"; what = e.what(); }
    std::printf("%s %s %zu", err, what.c_str(), ms.size());
    for (auto& m : ms) std::printf(" %u,%u,%u", (unsigned)m.pair.first, (unsigned)m.pair.second, (unsigned)m.new_token);
    std::printf("\\n");
    return 0;
}
''')
    exe = tmp_path / "rd"
    subprocess.run(["g++", "-O1", "-std=c++17", f"-I{ROOT}/include", f"-I{ROOT}/zig-bpe_b200/host", str(src), "-o", str(exe)], check=True)
    cases = [(b"1,2,3\n4,5,6", 2), (b"1,2,3,99\n", 1), (b"1,2\n", -106), (b"1,2,\n", -20), (b"\n", -20), (b"1,2,3\r\n", -20),
             (b"1,2,65536\n", -21), (b"1, 2,3\n", -20), (b"+1,2,3\n", 1), (b"1_0,2,3\n", 1), (b"1,2,3\n" + b"9" * 100 + b"\n", -22),
             (b"1,2,3\n" + b"7," * 49 + b"7\n", 2), (b"", 0), (b"1,2,3", 1)]
    name = {-20: "InvalidCharacter", -21: "Overflow", -22: "StreamTooLong", -106: "InvalidFormat"}
    for content, err in cases:
        p = tmp_path / "m.txt"
        p.write_bytes(content)
        r, parsed = ora.deserialize(p)
        assert r == err
        out = subprocess.run([str(exe), str(p)], capture_output=True, text=True, check=True).stdout.split()
        got = [tuple(int(x) for x in t.split(",")) for t in out[3:]] if err < 0 else [tuple(int(x) for x in t.split(",")) for t in out[2:]]
        if err >= 0:
            assert out[0] == "OK" and got == parsed, (content, out)
        else:
            assert out[0] == "ERR" and name[err] in out[1] and got == parsed, (content, out)
