"""The oracle against every golden vector / known answer the reference holds for this path
(SURVEY.md §8c): Wyhash vectors, the five in-file unit tests (basic_tokenizer.zig:351-461), the
committed merges.txt, main.zig's encode, plus the model-derived merges_512 (Appendix B)."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN

MAIN_ZIG_STRING = "hello world!!!? (안녕하세요!) lol123 😉".encode()
MAIN_ZIG_TOKENS = [104, 101, 108, 108, 111, 32, 119, 291, 108, 100, 33, 33, 33, 63, 32, 40, 236, 149, 136, 235, 133, 149,
                   237, 149, 152, 236, 132, 184, 236, 154, 148, 33, 41, 32, 108, 111, 108, 49, 50, 51, 32, 240, 159, 152, 137]
REF_TEST_MERGES = [(ord("h"), ord("e"), 256), (256, ord("l"), 257), (ord("w"), ord("o"), 258)]


def test_wyhash_vectors(ora):  # Zig std/hash/wyhash.zig test vectors
    vec = [(0, b"", 0x0409638EE2BDE459), (1, b"a", 0xA8412D091B5FE0A9), (2, b"abc", 0x32DD92E4B2915153),
           (3, b"message digest", 0x8619124089A3A16B), (4, b"abcdefghijklmnopqrstuvwxyz", 0x7A43AFB61D7F5F40),
           (5, b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789", 0xFF42329B90E50D58),
           (6, b"1234567890" * 8, 0xC39CAB13B115AAD3)]
    for seed, data, want in vec:
        assert ora.wyhash(seed, data) == want, (seed, data)


def test_pair_hash_samples(ora):  # Appendix A.2
    samples = {(101, 32): 0x544D3DA525A0DEDA, (46, 10): 0xDA1944F769904CC9, (265, 101): 0x3A6CB6F333C4024E,
               (0, 0): 0x14B83016DC460955, (65535, 65535): 0x53AC8FA6824BCAD0, (97, 97): 0x49A7A1E6A8804B35}
    for (a, b), want in samples.items():
        assert ora.hash_pair(a, b) == want
    assert ora.hash_pair(46, 10) & 8191 == 3273 and ora.hash_pair(265, 101) & 8191 == 590


def test_zig_map_growth_thresholds(ora):  # Appendix A.4: capacity doubles on the getOrPut after 6, 12, 25, 51, ... keys
    def cap_after(n, extra_call):
        pairs = [(i, i + 1) for i in range(n)]
        if extra_call:
            pairs.append(pairs[0])  # an existing key still triggers growIfNeeded
        return ora.zigmap_slots(pairs)[0]
    for n, cap in [(1, 8), (6, 8), (7, 16), (12, 16), (13, 32), (25, 32), (26, 64), (51, 64), (52, 128)]:
        assert cap_after(n, False) == cap, n
    assert cap_after(6, True) == 16 and cap_after(12, True) == 32 and cap_after(5, True) == 8


def test_ref_generate_initial_tokens(zb):  # :351-360
    t = zb.BasicTokenizer.generateInitialTokens(None, "hello world")
    assert list(t) == [ord(c) for c in "hello world"] and t.dtype == np.uint16


def test_ref_encode(ora):  # :362-378
    want = [257, ord("l"), ord("o"), ord(" "), 258, ord("r"), ord("l"), ord("d")]
    assert list(ora.encode(b"hello world", REF_TEST_MERGES, linear=False)) == want
    assert list(ora.encode(b"hello world", REF_TEST_MERGES, linear=True)) == want


def test_ref_decode(ora):  # :380-397
    rc, out = ora.decode([257, ord("l"), ord("o"), ord(" "), 258, ord("r"), ord("l"), ord("d")], REF_TEST_MERGES)
    assert rc == 0 and out == b"hello world"
    assert ora.decode([300], REF_TEST_MERGES)[0] == 3  # error.InvalidToken


def test_ref_train(ora):  # :399-432
    for fast in (False, True):
        m, c = ora.train(b"hello world hello", 300, fast=fast)
        assert len(m) == 12  # collapses to one token, then "No more pairs to merge"
        assert [tuple(r) for r in m] == [(101, 108, 256), (104, 256, 257), (257, 108, 258), (258, 111, 259), (119, 111, 260),
                                         (32, 260, 261), (261, 114, 262), (100, 32, 263), (259, 262, 264), (264, 108, 265),
                                         (265, 263, 266), (266, 259, 267)]
        enc = ora.encode(b"hello", m)
        assert list(enc) == [259]
        assert ora.decode(enc, m) == (0, b"hello")
    with pytest.raises(ValueError):
        ora.train(b"abc", 255)  # InvalidVocabSize (:147-149)


def test_ref_serialize_roundtrip(ora, tmp_path):  # :434-461
    p = tmp_path / "test_merges.txt"
    assert ora.serialize(REF_TEST_MERGES, p) == 0
    assert p.read_bytes() == b"104,101,256\n256,108,257\n119,111,258\n"
    n, back = ora.deserialize(p)
    assert n == 3 and back == REF_TEST_MERGES


def test_golden_merges_txt(ora, taylor):
    """main.zig workload: train(taylorswift.txt, 300) must reproduce the committed merges.txt byte for byte."""
    for fast in (False, True):
        m, _ = ora.train(taylor, 300, fast=fast)
        text = "".join(f"{a},{b},{c}\n" for a, b, c in m).encode()
        assert text == open(os.path.join(GOLDEN, "merges_300.txt"), "rb").read()
        assert hashlib.sha256(text).hexdigest() == "f1a9b78b2be24bf3c6813cb0efd4920f0e40347ac7da7a098df61bd215d5f8d0"


def test_merges_512_model(ora, taylor):  # Appendix B (model-derived; C1 of BASELINE.md)
    m, c = ora.train(taylor, 512, fast=False)
    text = "".join(f"{a},{b},{cc}\n" for a, b, cc in m).encode()
    assert len(m) == 256
    assert hashlib.sha256(text).hexdigest() == "fb698f9c0610b9779d549a59162c3d862ba0f9180a5687da4ce8c6e96dc9f588"
    mf, cf = ora.train(taylor, 512, fast=True)
    assert np.array_equal(m, mf) and np.array_equal(c, cf)


def test_main_zig_encode(ora, golden_merges, taylor):
    enc = ora.encode(MAIN_ZIG_STRING, golden_merges)
    assert list(enc) == MAIN_ZIG_TOKENS
    assert ora.decode(enc, golden_merges) == (0, MAIN_ZIG_STRING)
    whole = ora.encode(taylor, golden_merges)
    assert len(whole) == 128451
    assert hashlib.sha256(whole.astype("<u2").tobytes()).hexdigest() == "71f70e539b770e3cf87adb5b9e04bf092f9b38d767351b7954a81d709c4facc3"
    assert ora.decode(whole, golden_merges, cap=len(taylor) + 16) == (0, taylor)


def test_encode_verbatim_vs_linear(ora):
    rng = np.random.default_rng(7)
    for trial in range(20):
        data = bytes(rng.integers(97, 100, size=int(rng.integers(1, 200)), dtype=np.uint8))
        merges = [(97, 97, 256), (256, 98, 257), (98, 98, 258), (257, 257, 259), (99, 97, 99)][: int(rng.integers(1, 6))]
        assert np.array_equal(ora.encode(data, merges, linear=False), ora.encode(data, merges, linear=True))
    # a merge whose new token equals its first component keeps absorbing (i is not advanced, :78-81)
    assert list(ora.encode(b"Xbbbb", [(ord("X"), ord("b"), ord("X"))], linear=False)) == [ord("X")]
    assert list(ora.encode(b"aaa", [(97, 97, 256)], linear=False)) == [256, 97]
    assert list(ora.encode(b"aaaa", [(97, 97, 256)], linear=False)) == [256, 256]


def test_fast_equals_faithful_random(ora):
    rng = np.random.default_rng(11)
    for k in (2, 3, 5, 40, 256):
        data = bytes(rng.integers(0, k, size=4000, dtype=np.uint8))
        a = ora.train(data, 256 + 60, fast=False)
        b = ora.train(data, 256 + 60, fast=True)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    for data in (b"a" * 777, b"ab" * 400 + b"a", b"abc" * 300, b"aab" * 300 + b"aa"):
        a = ora.train(data, 300, fast=False)
        b = ora.train(data, 300, fast=True)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("content,err", [
    (b"1,2,3\n4,5,6", 2), (b"1,2,3,99\n", 1), (b"1,2\n", -106), (b"1,2,\n", -20), (b"\n", -20), (b"1,2,3\r\n", -20),
    (b"1,2,65536\n", -21), (b"1, 2,3\n", -20), (b"+1,2,3\n", 1), (b"1_0,2,3\n", 1), (b"1,2,3\n" + b"9" * 100 + b"\n", -22),
    (b"1,2,3\n" + b"7," * 49 + b"7\n", 2),
])
def test_deserialize_edge_cases(ora, zb, tmp_path, content, err):
    """Format limits of deserializeMerges (:332-348): 100-byte line buffer, parseInt(u16), field count.
    The host mirror (zig-bpe_b200.read_merges) must agree with the oracle."""
    p = tmp_path / "m.txt"
    p.write_bytes(content)
    r, parsed = ora.deserialize(p)
    assert r == err
    got = []
    name = {-20: "InvalidCharacter", -21: "Overflow", -22: "StreamTooLong", -106: "InvalidFormat"}
    if err >= 0:
        zb.read_merges(p, got)
        assert got == parsed and len(got) == err
    else:
        with pytest.raises(zb.MergesFormatError) as ei:
            zb.read_merges(p, got)
        assert ei.value.zig_error == name[err]
        assert got == parsed  # merges before the bad line stay appended
