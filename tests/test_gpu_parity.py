"""Parity tests proper: the CUDA path through the C ABI (libbpe_b200.so on cuda:0) against the
oracle on the same inputs, bit-exact (integer / byte / index work), plus size-independent
properties at BASELINE.json's C2 size. Reference tests restated: basic_tokenizer.zig:351-461."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN, merges_array
from test_oracle import MAIN_ZIG_STRING, MAIN_ZIG_TOKENS, REF_TEST_MERGES

pytestmark = pytest.mark.gpu


def _opts(eng, **opts):
    for k, v in {"verify_recount": 0, "check_tiebreak": 0, "force_slow_tiebreak": 0, "compact_pct": 93, "table_log2": 0,
                 "max_steps": 0, "merge_impl": 0, **opts}.items():
        eng.set_option(k, v)


def _train_check(gpu, ora, data, vocab, **opts):
    _opts(gpu, **opts)
    m, c = gpu.train(data, vocab)
    om, oc = ora.train(data, vocab, fast=True)
    assert np.array_equal(merges_array(m), om)
    assert np.array_equal(c, oc)
    _opts(gpu)
    return gpu.last_stats


# ---- the reference's own unit tests, through the host mirror --------------------------------
def test_ref_train(zb, gpu):  # :399-432
    tk = zb.BasicTokenizer(engine=gpu, quiet=True)
    tk.train("hello world hello", 300, True)
    assert len(tk.merges) > 0
    enc = tk.encode("hello")
    assert len(enc) == 1 and enc[0] == 259
    assert tk.decode(enc) == b"hello"
    assert len(tk.merges) == 12  # early stop once one token is left (:188-191)
    with pytest.raises(zb.InvalidVocabSize):
        tk.train("abc", 255)


def test_ref_encode_decode(zb, gpu):  # :362-397
    tk = zb.BasicTokenizer(engine=gpu, quiet=True)
    for a, b, c in REF_TEST_MERGES:
        tk.put((a, b), c)
    want = [257, ord("l"), ord("o"), ord(" "), 258, ord("r"), ord("l"), ord("d")]
    assert list(tk.encode("hello world")) == want
    assert tk.decode(np.array(want, dtype=np.uint16)) == b"hello world"
    with pytest.raises(zb.InvalidToken):
        tk.decode([300])


def test_ref_serialize_roundtrip(zb, gpu, tmp_path):  # :434-461
    tk = zb.BasicTokenizer(engine=gpu, quiet=True)
    for a, b, c in REF_TEST_MERGES:
        tk.put((a, b), c)
    p = tmp_path / "test_merges.txt"
    tk.serializeMerges(p)
    tk2 = zb.BasicTokenizer(engine=gpu, quiet=True)
    tk2.deserializeMerges(p)
    assert tk2.merges == tk.merges


def test_verbose_output_format(zb, gpu, capsys):
    tk = zb.BasicTokenizer(engine=gpu)
    tk.train("hello world hello", 300, True)
    err = capsys.readouterr().err
    assert "merge 1/44: (101,108) -> 256 had 2 occurrences" in err  # printMergeInfo :308-317
    assert "No more pairs to merge. Stopping early." in err  # :189
    assert "Time statistics:" in err and "sortCodePointPairs:" in err
    assert "generateInitialTokens runtime: " in err and " seconds" in err  # :156-160


# ---- golden vectors ---------------------------------------------------------------------------
def test_golden_merges_txt(zb, gpu, taylor, tmp_path):
    """main.zig: train(taylorswift.txt, 300) + serializeMerges == the committed merges.txt."""
    tk = zb.BasicTokenizer(engine=gpu, quiet=True)
    tk.train(taylor, 300, False)
    p = tmp_path / "merges.txt"
    tk.serializeMerges(p)
    assert p.read_bytes() == open(os.path.join(GOLDEN, "merges_300.txt"), "rb").read()
    enc = tk.encode(MAIN_ZIG_STRING)
    assert list(enc) == MAIN_ZIG_TOKENS
    assert tk.decode(enc) == MAIN_ZIG_STRING


def test_c1_taylorswift_512(gpu, ora, taylor):
    """BASELINE config 1: vocab 512, then encode / decode round trip of the whole file."""
    _opts(gpu, check_tiebreak=1)
    m, c = gpu.train(taylor, 512)
    _opts(gpu)
    text = "".join(f"{a},{b},{t}\n" for a, b, t in merges_array(m)).encode()
    assert hashlib.sha256(text).hexdigest() == "fb698f9c0610b9779d549a59162c3d862ba0f9180a5687da4ce8c6e96dc9f588"
    assert gpu.last_stats["tie_steps"] == 69
    om, oc = ora.train(taylor, 512, fast=True)
    assert np.array_equal(c, oc)
    ids = gpu.encode(taylor, m)
    assert np.array_equal(ids, ora.encode(taylor, om))
    assert len(ids) == 78746  # Appendix B: final length after 256 merges
    assert gpu.decode(ids, m) == taylor


def test_whole_file_encode_golden(gpu, golden_merges, taylor):
    ids = gpu.encode(taylor, golden_merges)
    assert len(ids) == 128451
    assert hashlib.sha256(ids.astype("<u2").tobytes()).hexdigest() == "71f70e539b770e3cf87adb5b9e04bf092f9b38d767351b7954a81d709c4facc3"
    assert gpu.decode(ids, golden_merges) == taylor


# ---- edge cases -------------------------------------------------------------------------------
@pytest.mark.parametrize("data,vocab", [
    (b"", 300), (b"x", 300), (b"xy", 300), (b"abcabc", 256), (b"a" * 100000, 280), (b"ab" * 50000 + b"a", 280),
    (b"aaab" * 30000 + b"aa", 290), (bytes(range(256)) * 64, 300), (b"abc" * 4096 + b"ab", 270),
])
def test_train_edge_cases(gpu, ora, data, vocab):
    _train_check(gpu, ora, data, vocab, verify_recount=1, check_tiebreak=1)


def test_train_long_runs_across_tiles(gpu, ora):
    # runs of one byte spanning many 4096-slot tiles, with odd/even lengths and odd offsets
    data = b"xyz" + b"a" * 40961 + b"b" + b"a" * 8192 + b"cc" + b"a" * 12287 + b"q"
    _train_check(gpu, ora, data, 300, verify_recount=1, check_tiebreak=1)
    _train_check(gpu, ora, data, 300, compact_pct=0)


def test_train_random(gpu, ora):
    rng = np.random.default_rng(21)
    for k, n, vocab in [(2, 50000, 300), (4, 100000, 400), (26, 200000, 600), (256, 300000, 700)]:
        data = bytes(rng.integers(0, k, size=n, dtype=np.uint8))
        _train_check(gpu, ora, data, vocab, check_tiebreak=1)


def test_train_tiebreak_paths_agree(gpu, ora):
    rng = np.random.default_rng(22)
    data = bytes(rng.integers(97, 123, size=150000, dtype=np.uint8))
    st_fast = _train_check(gpu, ora, data, 500)
    st_slow = _train_check(gpu, ora, data, 500, force_slow_tiebreak=1)
    assert st_fast["tie_steps"] == st_slow["tie_steps"] > 0
    assert st_slow["tie_slow_steps"] == st_slow["tie_steps"]
    assert st_fast["tie_slow_steps"] < st_fast["tie_steps"]


def test_train_compaction_and_table_growth(gpu, ora, synth):
    data = bytes(synth.generate(1_500_000, synth.SEED_C3, synth.BYTE))
    st = _train_check(gpu, ora, data, 700, table_log2=12, compact_pct=95, verify_recount=0)
    assert st["compactions"] >= 3
    _train_check(gpu, ora, data[:300000], 400, compact_pct=0, verify_recount=1)


def test_train_synth_utf8_4mb(gpu, ora, synth):
    """C2-shaped input at a size the oracle finishes in seconds."""
    data = bytes(synth.generate(4_000_000, synth.SEED_C2, synth.UTF8))
    st = _train_check(gpu, ora, data, 256 + 400)
    assert st["scanned_slots"] > 0 and st["kernel_launches"] > 400


def test_encode_parity_synth(gpu, ora, synth):
    data = bytes(synth.generate(1_000_000, synth.SEED_C3, synth.BYTE))
    m, _ = gpu.train(data, 256 + 300)
    ids = gpu.encode(data, m)
    assert np.array_equal(ids, ora.encode(data, merges_array(m), linear=True))
    assert np.array_equal(gpu.encode(data[:3000], m), ora.encode(data[:3000], merges_array(m), linear=False))
    assert gpu.decode(ids, m) == data
    # text the merges were not trained on, including bytes they never saw
    other = bytes(synth.generate(500_000, synth.SEED_C5, synth.BYTE))
    ids2 = gpu.encode(other, m)
    assert np.array_equal(ids2, ora.encode(other, merges_array(m), linear=True))
    assert gpu.decode(ids2, m) == other


@pytest.mark.parametrize("data,merges", [
    (b"a" * 100001, [(97, 97, 256), (256, 256, 257), (257, 97, 258)]),
    (b"Xbbbbb cXbb" * 999, [(ord("X"), ord("b"), ord("X"))]),
    (b"a" * 9999, [(97, 97, 97)]),
    (b"aaXaX" * 5000, [(97, ord("X"), ord("X"))]),
    (b"abcabcab" * 3000, [(97, 98, 65535), (65535, 99, 300), (300, 65535, 301)]),
    (b"", [(97, 98, 256)]), (b"z", [(97, 98, 256)]), (b"abab", []),
])
def test_encode_edge_cases(gpu, ora, data, merges):
    ids = gpu.encode(data, merges)
    assert np.array_equal(ids, ora.encode(data, merges, linear=True))
    if len(ids):
        rc, want = ora.decode(ids, merges, cap=len(data) + 64)
        assert rc == 0 and gpu.decode(ids, merges) == want


def test_encode_levels_vs_single_passes(gpu, ora, synth):
    """the level-scheduled encode (default) against one pass per merge (encode_impl = 1) and the oracle, on
    a list long enough for ~50 levels, text with and without the training corpus"""
    data = bytes(synth.generate(3_000_000, synth.SEED_C3, synth.BYTE))
    m, _ = gpu.train(data, 256 + 1500)
    other = bytes(synth.generate(2_000_000, synth.SEED_C4, synth.BYTE)) + data[:100_000]
    want = ora.encode(other, merges_array(m), linear=True)
    try:
        for impl in (0, 1):
            gpu.set_option("encode_impl", impl)
            ids = gpu.encode(other, m)
            launches = gpu.last_stats["kernel_launches"]
            assert np.array_equal(ids, want), impl
            if impl == 0:
                lvl_launches = launches
        assert lvl_launches * 4 < launches  # far fewer passes
    finally:
        gpu.set_option("encode_impl", 0)
    assert gpu.decode(ids, m) == other


@pytest.mark.parametrize("seed", range(8))
def test_encode_arbitrary_lists(gpu, ora, seed):
    """merge lists no training run writes (repeated pairs, ids reused, id 65535, new id equal to a component)"""
    rng = np.random.default_rng(900 + seed)
    nsym = int(rng.integers(2, 7))
    pool = list(range(97, 97 + nsym))
    merges = []
    for k in range(int(rng.integers(4, 60))):
        a, b = int(rng.choice(pool)), int(rng.choice(pool))
        mode = rng.random()
        z = 256 + k if mode < 0.75 else int(rng.choice(pool)) if mode < 0.85 else 256 + int(rng.integers(0, k + 1)) if mode < 0.95 \
            else (65535 if rng.random() < 0.5 else a)
        merges.append((a, b, z))
        if z not in pool and rng.random() < 0.8:
            pool.append(z)
    data = bytes(rng.integers(97, 97 + nsym, size=int(rng.integers(20_000, 60_000)), dtype=np.uint8))
    assert np.array_equal(gpu.encode(data, merges), ora.encode(data, merges, linear=False))


def test_tile_encoder_on_gpu(gpu, ora, synth, taylor, golden_merges):
    """the tile-resident kernel (encode_impl = 4) against the oracle: the golden list, a trained list on unseen text, every
    tile size, runs of equal tokens, odd sizes; and the default dispatch must pick it for regular lists"""
    def enc(data, merges, tile=8192, impl=4):
        try:
            gpu.set_option("encode_impl", impl)
            gpu.set_option("encode_tile", tile)
            return gpu.encode(data, merges)
        finally:
            gpu.set_option("encode_impl", 0)
            gpu.set_option("encode_tile", 8192)
    want = ora.encode(taylor, golden_merges, linear=True)
    for tile in (512, 2048, 8192):
        assert np.array_equal(enc(taylor, golden_merges, tile), want)
    try:
        gpu.set_option("encode_try_tiles", 2)  # encode_impl 0 picks by cost; 2: the tile kernel first whatever the list
        ids = gpu.encode(taylor, golden_merges)
        assert np.array_equal(ids, want) and gpu.last_stats["kernel_calls"][10] == 2 and gpu.last_stats["kernel_calls"][11] == 1
    finally:
        gpu.set_option("encode_try_tiles", 1)
    train = bytes(synth.generate(2_000_000, synth.SEED_C3, synth.BYTE))
    om, _ = ora.train(train, 256 + 2000, fast=True)
    other = bytes(synth.generate(3_000_000, synth.SEED_C5, synth.BYTE))
    want = ora.encode(other, om, linear=True)
    for tile in (1024, 4096, 8192):
        assert np.array_equal(enc(other, om, tile), want), tile
    for n in (1, 2, 65, 8191, 8192, 8193, 12289, 100_001):
        assert np.array_equal(enc(other[:n], om), ora.encode(other[:n], om, linear=True)), n
    merges = [(97, 97, 256), (256, 256, 257), (98, 257, 258), (257, 257, 259)]
    rng = np.random.default_rng(5)
    data = b"".join(b"a" * int(rng.integers(1, 12)) + bytes(rng.integers(98, 102, size=int(rng.integers(1, 6)), dtype=np.uint8)) for _ in range(40000))
    assert np.array_equal(enc(data, merges), ora.encode(data, merges, linear=False))
    data = b"q" + b"a" * 50001 + b"b"  # a run no window can take: the dispatch falls back to the other encoders
    try:
        gpu.set_option("encode_try_tiles", 2)
        ids = gpu.encode(data, merges)
        assert np.array_equal(ids, ora.encode(data, merges, linear=False)) and gpu.last_stats["kernel_calls"][11] == 2
    finally:
        gpu.set_option("encode_try_tiles", 1)


def test_streaming_encode_from_host_buffers(gpu, ora, synth):
    """bpe_encode on an input of several chunks (stream_chunk_bytes lowered to 4 MiB): chunks travel host -> device while the tile
    kernel encodes the previous one; same ids as the resident path and as the oracle on slices"""
    train = bytes(synth.generate(2_000_000, synth.SEED_C3, synth.BYTE))
    om, _ = ora.train(train, 256 + 1500, fast=True)
    data = synth.generate(40_000_003, synth.SEED_C5, synth.BYTE)
    try:
        gpu.set_option("stream_chunk_bytes", 4 << 20)
        ids = gpu.encode(data, om)
        st = dict(gpu.last_stats)
        assert st["kernel_calls"][10] == 2 and st["kernel_calls"][11] == 1 and st["kernel_calls"][9] == 10
    finally:
        gpu.set_option("stream_chunk_bytes", 0)
    resident = gpu.encode(data[:30_000_000], om)  # one chunk: the resident path
    assert gpu.last_stats["kernel_calls"][9] == 0
    k = len(resident) - 100
    assert np.array_equal(ids[:k], resident[:k])
    for lo in (0, 4 * (1 << 20) - 300, 17_000_000):
        sl = data[lo:lo + 200_000]
        want = ora.encode(sl, om, linear=True)
        # the slice's own encoding agrees with the whole text's away from the slice ends: compare through decode
        assert gpu.decode(want, om) == sl.tobytes()
    assert gpu.decode(ids, om) == data.tobytes()


def test_count_overflow_is_detected_on_gpu(gpu, zb, taylor):
    try:
        gpu.set_option("count_limit_log2", 9)
        with pytest.raises(zb.BpeError) as e:
            gpu.train(taylor, 300)
        assert e.value.code == zb.BPE_ERR_INTERNAL
    finally:
        gpu.set_option("count_limit_log2", 32)
    m, c = gpu.train(taylor, 300)
    assert len(m) == 44


def test_decode_semantics(gpu, zb):
    with pytest.raises(zb.InvalidToken):
        gpu.decode([256], [(256, 97, 256)])  # cyclic definition (stack overflow in the reference)
    assert gpu.decode([257], [(97, 98, 257), (99, 99, 257)]) == b"ab"  # first matching merge wins (:109-116)
    assert gpu.decode([], REF_TEST_MERGES) == b""
    assert gpu.decode(np.arange(256, dtype=np.uint16), []) == bytes(range(256))


# ---- BASELINE configs 2 and 3 at full size: the whole merge list against the oracle's committed golden run --------
def _golden_big(cfg):
    """(merges (k,3), counts (k,), record) produced by tools/make_golden_big.py (oracle fast mode over the whole config)"""
    import json
    rec = json.load(open(os.path.join(GOLDEN, "big_sha256.json")))[cfg]
    text = open(os.path.join(GOLDEN, cfg + "_merges.txt")).read()
    assert hashlib.sha256(text.encode()).hexdigest() == rec["merges_sha256"]
    m = np.array([[int(x) for x in ln.split(",")] for ln in text.split()], dtype=np.uint16).reshape(-1, 3)
    c = np.array([int(x) for x in open(os.path.join(GOLDEN, cfg + "_counts.txt")).read().split()], dtype=np.uint64)
    assert len(m) == len(c) == rec["merges"]
    return m, c, rec


def _merges_text(ma):
    return "".join(f"{int(a)},{int(b)},{int(z)}\n" for a, b, z in ma)


def test_c2_full_size_golden(gpu, ora, synth):
    """C2: 100 MB UTF-8 corpus, vocab 4096. All 3,840 merges and their counts equal the oracle's full run
    (tests/golden/c2_merges.txt, sha256 pinned), then encode / decode at full size."""
    n = 100_000_000
    gm, gc, rec = _golden_big("c2")
    assert rec["complete"] and len(gm) == 3840
    data = synth.generate(n, synth.SEED_C2, synth.UTF8)
    assert hashlib.sha256(data.tobytes()).hexdigest() == rec["corpus_sha256"]
    m, c = gpu.train(data, 4096)
    st = dict(gpu.last_stats)
    ma = merges_array(m)
    assert np.array_equal(ma, gm), f"first differing merge: {int(np.argmax((ma != gm).any(axis=1)))}"
    assert np.array_equal(c, gc)
    assert hashlib.sha256(_merges_text(ma).encode()).hexdigest() == rec["merges_sha256"]
    # encode -> decode round trip of the whole corpus; slices against the oracle's encode
    ids = gpu.encode(data, m)
    assert len(ids) < n // 2
    assert gpu.decode(ids, m) == data.tobytes()
    for lo in (0, 12_345_678, n - 200_000):
        sl = slice(lo, lo + 200_000)
        assert np.array_equal(gpu.encode(data[sl], m), ora.encode(data[sl], ma, linear=True))
    assert st["tie_slow_steps"] <= st["tie_steps"]


def test_cpp_host_mirror_runs_main_zig_workload(zb, taylor, tmp_path):
    """The C++ host mirror (zig-bpe_b200/host) binds the same C symbols the Zig shim declares; its main.cpp is
    the reference's main.zig workload: train to vocab 300, write merges.txt, encode + decode the sample."""
    import subprocess
    from conftest import ROOT
    exe = tmp_path / "zig-bpe-cpp"
    lib_dir = os.path.dirname(zb.LIB_PATH)
    subprocess.run(["g++", "-O2", "-std=c++17", f"-I{ROOT}/include", f"{ROOT}/zig-bpe_b200/host/main.cpp", f"-L{lib_dir}",
                    "-lbpe_b200", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)], check=True)
    out = tmp_path / "merges.txt"
    r = subprocess.run([str(exe), os.path.join(GOLDEN, "taylorswift.txt"), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert out.read_bytes() == open(os.path.join(GOLDEN, "merges_300.txt"), "rb").read()
    assert " ".join(str(t) for t in MAIN_ZIG_TOKENS) in r.stderr
    assert MAIN_ZIG_STRING.decode() in r.stderr
    assert "Training completed in" in r.stderr and "Time statistics:" in r.stderr


def test_c3_full_size_golden(gpu, ora, synth):
    """BASELINE config 3 (the headline): 1 GB byte corpus, vocab 8192, on one GPU. Merges and counts equal the
    oracle's golden run (tests/golden/c3_merges.txt: the whole list when big_sha256.json says complete, else a prefix),
    then encode -> decode round trip of the whole corpus. The oracle job kept running after the last GPU lease of the
    round: `gpu_verified_merges` is how much of the list a B200 training had been compared with when it was committed
    (tools/check_c3_golden.py), and that much is asserted here; bench.py reports the comparison with the longest prefix
    hash of the whole list in its `parity` block (true / false, no assertion)."""
    n = 1_000_000_000
    gm, gc, rec = _golden_big("c3")
    k = min(len(gm), int(rec.get("gpu_verified_merges", len(gm))))
    gm, gc = gm[:k], gc[:k]
    assert k >= 4971
    data = synth.generate(n, synth.SEED_C3, synth.BYTE)
    assert hashlib.sha256(data.tobytes()).hexdigest() == rec["corpus_sha256"]
    m, c = gpu.train(data, 8192)
    st = dict(gpu.last_stats)
    ma = merges_array(m)
    assert len(m) == 7936 and list(ma[:, 2]) == list(range(256, 8192))
    assert np.array_equal(ma[:k], gm), f"first differing merge: {int(np.argmax((ma[:k] != gm).any(axis=1)))}"
    assert np.array_equal(c[:k], gc)
    assert hashlib.sha256(_merges_text(ma[:k]).encode()).hexdigest() == rec["prefix_sha256"].get(str(k), rec["merges_sha256"])
    assert (np.diff(c.astype(np.int64)) <= 0).all()
    ids = gpu.encode(data, m)
    # encode replays the training sequence: a merge removes at most its (overlapping) pair count of tokens
    assert n - int(c.astype(np.int64).sum()) <= len(ids) < n // 3
    out = gpu.decode(ids, m)
    assert len(out) == n and hashlib.sha256(out).digest() == hashlib.sha256(data.tobytes()).digest()
    sl = slice(500_000_000, 500_000_000 + 100_000)
    assert np.array_equal(gpu.encode(data[sl], m), ora.encode(data[sl], ma, linear=True))
    assert st["tie_steps"] > 0 and st["scanned_slots"] > 0


@pytest.mark.parametrize("impl", [1, 2])
def test_train_alternative_merge_paths(gpu, ora, synth, taylor, impl):
    """merge_impl = 1 (TMA ring) and 2 (candidate scan) must learn exactly what the default path learns."""
    rng = np.random.default_rng(31)
    _train_check(gpu, ora, taylor, 400, merge_impl=impl, check_tiebreak=1)
    _train_check(gpu, ora, bytes(rng.integers(97, 101, size=200000, dtype=np.uint8)), 350, merge_impl=impl, verify_recount=1)
    _train_check(gpu, ora, b"xyz" + b"a" * 40961 + b"b" + b"a" * 8192 + b"cc" + b"a" * 12287 + b"q", 300, merge_impl=impl, verify_recount=1)
    data = bytes(synth.generate(3_000_000, synth.SEED_C3, synth.BYTE))
    _train_check(gpu, ora, data, 256 + 300, merge_impl=impl)


@pytest.mark.parametrize("opts", [dict(merge_loop=0), dict(merge_loop=0, merge_prestage=0), dict(merge_loop=0, merge_direct=0),
                                  dict(merge_loop=1), dict(merge_loop=3, merge_direct=1 << 20, merge_pairfilter=0),
                                  dict(merge_loop=50, merge_direct=0), dict(merge_loop=4, merge_dbuf=1)])
def test_train_merge_pass_variants(gpu, ora, synth, taylor, opts):
    """The looped merge pass (merge_loop tiles per CTA, prefetch, optional second staging buffer), the one-CTA-per-tile pass
    with / without pre-staging, the queued and the queue-less staged path with / without the packed pair filter: all learn
    what the oracle learns, with a recount of all pairs after every step on the smaller inputs."""
    defaults = dict(merge_loop=6, merge_prestage=1, merge_direct=3, merge_pairfilter=1, merge_dbuf=0)
    rng = np.random.default_rng(32)
    try:
        _train_check(gpu, ora, taylor, 400, check_tiebreak=1, **opts)
        _train_check(gpu, ora, bytes(rng.integers(97, 101, size=200000, dtype=np.uint8)), 350, verify_recount=1, **opts)
        _train_check(gpu, ora, b"xyz" + b"a" * 40961 + b"b" + b"a" * 8192 + b"cc" + b"a" * 12287 + b"q" + b"ab" * 9000, 300,
                     verify_recount=1, **opts)
        data = bytes(synth.generate(3_000_000, synth.SEED_C3, synth.BYTE))
        _train_check(gpu, ora, data, 256 + 300, **opts)
    finally:
        _opts(gpu, **defaults)


def test_train_long_run_on_random_bytes(gpu, ora):
    """A long training on random bytes ends in the regime where the maximum is 1-2 and (hundreds of) thousands of
    pairs tie: the heavy list holds the whole table there and ties beyond 1,024 keys go to the replay."""
    rng = np.random.default_rng(41)
    data = bytes(rng.integers(0, 256, size=40000, dtype=np.uint8))
    st = _train_check(gpu, ora, data, 256 + 1500)
    assert st["tie_steps"] > 500


def test_encode_levels_repeatable_at_size(gpu, synth):
    """50 MB, 2,000 merges: the in-place level pass gives the same ids on every run (a race would show up as run-to-run
    differences) and the same ids as one pass per merge"""
    data = synth.generate(50_000_000, synth.SEED_C3, synth.BYTE)
    m, _ = gpu.train(data, 256 + 2000)
    runs = [gpu.encode(data, m) for _ in range(3)]
    assert np.array_equal(runs[0], runs[1]) and np.array_equal(runs[0], runs[2])
    try:
        gpu.set_option("encode_impl", 1)
        assert np.array_equal(gpu.encode(data, m), runs[0])
    finally:
        gpu.set_option("encode_impl", 0)
    assert gpu.decode(runs[0], m) == data.tobytes()


def test_encode_filter_variant(gpu, ora, synth):
    """encode_filter = 1 (byte role map + in-register successor filter in the level pass) gives the same ids"""
    data = bytes(synth.generate(2_000_000, synth.SEED_C3, synth.BYTE))
    m, _ = gpu.train(data, 256 + 1200)
    other = bytes(synth.generate(1_500_000, synth.SEED_C5, synth.BYTE)) + data[:50_000]
    want = ora.encode(other, merges_array(m), linear=True)
    try:
        gpu.set_option("encode_impl", 2)
        for filt in (0, 1, 2):  # role bitmap / role bytes + successor filter / pair Bloom filter
            gpu.set_option("encode_filter", filt)
            assert np.array_equal(gpu.encode(other, m), want), filt
            assert np.array_equal(gpu.encode(b"ab" * 5000 + b"a", [(97, 98, 256), (256, 256, 257), (98, 97, 258), (257, 97, 259)]),
                                  ora.encode(b"ab" * 5000 + b"a", [(97, 98, 256), (256, 256, 257), (98, 97, 258), (257, 97, 259)], linear=False)), filt
    finally:
        gpu.set_option("encode_filter", 2)
        gpu.set_option("encode_impl", 0)


@pytest.mark.xfail(strict=False, reason="bpe_groupenc.cuh was written after the round's last GPU lease: equal to the oracle under the CPU "
                                        "emulation, never run on a GPU; off by default (encode_fuse = 0)")
def test_encode_fused_level_groups_experimental():
    """LAST test of the GPU tier, in a process of its own (a fault in the new kernel must not touch the tests above):
    `encode_fuse` = 2 / 4 / 8 must give the ids of the level passes on a 50 MB corpus with 1,792 merges
    (tools/fuse_gpu_check.py prints the timings)."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(GOLDEN), "..", "tools", "fuse_gpu_check.py"), "5e7", "2048"],
                       capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
