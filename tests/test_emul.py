"""Kernel logic without a GPU: the same .cu sources compiled for the CPU emulation in
tests/emul (small tiles so tile-boundary code paths are hit by tiny inputs), compared with the
oracle. This tier only checks logic; the parity tests proper are the -m gpu tests."""
import numpy as np
import pytest

from conftest import merges_array
from test_oracle import MAIN_ZIG_STRING, MAIN_ZIG_TOKENS, REF_TEST_MERGES


def _train_check(emu, ora, data, vocab, **opts):
    for k, v in {"verify_recount": 0, "check_tiebreak": 1, "force_slow_tiebreak": 0, "compact_pct": 93, "merge_impl": 0, **opts}.items():
        emu.set_option(k, v)
    m, c = emu.train(data, vocab)
    om, oc = ora.train(data, vocab, fast=True)
    assert np.array_equal(merges_array(m), om)
    assert np.array_equal(c, oc)
    return emu.last_stats


@pytest.mark.parametrize("data,vocab", [
    (b"hello world hello", 300), (b"a" * 1000, 270), (b"ab" * 700 + b"a", 270), (b"aaab" * 300 + b"aa", 280),
    (b"x", 300), (b"xy", 300), (b"", 300), (b"abcabc", 256),
])
def test_train_small(emu, ora, data, vocab):
    _train_check(emu, ora, data, vocab, verify_recount=1)


def test_train_random_alphabets(emu, ora):
    rng = np.random.default_rng(3)
    _train_check(emu, ora, bytes(rng.integers(97, 101, size=3000, dtype=np.uint8)), 290, verify_recount=1)
    st = _train_check(emu, ora, bytes(rng.integers(0, 256, size=4000, dtype=np.uint8)), 330)
    assert st["tie_steps"] > 0


def test_train_forced_replay_and_no_compaction(emu, ora):
    rng = np.random.default_rng(5)
    data = bytes(rng.integers(97, 103, size=2500, dtype=np.uint8))
    st = _train_check(emu, ora, data, 300, force_slow_tiebreak=1, compact_pct=0)
    assert st["compactions"] == 0 and st["tie_slow_steps"] == st["tie_steps"]
    st = _train_check(emu, ora, data, 300, compact_pct=100)
    assert st["compactions"] > 0


@pytest.mark.parametrize("loop,direct,pairfilter,dbuf", [(0, 3, 1, 0), (1, 3, 1, 0), (2, 0, 1, 1), (3, 1 << 20, 0, 0), (7, 3, 1, 0), (1000, 3, 1, 0),
                                                        (3, 3, 1, 1), (2, 1 << 20, 1, 1), (1000, 3, 1, 1)])
def test_train_resident_merge_ctas(emu, ora, taylor, loop, direct, pairfilter, dbuf):
    """merge_loop = tiles per CTA of the looped merge pass (a CTA takes every gridDim-th tile, the next tile and its halo are
    prefetched, the delta bins live as long as the CTA); 1 = one tile per CTA, 1000 = one CTA takes them all; merge_dbuf = two staging buffers. With a
    recount of all pairs after every step."""
    rng = np.random.default_rng(12)
    try:
        for data, vocab in ((taylor[:14000], 300), (bytes(rng.integers(97, 100, size=5000, dtype=np.uint8)), 290),
                            (b"abab" * 300 + b"ba" * 200 + b"aab" * 100 + b"a" * 700, 275)):
            _train_check(emu, ora, data, vocab, verify_recount=1, merge_loop=loop, merge_direct=direct, merge_pairfilter=pairfilter,
                         merge_dbuf=dbuf)
    finally:
        for k, v in (("merge_loop", 6), ("merge_direct", 3), ("merge_pairfilter", 1), ("merge_dbuf", 0)):
            emu.set_option(k, v)


@pytest.mark.parametrize("direct", [0, 8, 1 << 20])
def test_train_queued_and_queueless_step_paths(emu, ora, taylor, direct):
    """merge_direct: 0 = every staged tile fills the candidate queue, 2^20 = no step does, 8 = the default mix (by the
    pair's count per tile); with a recount of all pairs after every step"""
    rng = np.random.default_rng(11)
    try:
        for data, vocab in ((taylor[:12000], 300), (bytes(rng.integers(97, 100, size=4000, dtype=np.uint8)), 290),
                            (b"abab" * 300 + b"ba" * 200 + b"aab" * 100, 275)):
            _train_check(emu, ora, data, vocab, verify_recount=1, merge_direct=direct)
    finally:
        emu.set_option("merge_direct", 3)


def test_train_golden_prefix(emu, ora, taylor):
    st = _train_check(emu, ora, taylor[:40000], 300)
    assert st["kernel_launches"] > 0


def test_count_overflow_is_detected(emu, zb, ora, taylor):
    """The pair table counts in 32 bits (the reference in usize, :47,265): a corpus whose most frequent byte pair does not
    fit is refused with BPE_ERR_INTERNAL instead of wrapping silently. `count_limit_log2` scales the limit down so that
    the check runs on a small input; at the real limit (2^32 - 1) the same input trains as usual."""
    data = taylor[:30000]
    try:
        emu.set_option("count_limit_log2", 9)  # ('e',' ') occurs far more than 511 times in 30 KB of English
        with pytest.raises(zb.BpeError) as e:
            emu.train(data, 270)
        assert e.value.code == zb.BPE_ERR_INTERNAL and "32-bit pair counts" in str(e.value)
        emu.set_option("count_limit_log2", 12)  # 4,095: every pair fits
        m, c = emu.train(data, 270)
        assert int(c.max()) <= 4095
    finally:
        emu.set_option("count_limit_log2", 32)
    m2, c2 = emu.train(data, 270)
    assert np.array_equal(m, m2) and np.array_equal(c, c2)


def test_fused_halo_equals_a_halo_pass_per_step(emu, ora, taylor):
    """fuse_halo = 1 (default: the apply kernel gathers the next step's halos, first == second steps take the
    run-chaining halo pass through a halt) learns what fuse_halo = 0 (one halo launch per step) learns"""
    rng = np.random.default_rng(3)
    for data, vocab in ((taylor[:20000], 330), (bytes(rng.integers(97, 100, size=5000, dtype=np.uint8)), 300),
                        (b"xyz" + b"a" * 700 + b"b" + b"a" * 513 + b"cc" + b"a" * 1024 + b"q", 280)):
        om, oc = ora.train(data, vocab, fast=True)
        for fh in (1, 0):
            try:
                for k, v in {"verify_recount": 0, "check_tiebreak": 0, "force_slow_tiebreak": 0, "compact_pct": 93, "merge_impl": 0}.items():
                    emu.set_option(k, v)  # (ties settled by the replay do not pass through the first == second halt)
                emu.set_option("fuse_halo", fh)
                m, c = emu.train(data, vocab)
                assert np.array_equal(merges_array(m), om) and np.array_equal(c, oc), fh
                if fh:
                    assert emu.last_stats["aeqb_steps"] + emu.last_stats["tie_slow_steps"] >= int((om[:, 0] == om[:, 1]).sum()) >= emu.last_stats["aeqb_steps"]
            finally:
                emu.set_option("fuse_halo", 1)


def test_invalid_vocab(emu, zb):
    with pytest.raises(zb.InvalidVocabSize):
        emu.train(b"abc", 255)


def test_encode_decode(emu, ora, golden_merges, taylor):
    want = [257, ord("l"), ord("o"), ord(" "), 258, ord("r"), ord("l"), ord("d")]
    assert list(emu.encode(b"hello world", REF_TEST_MERGES)) == want  # reference test :362-378
    assert emu.decode(want, REF_TEST_MERGES) == b"hello world"  # :380-397
    assert list(emu.encode(MAIN_ZIG_STRING, golden_merges)) == MAIN_ZIG_TOKENS
    assert emu.decode(MAIN_ZIG_TOKENS, golden_merges) == MAIN_ZIG_STRING
    chunk = taylor[:6000]
    ids = emu.encode(chunk, golden_merges)
    assert np.array_equal(ids, ora.encode(chunk, golden_merges, linear=False))
    assert emu.decode(ids, golden_merges) == chunk


@pytest.mark.parametrize("data,merges", [
    (b"a" * 1001, [(97, 97, 256), (256, 256, 257), (257, 97, 258)]),
    (b"Xbbbbb cXbb", [(ord("X"), ord("b"), ord("X"))]),
    (b"aaaaaaa", [(97, 97, 97)]),
    (b"aaXaX", [(97, ord("X"), ord("X"))]),
    (b"abcabcab", [(97, 98, 65535), (65535, 99, 300), (300, 65535, 301)]),
    (b"", [(97, 98, 256)]), (b"z", [(97, 98, 256)]), (b"abab", []),
])
def test_encode_edge_cases(emu, ora, data, merges):
    ids = emu.encode(data, merges)
    assert np.array_equal(ids, ora.encode(data, merges, linear=False))
    if len(ids):
        rc, want = ora.decode(ids, merges)
        assert rc == 0 and emu.decode(ids, merges) == want


def test_decode_errors(emu, zb):
    with pytest.raises(zb.InvalidToken):
        emu.decode([300], REF_TEST_MERGES)
    with pytest.raises(zb.InvalidToken):  # a cycle would overflow the reference's stack; defined as InvalidToken
        emu.decode([256], [(256, 97, 256)])
    assert emu.decode([257], [(97, 98, 257), (99, 99, 257)]) == b"ab"  # first matching merge wins (:109-116)
    assert emu.decode([], REF_TEST_MERGES) == b""


@pytest.mark.parametrize("data,vocab", [
    (b"hello world hello", 300), (b"ab" * 700 + b"a", 270), (b"aaab" * 300 + b"aa", 280), (b"a" * 1000, 266),
])
def test_train_candidate_scan_path(emu, ora, data, vocab):
    """merge_impl = 2: scan / resolve / write kernels instead of halo + tiled merge (A == B and dense steps
    fall back to the tiled pass through H_CLASSIC)."""
    try:
        _train_check(emu, ora, data, vocab, verify_recount=1, merge_impl=2)
        _train_check(emu, ora, data, vocab, merge_impl=2)
    finally:
        emu.set_option("merge_impl", 0)


def test_train_candidate_scan_path_random(emu, ora, taylor):
    rng = np.random.default_rng(9)
    try:
        _train_check(emu, ora, bytes(rng.integers(97, 101, size=3000, dtype=np.uint8)), 290, verify_recount=1, merge_impl=2)
        _train_check(emu, ora, bytes(rng.integers(0, 256, size=4000, dtype=np.uint8)), 330, merge_impl=2)
        _train_check(emu, ora, taylor[:30000], 300, merge_impl=2)
    finally:
        emu.set_option("merge_impl", 0)


def test_train_until_everything_ties(emu, ora):
    """Late phase of a long training: the maximum count drops to 1 and every remaining pair ties (the heavy
    list then holds the whole table and the tie goes to the replay)."""
    rng = np.random.default_rng(17)
    data = bytes(rng.integers(0, 256, size=1200, dtype=np.uint8))
    st = _train_check(emu, ora, data, 256 + 120, check_tiebreak=0)
    assert st["tie_steps"] > 100 and st["tie_slow_steps"] > 0  # > 1,024 tied pairs go to the replay


def _encode_both(emu, ora, data, merges):
    """level-scheduled encode (default) and one-pass-per-merge encode against the oracle's verbatim loop"""
    want = ora.encode(data, merges, linear=False)
    try:
        for impl, filt in ((2, 0), (2, 1), (2, 2), (1, 0)):
            emu.set_option("encode_impl", impl)
            emu.set_option("encode_filter", filt)  # 1: byte role map + successor filter (MODE 1); 2: pair Bloom filter (MODE 2)
            got = emu.encode(data, merges)
            assert np.array_equal(got, want), (impl, filt, merges[:8])
        emu.set_option("encode_impl", 2)
        for fuse in (2, 3, 8):  # fused level groups (bpe_groupenc.cuh)
            emu.set_option("encode_fuse", fuse)
            got = emu.encode(data, merges)
            assert np.array_equal(got, want), ("fuse", fuse, merges[:8])
    finally:
        emu.set_option("encode_impl", 0)
        emu.set_option("encode_filter", 2)
        emu.set_option("encode_fuse", 0)
    return want


@pytest.mark.parametrize("seed", range(6))
def test_encode_levels_trained_lists(emu, ora, seed):
    """merge lists as training produces them, on other text than they were trained on (tiles of 256 slots:
    pairs straddle tile ends, levels are split by compactions)"""
    rng = np.random.default_rng(100 + seed)
    alpha = [2, 3, 5, 26, 64, 256][seed]
    train = bytes(rng.integers(0, alpha, size=3000, dtype=np.uint8))
    merges, _ = ora.train(train, 256 + [40, 60, 90, 120, 80, 40][seed])
    merges = [tuple(int(x) for x in r) for r in merges]
    text = bytes(rng.integers(0, alpha, size=2500, dtype=np.uint8)) + train[:700]
    ids = _encode_both(emu, ora, text, merges)
    assert emu.decode(ids, merges) == text


@pytest.mark.parametrize("seed", range(12))
def test_encode_levels_arbitrary_lists(emu, ora, seed):
    """lists no training run would write: repeated pairs, new ids that are bytes, that equal a component or that were
    used before, runs of equal tokens — the schedule must keep exactly the reference's sequential meaning"""
    rng = np.random.default_rng(500 + seed)
    nsym = int(rng.integers(2, 7))
    pool = list(range(97, 97 + nsym))
    merges = []
    for k in range(int(rng.integers(4, 40))):
        a, b = int(rng.choice(pool)), int(rng.choice(pool))
        mode = rng.random()
        if mode < 0.75:
            z = 256 + k
        elif mode < 0.85:
            z = int(rng.choice(pool))           # reuses an id (maybe a byte, maybe a or b)
        elif mode < 0.95:
            z = 256 + int(rng.integers(0, k + 1))  # an id an earlier merge may have produced
        else:
            z = 65535 if rng.random() < 0.5 else a
        merges.append((a, b, z))
        if z not in pool and rng.random() < 0.8:
            pool.append(z)
    data = bytes(rng.integers(97, 97 + nsym, size=int(rng.integers(300, 1500)), dtype=np.uint8))
    _encode_both(emu, ora, data, merges)


def test_encode_levels_large_level(emu, ora):
    """more pairs in one level than one pass holds (LVL_MAX = 1024): the level is split"""
    merges = [(a, b, 256 + i) for i, (a, b) in enumerate((a, b) for a in range(0, 40) for b in range(100, 130))]
    rng = np.random.default_rng(3)
    data = bytes(rng.choice(np.array(list(range(0, 40)) + list(range(100, 130)), dtype=np.uint8), size=3000))
    _encode_both(emu, ora, data, merges)


def test_encode_levels_id_65535_pair(emu, ora):
    """(65535,65535) packs to the hash's empty marker: it must never look like a hit"""
    merges = [(97, 98, 65535), (65535, 99, 300), (99, 99, 301), (65535, 97, 302)]
    data = b"ababc" * 40 + b"abab" * 30 + b"abc" * 20
    _encode_both(emu, ora, data, merges)


def test_encode_levels_round_robin_tiles(emu, ora):
    """fewer CTAs than tiles (encode_grid < 0: absolute CTA count): each CTA stages several tiles in turn"""
    rng = np.random.default_rng(77)
    train = bytes(rng.integers(97, 103, size=3000, dtype=np.uint8))
    merges = [tuple(int(x) for x in r) for r in ora.train(train, 256 + 70)[0]]
    text = bytes(rng.integers(97, 103, size=4000, dtype=np.uint8))
    try:
        for g in (-1, -2, -3):
            emu.set_option("encode_grid", g)
            assert np.array_equal(emu.encode(text, merges), ora.encode(text, merges, linear=False))
    finally:
        emu.set_option("encode_grid", 24)


def test_packed_halfword_primitives(tmp_path):
    """vec_has / vec_mask / vec_candidates / vec_pair_candidates (packed 16-bit minimum, funnel shifts, zero-halfword tricks)
    against their slot-by-slot definitions on 2 x 2,000,000 random vectors (tests/emul/vec_check.cpp)"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "vec_check"
    subprocess.run(["g++", "-O2", "-std=c++17", f"-I{root}/tests/emul", f"-I{root}/zig-bpe_b200/csrc", f"{root}/tests/emul/vec_check.cpp",
                    f"{root}/tests/emul/cuda_emul.cpp", "-o", str(exe)], check=True)
    for seed in ("1", "7"):
        out = subprocess.run([str(exe), seed], capture_output=True, text=True).stdout.split()
        assert out and int(out[0]) == 2000000 and int(out[1]) == 0, out


def test_encode_into_a_caller_owned_buffer(emu, ora, taylor):
    """Engine.encode(out=...) hands the caller's uint16 buffer to bpe_encode (the C ABI's own contract: the caller allocates
    n ids) and returns a view of it"""
    data = taylor[:3000]
    m, _ = emu.train(data, 290)
    want = emu.encode(data, m)
    buf = np.full(len(data) + 5, 0xABCD, dtype=np.uint16)
    got = emu.encode(data, m, out=buf)
    assert np.array_equal(got, want) and got.base is buf and (buf[len(data):] == 0xABCD).all()
    assert np.array_equal(got, ora.encode(data, merges_array(m)))
    with pytest.raises(ValueError):
        emu.encode(data, m, out=np.zeros(10, dtype=np.uint16))
    with pytest.raises(ValueError):
        emu.encode(data, m, out=np.zeros(len(data), dtype=np.int16))
