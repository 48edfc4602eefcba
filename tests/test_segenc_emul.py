"""Segment-resident encoder (zig-bpe_b200/csrc/bpe_segenc.cuh) in the CPU emulation of the kernels: same ids as the
oracle's verbatim encode loop (basic_tokenizer.zig:71-88), on every window geometry, with the seams between windows,
CTAs and shards exercised, and the fall-back to the level passes when two windows share no token."""
import numpy as np
import pytest

from conftest import merges_array

SEG_USED, SEG_FELL_BACK = 1, 2  # bpe_stats_t.kernel_calls[11]


def _path(eng):
    return eng.last_stats["kernel_calls"][11]


def _enc(eng, data, merges, geom=0, impl=0):
    try:
        eng.set_option("encode_geom", geom)
        eng.set_option("encode_impl", impl)
        eng.set_option("encode_seg_min_steps", 0)  # impl 0 picks by cost; 0 steps: always try the segment kernel first
        eng.set_option("encode_try_tiles", 0)      # (the tile-resident kernel has its own tests: test_tilenc_emul.py)
        return eng.encode(data, merges), _path(eng)
    finally:
        eng.set_option("encode_geom", 4)
        eng.set_option("encode_impl", 0)
        eng.set_option("encode_seg_min_steps", 450)
        eng.set_option("encode_try_tiles", 1)


@pytest.mark.parametrize("geom", [0, 1, 2, 3, 4, 5])
def test_taylor_golden_list_every_geometry(emu, ora, taylor, golden_merges, geom):
    data = taylor[:60000]
    ids, path = _enc(emu, data, golden_merges, geom)
    assert np.array_equal(ids, ora.encode(data, golden_merges, linear=True))
    assert path == (SEG_USED if geom != 3 else path)  # 12-byte margins (geometry 3) may or may not be enough
    assert path in (SEG_USED, SEG_FELL_BACK)


def test_trained_list_on_other_text(emu, ora, synth):
    """1,000 merges trained on one synthetic corpus, applied to another (pairs the list never saw, all 256 byte values)"""
    train = bytes(synth.generate(300_000, synth.SEED_C3, synth.BYTE))
    om, _ = ora.train(train, 256 + 1000, fast=True)
    other = bytes(synth.generate(150_000, synth.SEED_C5, synth.BYTE)) + train[:20_000]
    want = ora.encode(other, om, linear=True)
    for geom in (0, 1):
        ids, path = _enc(emu, other, om, geom)
        assert path == SEG_USED
        assert np.array_equal(ids, want)
    # odd sizes around the core / tile sizes: the clipped last window, an empty last tile
    for n in (1, 2, 63, 64, 65, 127, 128, 129, 64 * 126 - 1, 64 * 126, 64 * 126 + 1, 64 * 126 + 97, 2 * 64 * 126 + 31):
        ids, path = _enc(emu, other[:n], om)
        assert path == SEG_USED, n
        assert np.array_equal(ids, ora.encode(other[:n], om, linear=True)), n


def test_utf8_text(emu, ora, synth):
    data = bytes(synth.generate(120_000, synth.SEED_C2, synth.UTF8))
    om, _ = ora.train(data[:80_000], 256 + 600, fast=True)
    ids, path = _enc(emu, data, om)
    assert path == SEG_USED and np.array_equal(ids, ora.encode(data, om, linear=True))
    assert emu.decode(ids, om) == data


def test_runs_of_equal_tokens(emu, ora):
    """first == second merges are greedy from the start of the run (aaa -> Xa): runs that cross window cores,
    CTA tiles, and runs too long for any window (then the seams have no common token and the level passes run)"""
    merges = [(97, 97, 256), (256, 256, 257), (98, 257, 258), (257, 257, 259)]
    rng = np.random.default_rng(5)
    parts = []
    for _ in range(4000):
        parts.append(b"a" * int(rng.integers(1, 12)) + bytes(rng.integers(98, 102, size=int(rng.integers(1, 6)), dtype=np.uint8)))
    data = b"".join(parts)
    ids, path = _enc(emu, data, merges)
    assert path == SEG_USED and np.array_equal(ids, ora.encode(data, merges, linear=False))
    # a run of 70 bytes still fits the 128-byte windows; 5,000 bytes of 'a' do not
    data = b"xy" + b"a" * 70 + b"b" + b"a" * 33 + b"cc"
    ids, path = _enc(emu, data * 50, merges)
    assert np.array_equal(ids, ora.encode(data * 50, merges, linear=False))
    data = b"q" + b"a" * 5001 + b"b"
    ids, path = _enc(emu, data, merges)
    assert path == SEG_FELL_BACK and np.array_equal(ids, ora.encode(data, merges, linear=False))


def test_periodic_text_falls_back(emu, ora):
    # period 2, 16-byte tokens: the window grid (multiples of 64) has the same phase everywhere, seams are found
    merges = [(97, 98, 256), (256, 256, 257), (257, 257, 258), (258, 258, 259)]
    data = b"b" + b"ab" * 4000
    ids, path = _enc(emu, data, merges)
    assert path == SEG_USED and np.array_equal(ids, ora.encode(data, merges, linear=False))
    # period 3, 24-byte tokens: the phase of the (X,X) pairs depends on where a window starts, windows share no token
    merges = [(97, 98, 256), (256, 99, 257), (257, 257, 258), (258, 258, 259), (259, 259, 260)]
    data = b"c" + b"abc" * 3000
    ids, path = _enc(emu, data, merges)
    assert path == SEG_FELL_BACK
    assert np.array_equal(ids, ora.encode(data, merges, linear=False))
    with pytest.raises(Exception):
        _enc(emu, data, merges, impl=3)  # encode_impl = 3 insists on the segment path


def test_irregular_lists_take_the_pass_path(emu, ora):
    data = b"abcabcabcXbbbb" * 40
    for merges in ([(97, 98, 99)], [(97, 98, 256), (97, 98, 257)], [(97, 98, 256), (99, 99, 256)], [(ord("X"), 98, ord("X"))]):
        ids, path = _enc(emu, data, merges)
        assert path == 0 and np.array_equal(ids, ora.encode(data, merges, linear=False))
    # id 65535 is the hole marker of the u16 slot array
    merges = [(97, 98, 65535), (65535, 99, 300)]
    ids, path = _enc(emu, data, merges)
    assert path == 0 and np.array_equal(ids, ora.encode(data, merges, linear=False))


def test_fuzzed_regular_lists(emu, ora):
    """random regular lists over a 4-letter alphabet (dense in overlaps and runs) on random text, small windows too"""
    rng = np.random.default_rng(77)
    used = fell = 0
    for it in range(120):
        ids_avail = [97, 98, 99, 100]
        merges, seen = [], set()
        nxt = 256
        for _ in range(int(rng.integers(1, 40))):
            a, b = (int(rng.choice(ids_avail)) for _ in range(2))
            if (a, b) in seen:
                continue
            seen.add((a, b))
            z = nxt if rng.random() < 0.8 else nxt + int(rng.integers(1, 500))  # ids need not be consecutive
            nxt = z + 1
            merges.append((a, b, z))
            ids_avail.append(z)
        data = bytes(rng.integers(97, 101, size=int(rng.integers(1, 3000)), dtype=np.uint8))
        geom = int(rng.choice([0, 3]))
        ids, path = _enc(emu, data, merges, geom)
        assert path in (SEG_USED, SEG_FELL_BACK)
        used += path == SEG_USED
        fell += path == SEG_FELL_BACK
        assert np.array_equal(ids, ora.encode(data, merges, linear=False)), (it, merges, geom)
    assert used > 30 and used + fell == 120


def test_default_dispatch_by_cost(emu, ora, taylor, golden_merges):
    """encode_impl = 0: short lists (few levels) take the level passes, lists with more than encode_seg_min_steps
    schedule steps the segment kernel"""
    emu.set_option("encode_try_tiles", 0)
    ids = emu.encode(taylor[:20000], golden_merges)
    assert _path(emu) == 0 and np.array_equal(ids, ora.encode(taylor[:20000], golden_merges, linear=True))
    try:
        emu.set_option("encode_seg_min_steps", 5)
        ids = emu.encode(taylor[:20000], golden_merges)  # 44 merges -> 6 levels + 1 run merge
        assert _path(emu) == SEG_USED and np.array_equal(ids, ora.encode(taylor[:20000], golden_merges, linear=True))
    finally:
        emu.set_option("encode_seg_min_steps", 450)
        emu.set_option("encode_try_tiles", 1)


def test_matches_level_and_per_merge_paths(emu, ora, taylor):
    om, _ = ora.train(taylor[:50000], 256 + 200, fast=True)
    a, pa = _enc(emu, taylor[:90000], om, impl=0)
    b, pb = _enc(emu, taylor[:90000], om, impl=2)
    c, pc = _enc(emu, taylor[:30000], om, impl=1)
    assert pa == SEG_USED and pb == 0 and pc == 0
    assert np.array_equal(a, b) and np.array_equal(a[: len(c) - 50], c[: len(c) - 50])
