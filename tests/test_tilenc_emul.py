"""Tile-resident encoder (zig-bpe_b200/csrc/bpe_tilenc.cuh) in the CPU emulation of the kernels: same ids as the oracle's
verbatim encode loop (basic_tokenizer.zig:71-88) for every tile size, with the seams between windows exercised, the work
queue overflowing, runs of equal tokens, and the fall-back when a window gives up."""
import numpy as np
import pytest

TILE_PATH, SEG_PATH = 2, 1  # bpe_stats_t.kernel_calls[10]: the encoder that produced the ids (0: passes)
USED, FELL_BACK = 1, 2      # bpe_stats_t.kernel_calls[11]


def _enc(eng, data, merges, tile=512, impl=0):
    try:
        eng.set_option("encode_tile", tile)
        eng.set_option("encode_impl", impl)
        eng.set_option("encode_try_tiles", 2)  # impl 0 picks by cost (short lists on one GPU go to the level passes); 2: tiles first
        ids = eng.encode(data, merges)
        return ids, eng.last_stats["kernel_calls"][10], eng.last_stats["kernel_calls"][11]
    finally:
        eng.set_option("encode_tile", 8192)
        eng.set_option("encode_impl", 0)
        eng.set_option("encode_try_tiles", 1)


@pytest.mark.parametrize("tile", [512, 1024, 2048, 8192])
def test_taylor_golden_list_every_tile_size(emu, ora, taylor, golden_merges, tile):
    data = taylor[:40000]
    ids, who, verdict = _enc(emu, data, golden_merges, tile)
    assert (who, verdict) == (TILE_PATH, USED)
    assert np.array_equal(ids, ora.encode(data, golden_merges, linear=True))


def test_trained_list_on_other_text(emu, ora, synth):
    """1,000 merges trained on one synthetic corpus, applied to another (pairs the list never saw, all 256 byte values);
    odd sizes around the tile size: one tile, two tiles that share what is left, a clipped last window"""
    train = bytes(synth.generate(300_000, synth.SEED_C3, synth.BYTE))
    om, _ = ora.train(train, 256 + 1000, fast=True)
    other = bytes(synth.generate(60_000, synth.SEED_C5, synth.BYTE)) + train[:20_000]
    want = ora.encode(other, om, linear=True)
    for tile in (512, 2048):
        ids, who, verdict = _enc(emu, other, om, tile)
        assert (who, verdict) == (TILE_PATH, USED)
        assert np.array_equal(ids, want)
    for n in (1, 2, 63, 64, 65, 511, 512, 513, 1023, 1024, 1025, 1037, 1536, 3 * 512 + 1, 5 * 512 - 1):
        ids, who, verdict = _enc(emu, other[:n], om, 512)
        assert (who, verdict) == (TILE_PATH, USED), n
        assert np.array_equal(ids, ora.encode(other[:n], om, linear=True)), n


def test_utf8_text_and_round_trip(emu, ora, synth):
    data = bytes(synth.generate(60_000, synth.SEED_C2, synth.UTF8))
    om, _ = ora.train(data[:40_000], 256 + 600, fast=True)
    ids, who, verdict = _enc(emu, data, om, 1024)
    assert (who, verdict) == (TILE_PATH, USED) and np.array_equal(ids, ora.encode(data, om, linear=True))
    assert emu.decode(ids, om) == data


def test_runs_of_equal_tokens(emu, ora):
    """first == second merges are greedy from the start of the run (aaa -> Xa): runs inside tiles, across tile seams,
    and runs longer than 255 tokens (the window gives up and another encoder produces the ids)"""
    merges = [(97, 97, 256), (256, 256, 257), (98, 257, 258), (257, 257, 259)]
    rng = np.random.default_rng(5)
    parts = []
    for _ in range(3000):
        parts.append(b"a" * int(rng.integers(1, 12)) + bytes(rng.integers(98, 102, size=int(rng.integers(1, 6)), dtype=np.uint8)))
    data = b"".join(parts)
    ids, who, verdict = _enc(emu, data, merges)
    assert (who, verdict) == (TILE_PATH, USED) and np.array_equal(ids, ora.encode(data, merges, linear=False))
    data = (b"xy" + b"a" * 70 + b"b" + b"a" * 33 + b"cc") * 40
    ids, who, verdict = _enc(emu, data, merges)
    assert who == TILE_PATH and np.array_equal(ids, ora.encode(data, merges, linear=False))
    data = b"q" + b"a" * 5001 + b"b"
    ids, who, verdict = _enc(emu, data, merges)
    assert who != TILE_PATH and verdict == FELL_BACK and np.array_equal(ids, ora.encode(data, merges, linear=False))
    with pytest.raises(Exception):
        _enc(emu, data, merges, impl=4)  # encode_impl = 4 insists on the tile path


def test_queue_overflow_and_dense_levels(emu, ora):
    """one level whose occurrences exceed the work queue (2,048 entries) several times over in a single tile, then pairs
    created inside a level range"""
    merges = [(97, 98, 256), (99, 100, 257), (256, 257, 258), (258, 101, 259), (101, 258, 260)]
    data = b"abcd" * 2000 + b"e" + b"abcdeabcde" * 300
    ids, who, verdict = _enc(emu, data, merges, 8192)
    assert (who, verdict) == (TILE_PATH, USED) and np.array_equal(ids, ora.encode(data, merges, linear=False))
    merges = [(97, 97, 256)]
    rng = np.random.default_rng(9)
    data = b"".join(b"a" * int(rng.integers(2, 9)) + b"b" for _ in range(1500))
    ids, who, verdict = _enc(emu, data, merges, 8192)
    assert (who, verdict) == (TILE_PATH, USED) and np.array_equal(ids, ora.encode(data, merges, linear=False))


def test_periodic_text(emu, ora):
    # 16-byte tokens with period 2: every window sees the same phase, the seams exist
    merges = [(97, 98, 256), (256, 256, 257), (257, 257, 258), (258, 258, 259)]
    data = b"b" + b"ab" * 4000
    ids, who, verdict = _enc(emu, data, merges)
    assert np.array_equal(ids, ora.encode(data, merges, linear=False))
    # 24-byte tokens with period 3: the phase of the (X,X) pairs depends on where a window starts
    merges = [(97, 98, 256), (256, 99, 257), (257, 257, 258), (258, 258, 259), (259, 259, 260)]
    data = b"c" + b"abc" * 3000
    ids, who, verdict = _enc(emu, data, merges)
    assert np.array_equal(ids, ora.encode(data, merges, linear=False))


def test_irregular_lists_take_the_pass_path(emu, ora):
    data = b"abcabcabcXbbbb" * 40
    for merges in ([(97, 98, 99)], [(97, 98, 256), (97, 98, 257)], [(97, 98, 256), (99, 99, 256)], [(ord("X"), 98, ord("X"))],
                   [(97, 98, 65535), (65535, 99, 300)]):
        ids, who, verdict = _enc(emu, data, merges)
        assert who == 0 and verdict == 0 and np.array_equal(ids, ora.encode(data, merges, linear=False))


def test_fuzzed_regular_lists(emu, ora):
    """random regular lists over a 4-letter alphabet (dense in overlaps and runs) on random text"""
    rng = np.random.default_rng(177)
    used = 0
    for it in range(150):
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
        data = bytes(rng.integers(97, 101, size=int(rng.integers(1, 4000)), dtype=np.uint8))
        ids, who, verdict = _enc(emu, data, merges, int(rng.choice([512, 1024])))
        used += who == TILE_PATH
        assert np.array_equal(ids, ora.encode(data, merges, linear=False)), (it, merges)
    assert used > 100


def test_matches_the_other_encoders(emu, ora, taylor):
    om, _ = ora.train(taylor[:50000], 256 + 200, fast=True)
    a, who, _ = _enc(emu, taylor[:60000], om, 1024, impl=0)
    b, _, _ = _enc(emu, taylor[:60000], om, impl=2)
    assert who == TILE_PATH and np.array_equal(a, b)


def test_streaming_from_host_buffers(emu, ora, synth, taylor, golden_merges):
    """bpe_encode on an input of several chunks: the chunks stream through the device (text in, ids out) and are encoded like
    the shards of a multi-GPU run — halos from the neighbouring chunks, bridge windows at both ends — so their ids
    concatenate to the encoding of the whole text; sizes that leave a short last chunk, runs across chunk boundaries"""
    train = bytes(synth.generate(200_000, synth.SEED_C3, synth.BYTE))
    om, _ = ora.train(train, 256 + 500, fast=True)
    try:
        emu.set_option("stream_chunk_bytes", 8192)
        emu.set_option("encode_tile", 1024)
        for data, merges in ((train[:70_001], om), (taylor[:33_333], golden_merges), (taylor[:16_384], golden_merges)):
            ids = emu.encode(data, merges)
            st = emu.last_stats
            assert st["kernel_calls"][10] == 2 and st["kernel_calls"][11] == 1 and st["kernel_calls"][9] == -(-len(data) // 8192)
            assert np.array_equal(ids, ora.encode(data, merges, linear=True))
        merges = [(97, 97, 256), (256, 256, 257), (98, 257, 258), (257, 257, 259)]
        rng = np.random.default_rng(15)
        data = b"".join(b"a" * int(rng.integers(1, 12)) + bytes(rng.integers(98, 102, size=int(rng.integers(1, 6)), dtype=np.uint8)) for _ in range(6000))
        ids = emu.encode(data, merges)
        assert emu.last_stats["kernel_calls"][10] == 2 and np.array_equal(ids, ora.encode(data, merges, linear=False))
        # a window that gives up in some chunk: the whole call falls back to the resident encoders
        data = data[:20000] + b"q" + b"a" * 5001 + b"b" + data[20000:40000]
        ids = emu.encode(data, merges)
        assert emu.last_stats["kernel_calls"][10] != 2 and np.array_equal(ids, ora.encode(data, merges, linear=False))
        # small inputs (fewer than two chunks) do not stream
        ids = emu.encode(taylor[:9000], golden_merges)
        assert emu.last_stats["kernel_calls"][9] == 0 and np.array_equal(ids, ora.encode(taylor[:9000], golden_merges, linear=True))
    finally:
        emu.set_option("stream_chunk_bytes", 0)
        emu.set_option("encode_tile", 8192)


def test_default_dispatch_by_cost(emu, ora, taylor, golden_merges):
    """encode_impl = 0 on one GPU: a short list (few levels) takes the level passes, a list whose schedule is longer than
    encode_seg_min_steps the tile-resident kernel"""
    ids = emu.encode(taylor[:20000], golden_merges)
    assert emu.last_stats["kernel_calls"][10] == 0 and np.array_equal(ids, ora.encode(taylor[:20000], golden_merges, linear=True))
    try:
        emu.set_option("encode_seg_min_steps", 5)
        ids = emu.encode(taylor[:20000], golden_merges)  # 44 merges -> 6 levels + 1 run merge
        assert emu.last_stats["kernel_calls"][10] == TILE_PATH and np.array_equal(ids, ora.encode(taylor[:20000], golden_merges, linear=True))
    finally:
        emu.set_option("encode_seg_min_steps", 450)
