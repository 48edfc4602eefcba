"""Multi-GPU host logic on the CPU: world_size-2 (and 3) runs of the sharded training path over gloo.
Kernels run in the emulation build; see tests/dist_worker.py. All ranks must return the merges the
oracle learns from the unsharded corpus (so sharding is invisible in the result)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests"))
import dist_worker  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(case, vocab, world, tmp_path, *opts, mode="train", peer_shm=False):
    out = str(tmp_path / f"{case}_{world}_{mode}")
    port = _free_port()
    procs = []
    shm_path = f"/dev/shm/bpe_b200_test_{os.getpid()}_{port}" if peer_shm else ""
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1", DIST_MODE=mode, PEER_SHM=shm_path)
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dist_worker.py"), case, str(vocab), out, *opts],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = [p.communicate(timeout=600)[0] for p in procs]
    if shm_path and os.path.exists(shm_path):
        os.unlink(shm_path)
    for r, p in enumerate(procs):
        assert p.returncode == 0, f"rank {r} failed:\n{logs[r][-3000:]}"
    return [np.load(f"{out}.{r}.npz") for r in range(world)]


@pytest.mark.parametrize("case,vocab,world", [
    ("taylor", 290, 2), ("aaaa", 266, 2), ("abab", 268, 2), ("runs", 280, 2), ("runs", 280, 3),
    ("rand4", 290, 2), ("rand256", 300, 3), ("tiny", 270, 2),
])
def test_sharded_train_matches_oracle(ora, tmp_path, case, vocab, world):
    res = _run(case, vocab, world, tmp_path)
    om, oc = ora.train(dist_worker.make_case(case), vocab, fast=True)
    for r in range(world):
        assert np.array_equal(res[r]["merges"], om), f"rank {r}"
        assert np.array_equal(res[r]["counts"], oc), f"rank {r}"


def test_sharded_train_forced_replay(ora, tmp_path):
    """Every tie through the cross-rank replay (first-occurrence positions min-reduced in key order)."""
    res = _run("rand4", 285, 2, tmp_path, "force_slow_tiebreak=1")
    om, oc = ora.train(dist_worker.make_case("rand4"), 285, fast=True)
    for r in range(2):
        assert np.array_equal(res[r]["merges"], om) and np.array_equal(res[r]["counts"], oc)
        assert res[r]["tie_slow"] == res[r]["tie_steps"] > 0
    res = _run("taylor", 280, 2, tmp_path, "check_tiebreak=1")
    om, _ = ora.train(dist_worker.make_case("taylor"), 280, fast=True)
    assert np.array_equal(res[0]["merges"], om)


@pytest.mark.parametrize("case,world", [("aaaa", 2), ("abab", 3), ("runs", 2), ("runs", 3), ("cascade", 2), ("tiny", 2), ("rand4", 2),
                                        ("rand4", 3)])
def test_sharded_encode_matches_single_process(ora, tmp_path, case, world):
    """Sharded encode: tokens that straddle shard boundaries (also runs cut at odd offsets, shards of one byte,
    the cascade rule) must come out exactly as in the reference's single-sequence encode."""
    res = _run(case, 300, world, tmp_path, mode="encode")
    got = np.concatenate([res[r]["ids"] for r in range(world)])
    want = ora.encode(dist_worker.make_case(case), dist_worker.ENCODE_MERGES[case], linear=False)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("case,world", [("taylor", 2), ("taylor", 3), ("taylor_odd", 3), ("taylor_odd", 4), ("rand4", 3)])
def test_sharded_segment_encode(ora, tmp_path, case, world):
    """The segment-resident encoder across shards: every rank sees 512 bytes of its neighbours' text, windows lie on
    a global grid, and the concatenation of the ranks' ids is the single-process encoding (tokens that straddle a
    shard boundary included). All ranks must have taken the segment path."""
    res = _run(case, 300, world, tmp_path, "encode_seg_min_steps=0", "encode_try_tiles=0", mode="encode")
    got = np.concatenate([res[r]["ids"] for r in range(world)])
    want = ora.encode(dist_worker.make_case(case), dist_worker.encode_merges(case), linear=True)
    assert np.array_equal(got, want)
    assert all(int(res[r]["path"]) == 1 for r in range(world))
    # each rank holds exactly the tokens that start inside its shard
    data = dist_worker.make_case(case)
    bounds = dist_worker.shard_bounds(case, len(data), world)
    lens = {b: 1 for b in range(256)}
    for a, b, z in dist_worker.encode_merges(case):
        lens[z] = lens[a] + lens[b]
    pos = 0
    for r in range(world):
        for t in res[r]["ids"]:
            assert bounds[r] <= pos < bounds[r + 1]
            pos += lens[int(t)]
    assert pos == len(data)


@pytest.mark.parametrize("case,world,tile", [("taylor", 2, 1024), ("taylor", 3, 2048), ("taylor_even", 4, 1024), ("rand4", 3, 1024), ("rand4", 2, 8192)])
def test_sharded_tile_encode(ora, tmp_path, case, world, tile):
    """The tile-resident encoder across shards: around every shard boundary both ranks encode the same bridge window and
    stitch it between their edge tiles, so the concatenation of the ranks' ids is the single-process encoding and every
    rank holds exactly the tokens that start inside its shard — with nothing exchanged but 512 bytes of text."""
    res = _run(case, 300, world, tmp_path, f"encode_tile={tile}", mode="encode")
    data = dist_worker.make_case(case)
    got = np.concatenate([res[r]["ids"] for r in range(world)])
    want = ora.encode(data, dist_worker.encode_merges(case), linear=True)
    assert np.array_equal(got, want)
    assert all(int(res[r]["who"]) == 2 and int(res[r]["path"]) == 1 for r in range(world))
    bounds = dist_worker.shard_bounds(case, len(data), world)
    lens = {b: 1 for b in range(256)}
    for a, b, z in dist_worker.encode_merges(case):
        lens[z] = lens[a] + lens[b]
    pos = 0
    for r in range(world):
        for t in res[r]["ids"]:
            assert bounds[r] <= pos < bounds[r + 1]
            pos += lens[int(t)]
    assert pos == len(data)


@pytest.mark.parametrize("case,vocab,world", [("taylor", 290, 2), ("runs", 280, 3), ("rand4", 290, 2), ("rand256", 300, 3), ("abab", 268, 2),
                                              ("tiny", 270, 2)])
def test_sharded_train_through_peer_mailboxes(ora, tmp_path, case, vocab, world):
    """The default multi-GPU data plane: the last CTA of every rank's merge pass writes its deltas and shard ends into
    every peer's mailbox and raises a flag, the apply kernel waits for the flags and sums the slots (and its halo CTAs
    read the neighbours' shard ends straight from the mailbox). Here the ranks are processes and the mailboxes a
    shared-memory file; on GPUs it is NVLink peer memory. Same merges as the oracle on the unsharded corpus."""
    res = _run(case, vocab, world, tmp_path, peer_shm=True)
    om, oc = ora.train(dist_worker.make_case(case), vocab, fast=True)
    for r in range(world):
        assert np.array_equal(res[r]["merges"], om), f"rank {r}"
        assert np.array_equal(res[r]["counts"], oc), f"rank {r}"


def test_sharded_streaming_encode(ora, tmp_path):
    """several ranks, each streaming its shard through the device in chunks: shard boundaries and chunk boundaries are
    handled by the same bridge windows"""
    res = _run("taylor", 300, 2, tmp_path, "encode_tile=1024", "stream_chunk_bytes=4096", mode="encode")
    data = dist_worker.make_case("taylor")
    got = np.concatenate([res[r]["ids"] for r in range(2)])
    assert np.array_equal(got, ora.encode(data, dist_worker.encode_merges("taylor"), linear=True))
    assert all(int(res[r]["who"]) == 2 and int(res[r]["path"]) == 1 for r in range(2))
