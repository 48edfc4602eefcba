"""tests/dist_worker.py — one rank of the world_size-N CPU test of sharded training.

The kernels run in the CPU emulation build (tests/emul); the per-step exchange that NCCL performs on
GPUs is delegated to a callback which all-reduces the buffer over torch.distributed/gloo. Every rank
trains on its contiguous shard and must obtain the merges the oracle learns from the whole corpus."""
import ctypes
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    case, vocab, out_path = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    dist.init_process_group("gloo", rank=rank, world_size=world)
    zb = importlib.import_module("zig-bpe_b200")
    lib = zb.load_library(os.path.join(ROOT, "tests", "emul", "libbpe_emul.so"))

    CB = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int)

    def allreduce(buf, count, kind):
        try:
            if kind == 0:
                a = np.ctypeslib.as_array(ctypes.cast(buf, ctypes.POINTER(ctypes.c_uint32)), shape=(count,))
                t = torch.from_numpy(a.astype(np.int64))  # exact: sums stay far below 2^63
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                a[:] = (t.numpy() & 0xFFFFFFFF).astype(np.uint32)
            elif kind == 3:  # u64 sum (the byte-pair histogram; test inputs stay far below 2^63)
                a = np.ctypeslib.as_array(ctypes.cast(buf, ctypes.POINTER(ctypes.c_uint64)), shape=(count,))
                t = torch.from_numpy(a.view(np.int64).copy())
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                a[:] = t.numpy().view(np.uint64)
            else:
                a = np.ctypeslib.as_array(ctypes.cast(buf, ctypes.POINTER(ctypes.c_uint64)), shape=(count,))
                # order-preserving map of u64 onto i64 so that gloo's signed min/max are exact
                t = torch.from_numpy((a ^ np.uint64(1 << 63)).view(np.int64).copy())
                dist.all_reduce(t, op=dist.ReduceOp.MIN if kind == 1 else dist.ReduceOp.MAX)
                a[:] = t.numpy().view(np.uint64) ^ np.uint64(1 << 63)
            return 0
        except Exception as e:  # pragma: no cover
            print("allreduce callback failed:", e, file=sys.stderr)
            return 1

    cb = CB(allreduce)
    lib.bpe_ctx_create_dist_cb.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, CB]
    ctx = ctypes.c_void_p()
    assert lib.bpe_ctx_create_dist_cb(ctypes.byref(ctx), rank, world, cb) == 0
    eng = zb.Engine.__new__(zb.Engine)
    eng.lib, eng._ctx, eng.rank, eng.world, eng.last_stats = lib, ctx, rank, world, {}
    shm = None
    if os.environ.get("PEER_SHM"):
        # the default data plane of the GPU build — peer mailboxes written by the last CTA of the merge pass, summed by
        # the apply kernel — with the mailboxes in a shared-memory file instead of NVLink peer memory
        lib.bpe_peer_shm_bytes.restype = ctypes.c_size_t
        lib.bpe_peer_shm_bytes.argtypes = [ctypes.c_int]
        lib.bpe_ctx_set_peer_shm.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        nbytes = lib.bpe_peer_shm_bytes(world)
        path = os.environ["PEER_SHM"]
        if rank == 0:
            with open(path, "wb") as f:
                f.truncate(nbytes)
        dist.barrier()
        shm = np.memmap(path, dtype=np.uint8, mode="r+", shape=(nbytes,))
        assert lib.bpe_ctx_set_peer_shm(ctx, shm.ctypes.data, nbytes) == 0
        dist.barrier()
    eng.set_option("table_log2", 13)
    for opt in sys.argv[4:]:
        k, v = opt.split("=")
        eng.set_option(k, int(v))

    data = make_case(case)
    bounds = shard_bounds(case, len(data), world)
    shard = data[bounds[rank]:bounds[rank + 1]]
    if os.environ.get("DIST_MODE") == "encode":
        # sharded encode with a fixed merge list: the concatenation of the ranks' ids must equal the
        # single-process encoding of the whole text
        ids = eng.encode(shard, encode_merges(case))
        np.savez(f"{out_path}.{rank}.npz", ids=ids, path=eng.last_stats["kernel_calls"][11], who=eng.last_stats["kernel_calls"][10])
        dist.barrier()
        dist.destroy_process_group()
        return
    merges, counts = eng.train(shard, vocab)
    got = np.stack([merges["first"], merges["second"], merges["new_token"]], axis=1) if len(merges) else np.zeros((0, 3), np.uint16)
    np.savez(f"{out_path}.{rank}.npz", merges=got, counts=counts, tie_steps=eng.last_stats["tie_steps"],
             tie_slow=eng.last_stats["tie_slow_steps"])
    dist.barrier()
    dist.destroy_process_group()


ENCODE_MERGES = {
    "aaaa": [(97, 97, 256), (256, 256, 257), (257, 97, 258)],
    "abab": [(97, 98, 256), (256, 256, 257), (98, 97, 258), (257, 97, 259)],
    "runs": [(97, 97, 256), (256, 256, 257), (257, 257, 258), (98, 258, 259), (99, 99, 260)],
    "cascade": [(ord("X"), ord("b"), ord("X")), (ord("a"), ord("a"), ord("a"))],
    "tiny": [(97, 98, 256), (256, 99, 257), (257, 257, 258)],
    # several commuting pairs per level (one level pass each), pairs that straddle the shard ends
    "rand4": [(97, 98, 256), (99, 100, 257), (100, 97, 258), (256, 257, 259), (98, 99, 260), (259, 258, 261), (97, 97, 262),
              (256, 99, 263), (100, 100, 264), (263, 264, 265), (99, 97, 266), (98, 98, 267), (266, 267, 268)],
}


def encode_merges(case):
    if case in ("taylor", "taylor_odd", "taylor_even"):  # the reference's committed merges.txt
        return [tuple(int(x) for x in line.split(",")) for line in open(os.path.join(ROOT, "tests", "golden", "merges_300.txt"))]
    return ENCODE_MERGES[case]


def make_case(case):
    rng = np.random.default_rng(1234)
    if case in ("taylor", "taylor_odd", "taylor_even"):
        return open(os.path.join(ROOT, "tests", "golden", "taylorswift.txt"), "rb").read()[:30000]
    if case == "aaaa":
        return b"a" * 3001
    if case == "abab":
        return b"ab" * 1500 + b"a"
    if case == "runs":
        return b"xyz" + b"a" * 700 + b"b" + b"a" * 513 + b"cc" + b"a" * 1024 + b"q"
    if case == "rand4":
        return bytes(rng.integers(97, 101, size=6000, dtype=np.uint8))
    if case == "rand256":
        return bytes(rng.integers(0, 256, size=6000, dtype=np.uint8))
    if case == "cascade":
        return b"Xbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbb c" + b"a" * 777 + b"Xbb"
    if case == "tiny":
        return b"abcabcabc"
    raise ValueError(case)


def shard_bounds(case, n, world):
    if case == "taylor_odd":  # shards of very different sizes, one shorter than the halo, cuts off the 64-byte core grid
        b = [0, 301, 301 + 77, 20011][: world] + [n]
        return b
    if case == "taylor_even":  # uneven shards, all above the tile encoder's minimum, cut at odd offsets
        return [0, 1031, 1031 + 7777, 20011][: world] + [n]
    if case == "tiny":  # some ranks get one byte or nothing
        b = [0, 1] + [n] * (world - 1)
        return b[: world + 1] if world >= 2 else [0, n]
    if case == "runs":  # cut inside the runs of 'a', at odd offsets
        cuts = sorted({0, n} | {min(n, 3 + 351 + 700 * i) for i in range(world - 1)})
        while len(cuts) < world + 1:
            cuts.insert(-1, cuts[-2])
        return cuts
    return [n * r // world for r in range(world + 1)]


if __name__ == "__main__":
    main()
