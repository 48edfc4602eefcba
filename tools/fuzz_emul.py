"""tools/fuzz_emul.py [seconds] [seed] — TEST INFRASTRUCTURE: random small corpora through the CPU emulation of the train kernels
with random merge-pass options (looped / one CTA per tile, queue-less / queued, pair filter, second staging buffer, compaction
threshold), every step recounted, merges and counts compared with the oracle. Prints one line per failure and a summary."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
zb = importlib.import_module("zig-bpe_b200")
from oracle import oracle_py as ora

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
emu = zb.Engine(lib_path=os.path.join(ROOT, "tests", "emul", "libbpe_emul.so"))
emu.set_option("table_log2", 13)
t0 = time.time(); runs = bad = 0
while time.time() - t0 < budget:
    kind = int(rng.integers(0, 5))
    n = int(rng.integers(2, 6000))
    if kind == 0: data = rng.integers(97, 97 + int(rng.integers(1, 5)), size=n, dtype=np.uint8)
    elif kind == 1: data = rng.integers(0, 256, size=n, dtype=np.uint8)
    elif kind == 2: data = np.frombuffer((b"ab" * (n // 2) + b"a" * int(rng.integers(0, 600)) + b"ba" * int(rng.integers(0, 300)))[:max(n, 2)], dtype=np.uint8).copy()
    elif kind == 3:
        words = [bytes(rng.integers(97, 105, size=int(rng.integers(1, 7)), dtype=np.uint8)) for _ in range(12)]
        data = np.frombuffer(b" ".join(words[int(i)] for i in rng.integers(0, 12, size=n // 3 + 1))[:max(n, 2)], dtype=np.uint8).copy()
    else: data = np.frombuffer((b"a" * int(rng.integers(1, 3000)) + b"b" + b"a" * int(rng.integers(1, 3000)))[:max(n, 2)], dtype=np.uint8).copy()
    vocab = 256 + int(rng.integers(1, 60))
    opts = {"verify_recount": 1, "check_tiebreak": int(rng.integers(0, 2)), "merge_loop": int(rng.choice([0, 1, 2, 3, 6, 50])),
            "merge_direct": int(rng.choice([0, 1, 3, 8, 1 << 20])), "merge_pairfilter": int(rng.integers(0, 2)), "merge_dbuf": int(rng.integers(0, 2)),
            "merge_prestage": int(rng.integers(0, 2)), "compact_pct": int(rng.choice([0, 85, 93, 100])), "fuse_halo": int(rng.integers(0, 2))}
    for k, v in opts.items(): emu.set_option(k, v)
    try:
        m, c = emu.train(bytes(data), vocab)
        om, oc = ora.train(bytes(data), vocab, fast=True)
        got = np.stack([m["first"], m["second"], m["new_token"]], axis=1)
        ok = np.array_equal(got, om) and np.array_equal(c, oc)
    except Exception as e:  # noqa: BLE001
        ok = False; print("EXC", repr(e)[:200])
    runs += 1
    if not ok:
        bad += 1
        print("FAIL", kind, n, vocab, opts, data[:40].tobytes(), flush=True)
print(f"{runs} runs, {bad} failures in {time.time() - t0:.0f} s")
