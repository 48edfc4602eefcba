"""tools/fuzz_emul_encode.py [seconds] [seed] — TEST INFRASTRUCTURE: trained merge lists on random / structured texts through the
CPU emulation of the level encoders — one pass per level and fused level groups (encode_fuse 2..8), few CTAs (a CTA walks
many tiles) or many, compaction between passes — against the oracle's encode. Prints failures and a summary."""
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
    alpha = int(rng.choice([2, 3, 4, 8, 26, 256]))
    ntrain = int(rng.integers(200, 6000))
    if rng.random() < 0.5:
        train = bytes(rng.integers(0, alpha, size=ntrain, dtype=np.uint8))
    else:
        words = [bytes(rng.integers(97, 97 + min(alpha, 26), size=int(rng.integers(1, 8)), dtype=np.uint8)) for _ in range(20)]
        train = b" ".join(words[int(i)] for i in rng.integers(0, 20, size=ntrain // 4 + 1))
    m, _ = ora.train(train, 256 + int(rng.integers(2, 200)), fast=True)
    merges = [tuple(int(x) for x in r) for r in m]
    if not merges:
        continue
    n = int(rng.integers(1, 9000))
    text = (train * (n // max(len(train), 1) + 1))[:n] if rng.random() < 0.5 else bytes(rng.integers(0, alpha, size=n, dtype=np.uint8))
    want = ora.encode(text, merges, linear=True)
    opts = {"encode_impl": 2, "encode_fuse": int(rng.choice([0, 2, 3, 5, 8])), "encode_grid": int(rng.choice([24, -1, -2, -3, 0])),
            "encode_compact_pct": int(rng.choice([0, 85, 100])), "encode_filter": int(rng.choice([0, 1, 2]))}
    for k, v in opts.items(): emu.set_option(k, v)
    try:
        got = emu.encode(text, merges)
        ok = np.array_equal(got, want)
    except Exception as e:  # noqa: BLE001
        ok = False; print("EXC", repr(e)[:200])
    runs += 1
    if not ok:
        bad += 1
        print("FAIL", alpha, len(merges), n, opts, flush=True)
print(f"{runs} runs, {bad} failures in {time.time() - t0:.0f} s")
