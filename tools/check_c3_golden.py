"""tools/check_c3_golden.py — one C3 training on cuda:0 compared with tests/golden/c3_merges.txt / c3_counts.txt (as many
merges as the oracle job has produced); prints the first differing merge if any."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
zb = importlib.import_module("zig-bpe_b200")
from tools import synthcorpus as sc
G = os.path.join(ROOT, "tests", "golden")
gm = np.array([[int(x) for x in ln.split(",")] for ln in open(os.path.join(G, "c3_merges.txt")).read().split()], dtype=np.int64).reshape(-1, 3)
gc = np.array([int(x) for x in open(os.path.join(G, "c3_counts.txt")).read().split()], dtype=np.int64)
data = sc.generate(1_000_000_000, sc.SEED_C3, sc.BYTE)
eng = zb.Engine(0)
m, c = eng.train(data, 8192)
ma = np.stack([m["first"], m["second"], m["new_token"]], axis=1).astype(np.int64)
k = len(gm)
bad_m = np.nonzero((ma[:k] != gm).any(axis=1))[0]
bad_c = np.nonzero(np.asarray(c[:k], dtype=np.int64) != gc)[0]
print(json.dumps({"golden_merges": int(k), "gpu_merges": int(len(m)), "merges_equal": bool(len(bad_m) == 0), "counts_equal": bool(len(bad_c) == 0),
                  "first_bad_merge": int(bad_m[0]) if len(bad_m) else None, "first_bad_count": int(bad_c[0]) if len(bad_c) else None,
                  "device_ms": round(eng.last_stats["device_ms"], 1)}))
