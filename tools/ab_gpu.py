"""tools/ab_gpu.py <bytes> <vocab> <set>[,<set>...] — one corpus, one context, one training per option set (after a warm-up
run); a set is opt=val[:opt=val...] or "default". Prints device time, merge-kernel time (profile 3 samples) and whether every
set learned the same merges."""
import hashlib
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

zb = importlib.import_module("zig-bpe_b200")
from tools import synthcorpus as sc

n = int(float(sys.argv[1])); vocab = int(sys.argv[2]); sets = sys.argv[3].split(",")
reps = int(os.environ.get("REPS", 1))
d_text = torch.from_numpy(sc.generate(n, sc.SEED_C3, sc.BYTE)).cuda()
eng = zb.Engine(0)
eng.set_option("profile", 3)
eng.train(None, vocab, device_ptr=d_text.data_ptr(), n=n)  # warm-up (buffers, clocks)
ref = None
touched = set()
for s in sets:
    opts = {} if s == "default" else {k: int(v) for k, v in (p.split("=") for p in s.split(":"))}
    for k, v in opts.items():
        eng.set_option(k, v); touched.add(k)
    best = None
    for _ in range(reps):
        m, c = eng.train(None, vocab, device_ptr=d_text.data_ptr(), n=n)
        st = eng.last_stats
        if best is None or st["device_ms"] < best["device_ms"]:
            best = dict(st)
    sha = hashlib.sha256(np.ascontiguousarray(m).tobytes() + np.ascontiguousarray(c).tobytes()).hexdigest()[:16]
    ref = ref or sha
    km = {name: round(t, 2) for name, t in zip(zb.KERNEL_CLASSES, best["kernel_ms"]) if t}
    print(json.dumps({"set": s, "merges": len(m), "device_ms": round(best["device_ms"], 1), "merges_per_s": round(len(m) / best["device_ms"] * 1e3, 1),
                      "scan_GBps": round(best["scanned_slots"] * 2 / 1e9 / (best["device_ms"] / 1e3), 1), "sampled_kernel_ms": km,
                      "same_as_first": sha == ref}), flush=True)
