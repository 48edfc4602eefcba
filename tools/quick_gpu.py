"""tools/quick_gpu.py — ad-hoc timing on a GPU box: python tools/quick_gpu.py <bytes> <vocab> [variant] [opt=val ...]"""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

zb = importlib.import_module("zig-bpe_b200")
from tools import synthcorpus as sc

n = int(float(sys.argv[1])); vocab = int(sys.argv[2]); variant = int(sys.argv[3]) if len(sys.argv) > 3 else 2
opts = dict(a.split("=") for a in sys.argv[4:])
t = time.time(); data = sc.generate(n, sc.SEED_C3 if variant == 2 else sc.SEED_C2, variant); print("gen", round(time.time() - t, 2), "s", flush=True)
eng = zb.Engine(0)
for k, v in opts.items():
    eng.set_option(k, int(v))
for rep in range(2):
    t = time.time(); m, c = eng.train(data, vocab); dt = time.time() - t
    st = eng.last_stats
    print(json.dumps({"rep": rep, "n": n, "merges": len(m), "wall_s": round(dt, 3), "merges_per_s": round(len(m) / dt, 1),
                      "scan_GBps": round(st["scanned_slots"] * 2 / 1e9 / (st["device_ms"] / 1e3), 1), **{k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items() if not k.startswith("kernel_") and not k.endswith("_calls") and not k.endswith("pairs_ms") and k != "replace_pair_ms"}, "kernel_ms": {n: round(t, 1) for n, t in zip(zb.KERNEL_CLASSES, st["kernel_ms"]) if t}}), flush=True)
print("last merges", m[-3:], c[-3:])
