"""tools/steps_gpu.py — cumulative train time vs. number of merge steps (where does the time go?)"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
zb = importlib.import_module("zig-bpe_b200")
from tools import synthcorpus as sc
n = int(float(sys.argv[1])); vocab = int(sys.argv[2])
data = sc.generate(n, sc.SEED_C3, sc.BYTE)
import torch
d = torch.from_numpy(data).cuda()
eng = zb.Engine(0)
eng.set_option("profile", 1)
prev = None
for ms in [1, 10, 30, 100, 300, 1000, 2000, 4000, vocab - 256]:
    eng.set_option("max_steps", ms)
    for rep in range(2):
        m, c = eng.train(None, vocab, device_ptr=d.data_ptr(), n=n)
    st = eng.last_stats
    k = {nm: round(t, 1) for nm, t in zip(zb.KERNEL_CLASSES, st["kernel_ms"]) if t}
    print(json.dumps({"steps": ms, "device_ms": round(st["device_ms"], 1), "scanned_Gslots": round(st["scanned_slots"] / 1e9, 1), "kernel_ms": k}), flush=True)
