"""tools/stream_gpu.py — streaming rate of the merge kernel: encode with merges that never match."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
zb = importlib.import_module("zig-bpe_b200")
from tools import synthcorpus as sc
n = int(float(sys.argv[1])); K = int(sys.argv[2])
data = sc.generate(n, sc.SEED_C3, sc.BYTE)
d = torch.from_numpy(data).cuda()
out = torch.empty(n, dtype=torch.int16, device="cuda")
eng = zb.Engine(0)
A = int(sys.argv[3]) if len(sys.argv) > 3 else -1
merges = [((60000 + i) if A < 0 else A, 61000 + i, 62000 + i) for i in range(K)]
if A >= 0: print("A =", A, "freq", float((data == A).mean()))
for rep in range(2):
    k = eng.encode_device(d.data_ptr(), n, merges, out.data_ptr())
    st = eng.last_stats
    print(json.dumps({"n": n, "passes": K, "device_ms": round(st["device_ms"], 2), "ms_per_pass": round(st["device_ms"] / K, 4),
                      "GBps_per_pass": round(2 * n / 1e9 / (st["device_ms"] / K / 1e3), 1), "launches": st["kernel_launches"]}), flush=True)
