"""tools/fuse_gpu_check.py [bytes] [vocab] — the fused level groups of encode ("encode_fuse", bpe_groupenc.cuh) on cuda:0 against the
level passes: identical ids required, timings printed. Run in its own process (tests/test_gpu_parity.py does) because the
kernel had not been run on a GPU when the round's last lease ended. Exit code 0: identical."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
zb = importlib.import_module("zig-bpe_b200")
from tools import synthcorpus as sc

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
vocab = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
data = sc.generate(n, sc.SEED_C3, sc.BYTE)
eng = zb.Engine(0)
m, _ = eng.train(data, vocab)
out = {"bytes": n, "merges": len(m)}
eng.set_option("encode_impl", 2)
ref = None
ok = True
for fuse in (0, 2, 4, 8):
    eng.set_option("encode_fuse", fuse)
    best = None
    for _ in range(2):
        t = time.time(); ids = eng.encode(data, m); dt = time.time() - t
        best = dt if best is None else min(best, dt)
    st = eng.last_stats
    if ref is None:
        ref = ids
    same = bool(np.array_equal(ids, ref))
    ok = ok and same
    out[f"fuse{fuse}"] = {"s_host_buffers": round(best, 4), "device_ms": round(st["device_ms"], 2), "launches": int(st["kernel_launches"]),
                          "groups_ok": int(st["kernel_calls"][9]), "groups_redone": int(st["kernel_calls"][8]), "same_ids": same}
print(json.dumps(out))
sys.exit(0 if ok else 1)
