"""tools/encode_gpu.py <bytes> <vocab> [variants] — train on the synthetic corpus, then time encode (device-resident text and ids)
with the segment-resident kernel (encode_impl = 0), the level schedule (2) and one pass per merge (1); all must give the same ids."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
zb = importlib.import_module("zig-bpe_b200")
from tools import synthcorpus as sc
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
vocab = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
# variants: impl[:grid[:filter[:geom]]] — encode_impl; for the level schedule encode_grid (CTAs per SM; 0 = one CTA per
# tile) and encode_filter; for the segment kernel encode_geom (0: 64+2x32 bytes per thread, 1: 128+2x32, 2: 96+2x32)
variants = sys.argv[3].split(",") if len(sys.argv) > 3 else ["0:6:0:0", "0:6:0:1", "0:6:0:2", "2:6"]
eng = zb.Engine(device=0)
d_text = torch.from_numpy(sc.generate(n, sc.SEED_C3, sc.BYTE)).cuda()
m, _ = eng.train(None, vocab, device_ptr=d_text.data_ptr(), n=n)
out = {"bytes": n, "merges": int(len(m)), "train_ms": round(eng.last_stats["device_ms"], 1)}
ref = None
for var in variants:
    impl = int(var.split(":")[0])
    eng.set_option("encode_impl", impl)
    parts = var.split(":")
    if len(parts) > 1:
        eng.set_option("encode_grid", int(parts[1]))
    eng.set_option("encode_filter", int(parts[2]) if len(parts) > 2 else 0)
    eng.set_option("encode_geom", int(parts[3]) if len(parts) > 3 else 0)
    d_ids = torch.empty(n, dtype=torch.int16, device="cuda")
    best = None
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time()
        k = eng.encode_device(d_text.data_ptr(), n, m, d_ids.data_ptr())
        torch.cuda.synchronize(); dt = time.time() - t
        best = dt if best is None else min(best, dt)
    st = eng.last_stats
    out[f"impl{var}"] = {"s": round(best, 4), "GBps": round(n / 1e9 / best, 2), "launches": int(st["kernel_launches"]), "compactions": int(st["compactions"]), "ids": int(k),
                          "seg_path": int(st["kernel_calls"][11]), "seg_kernel_ms": round(st["kernel_ms"][11], 3), "device_ms": round(st["device_ms"], 3)}
    if ref is None:
        ref = d_ids[:k].clone()
    else:
        out["same_ids"] = bool(k == ref.numel() and torch.equal(ref, d_ids[:k]))
print(json.dumps(out))
