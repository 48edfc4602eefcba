"""tools/encode_gpu.py <bytes> <vocab> [variants] — train on the synthetic corpus, then time encode (device-resident text and ids)
with the given variants; all must give the same ids. A variant is impl[:option=value[:option=value...]] with impl = encode_impl
(0 default dispatch, 1 one pass per merge, 2 level passes, 3 segment kernel, 4 tile kernel), e.g. 4:encode_tile=4096,2,3."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
zb = importlib.import_module("zig-bpe_b200")
from tools import synthcorpus as sc
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
vocab = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
variants = sys.argv[3].split(",") if len(sys.argv) > 3 else ["4", "2", "3"]
train_bytes = int(float(os.environ.get("TRAIN_BYTES", n)))
eng = zb.Engine(device=0)
d_text = torch.from_numpy(sc.generate(n, sc.SEED_C3, sc.BYTE)).cuda()
m, _ = eng.train(None, vocab, device_ptr=d_text.data_ptr(), n=min(n, train_bytes))
out = {"bytes": n, "merges": int(len(m)), "train_ms": round(eng.last_stats["device_ms"], 1)}
if os.environ.get("ENC_DEBUG"):
    eng.set_option("debug", 1)  # the tile encoder reports ranges / rounds per window on stderr
ref = None
DEFAULTS = {"encode_fuse": 0, "encode_grid": 24, "encode_filter": 1, "encode_geom": 4, "encode_tile": 7936}
for var in variants:
    parts = var.split(":")
    eng.set_option("encode_impl", int(parts[0]))
    opts = dict(DEFAULTS)
    opts.update({k: int(v) for k, v in (p.split("=") for p in parts[1:])})
    for k, v in opts.items():
        eng.set_option(k, v)
    d_ids = torch.empty(n, dtype=torch.int16, device="cuda")
    best = None
    for rep in range(3):
        torch.cuda.synchronize(); t = time.time()
        k = eng.encode_device(d_text.data_ptr(), n, m, d_ids.data_ptr())
        torch.cuda.synchronize(); dt = time.time() - t
        best = dt if best is None else min(best, dt)
    st = eng.last_stats
    out[f"impl{var}"] = {"s": round(best, 4), "GBps": round(n / 1e9 / best, 2), "launches": int(st["kernel_launches"]), "compactions": int(st["compactions"]), "ids": int(k),
                          "encoder": int(st["kernel_calls"][10]), "verdict": int(st["kernel_calls"][11]), "kernel_ms": round(st["kernel_ms"][11], 3), "device_ms": round(st["device_ms"], 3),
                          "fused_groups_ok_redone": [int(st["kernel_calls"][9]), int(st["kernel_calls"][8])]}
    if ref is None:
        ref = d_ids[:k].clone()
    else:
        out["same_ids"] = out.get("same_ids", True) and bool(k == ref.numel() and torch.equal(ref, d_ids[:k]))
    del d_ids
print(json.dumps(out))
