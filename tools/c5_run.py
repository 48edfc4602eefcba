"""tools/c5_run.py — BASELINE config 5: encode-only, synthetic 10 GB byte corpus (seed 0x5EED0005) with
the longest merge list the format allows, trained on the FIRST 1 GB of that corpus (vocab 65535 -> up to 65,279 merges; a
merges.txt cannot hold more than 65,280). Run it on 1 GPU with plain python, on N GPUs under torchrun; every rank holds the
r-th contiguous slice of the corpus. Reports encode input GB/s (device-resident), its roofline fraction (n + 2 n_out), the
N-invariant sha256 of all ids in shard order, the decode round trip, and the streaming host-buffer path on a slice.
  python tools/c5_run.py [bytes] [train_bytes] [vocab]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 tools/c5_run.py [bytes] [train_bytes] [vocab]"""
import hashlib, importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
zb = importlib.import_module("zig-bpe_b200")
from tools import synthcorpus as sc

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000_000
train_bytes = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000_000
vocab = int(sys.argv[3]) if len(sys.argv) > 3 else 65535
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
uid = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    box = [zb.Engine.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = box[0]
eng = zb.Engine(device=lr, rank=rank, world=world, nccl_unique_id=uid)
nthreads = max(1, (os.cpu_count() or 8) // world)
PIECE = 1 << 30
stage = torch.empty(PIECE, dtype=torch.uint8, pin_memory=True)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def load(lo, hi):
    """bytes [lo, hi) of the C5 corpus into a device tensor, generated piecewise through one pinned staging buffer"""
    d = torch.empty(hi - lo, dtype=torch.uint8, device=dev)
    for o in range(lo, hi, PIECE):
        k = min(PIECE, hi - o)
        sc.generate(k, sc.SEED_C5, sc.BYTE, offset=o, out=stage.numpy(), nthreads=nthreads)
        d[o - lo:o - lo + k].copy_(stage[:k])
    return d


# ---- the merge list: trained on the first train_bytes of the corpus, sharded over the ranks ----
t = time.time()
tlo, thi = train_bytes * rank // world, train_bytes * (rank + 1) // world
d_train = load(tlo, thi)
barrier()
t1 = time.time()
m, c = eng.train(None, vocab, device_ptr=d_train.data_ptr(), n=thi - tlo)
barrier()
train_s = time.time() - t1
del d_train
merges_sha = hashlib.sha256("".join(f"{int(a)},{int(b)},{int(z)}\n" for a, b, z in zip(m["first"], m["second"], m["new_token"])).encode()).hexdigest()

# ---- encode the whole corpus ----
lo, hi = total * rank // world, total * (rank + 1) // world
n = hi - lo
d_text = load(lo, hi)
gen_s = time.time() - t - train_s
d_ids = torch.empty(n, dtype=torch.int16, device=dev)
if os.environ.get("ENC_DEBUG"):
    eng.set_option("debug", 1)
times = []
for rep in range(2):
    barrier()
    t2 = time.time()
    n_ids = eng.encode_device(d_text.data_ptr(), n, m, d_ids.data_ptr())
    barrier()
    times.append(time.time() - t2)
est = dict(eng.last_stats)
tt = torch.tensor([min(times), float(n_ids)], dtype=torch.float64, device=dev)
tmax, tsum = tt.clone(), tt.clone()
if world > 1:
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
enc_s, ids_total = float(tmax[0]), int(tsum[1])

# ---- N-invariant hash of the ids (rank 0 hashes the ranks' ids in shard order, 256 MiB at a time) ----
CH = 1 << 27
def hash_into(h, buf_u8):
    for o in range(0, buf_u8.numel(), 2 * CH):
        h.update(buf_u8[o:o + 2 * CH].cpu().numpy().tobytes())
ids_sha = None
mine = d_ids[:n_ids].contiguous().view(torch.uint8)
if rank == 0:
    h = hashlib.sha256()
    hash_into(h, mine)
    for r in range(1, world):
        cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        dist.recv(cnt, src=r)
        buf = torch.empty(2 * int(cnt[0]), dtype=torch.uint8, device=dev)
        dist.recv(buf, src=r)
        hash_into(h, buf)
        del buf
    ids_sha = h.hexdigest()
else:
    dist.send(torch.tensor([n_ids], dtype=torch.int64, device=dev), dst=0)
    dist.send(mine, dst=0)

# ---- decode round trip (a token that straddles two shards belongs to the left one: compare at the decoded offset) ----
d_back = torch.empty(n + (1 << 16), dtype=torch.uint8, device=dev)
barrier()
t3 = time.time()
nb = eng.decode_device(d_ids.data_ptr(), n_ids, m, d_back.data_ptr(), n + (1 << 16))
barrier()
dec_s = time.time() - t3
sizes = [nb]
if world > 1:
    g = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(g, torch.tensor([nb], dtype=torch.int64, device=dev))
    sizes = [int(x) for x in g]
off = sum(sizes[:rank])
ok = True
for o in range(0, nb, PIECE):
    k = min(PIECE, nb - o)
    sc.generate(k, sc.SEED_C5, sc.BYTE, offset=off + o, out=stage.numpy(), nthreads=nthreads)
    ok = ok and bool(torch.equal(d_back[o:o + k], stage[:k].to(dev)))
okt = torch.tensor([1 if ok else 0], device=dev)
if world > 1:
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
del d_back

# ---- host buffers, streaming (bpe_encode): the first 2 GB of this rank's slice from pinned memory ----
e2e = None
if world == 1:
    k = min(n, 2 * PIECE)
    host = torch.empty(k, dtype=torch.uint8, pin_memory=True)
    sc.generate(k, sc.SEED_C5, sc.BYTE, offset=lo, out=host.numpy(), nthreads=nthreads)
    eng.set_option("debug", 0)
    eng.encode(host.numpy()[: 1 << 28], m)  # warm-up (tables, buffers)
    t4 = time.time()
    ids_h = eng.encode(host.numpy(), m)
    e2e_s = time.time() - t4
    st = dict(eng.last_stats)
    same = bool(np.array_equal(ids_h[:1000000], d_ids[:1000000].cpu().numpy().view(np.uint16)))
    e2e = {"bytes": k, "s": round(e2e_s, 3), "GBps": round(k / 1e9 / e2e_s, 2), "chunks": int(st["kernel_calls"][9]), "h2d_bytes": k, "d2h_bytes": 2 * len(ids_h),
           "first_1M_ids_equal_resident_run": same, "note": "python wrapper allocates and copies the result array inside the timed region"}

if rank == 0:
    peak = 6552.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    alg = total + 2 * ids_total
    print(json.dumps({"config": "C5", "bytes": total, "gpus": world, "merges": int(len(m)), "train_bytes": train_bytes, "train_s": round(train_s, 2),
                      "merges_sha256": merges_sha, "gen_s": round(gen_s, 1),
                      "encode_s": round(enc_s, 3), "encode_input_GBps": round(total / 1e9 / enc_s, 3), "ids_total": ids_total,
                      "encoder": {0: "level passes", 1: "segment kernel", 2: "tile kernel"}.get(int(est["kernel_calls"][10]), "?"),
                      "encode_launches_rank0": int(est["kernel_launches"]),
                      "roofline": {"bound": "hbm", "achieved_GBps_per_gpu": round(alg / 1e9 / enc_s / world, 2), "peak": peak, "frac": alg / 1e9 / enc_s / world / peak,
                                   "algorithmic_bytes": alg},
                      "encode_ids_sha256": ids_sha, "decode_s": round(dec_s, 3), "decode_GBps": round(total / 1e9 / dec_s, 2),
                      "round_trip_ok_all_ranks": bool(int(okt)) and sum(sizes) == total, "host_streaming": e2e}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
