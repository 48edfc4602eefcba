"""tools/c4_run.py — BASELINE config 4: synthetic 10 GB byte corpus (seed 0x5EED0004), vocab 32768, sharded over the
GPUs of one box (run under torchrun). No oracle can check this size, so it reports size-independent properties:
all ranks learn the same merges, ids are strictly new, winning counts never increase, and every shard's
encode -> decode round trip reproduces its bytes.
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 tools/c4_run.py [bytes] [vocab] [encode 0/1] [C4|C5]"""
import hashlib, importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
zb = importlib.import_module("zig-bpe_b200")
from tools import synthcorpus as sc

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000_000
vocab = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
do_encode = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
cfg = sys.argv[4] if len(sys.argv) > 4 else "C4"
SEED = {"C4": sc.SEED_C4, "C5": sc.SEED_C5}[cfg]
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
box = [zb.Engine.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
eng = zb.Engine(device=lr, rank=rank, world=world, nccl_unique_id=box[0])
lo, hi = total * rank // world, total * (rank + 1) // world
n = hi - lo
t = time.time()
pinned = torch.empty(n, dtype=torch.uint8, pin_memory=True)
sc.generate(n, SEED, sc.BYTE, offset=lo, out=pinned.numpy(), nthreads=max(1, (os.cpu_count() or 8) // world))
d_text = pinned.cuda()
gen_s = time.time() - t
torch.cuda.synchronize(); dist.barrier()
t = time.time()
m, c = eng.train(None, vocab, device_ptr=d_text.data_ptr(), n=n)
torch.cuda.synchronize(); dist.barrier()
train_s = time.time() - t
st = dict(eng.last_stats)
ma = np.stack([m["first"], m["second"], m["new_token"]], axis=1)
digest = hashlib.sha256(ma.tobytes() + c.tobytes()).digest()
h = torch.tensor(list(digest[:8]), dtype=torch.int64, device="cuda")
hmin, hmax = h.clone(), h.clone()
dist.all_reduce(hmin, op=dist.ReduceOp.MIN); dist.all_reduce(hmax, op=dist.ReduceOp.MAX)
same = bool(torch.equal(hmin, hmax))
props = {"merges": int(len(m)), "ids_strictly_new": bool((ma[:, 2] == np.arange(256, 256 + len(m))).all() and (ma[:, 0] < ma[:, 2]).all() and (ma[:, 1] < ma[:, 2]).all()),
         "counts_non_increasing": bool((np.diff(c.astype(np.int64)) <= 0).all()), "same_on_all_ranks": same}
enc = None
if do_encode:
    d_ids = torch.empty(n, dtype=torch.int16, device="cuda")  # u16 ids
    torch.cuda.synchronize(); dist.barrier()
    t = time.time()
    n_ids = eng.encode_device(d_text.data_ptr(), n, m, d_ids.data_ptr())
    torch.cuda.synchronize(); dist.barrier()
    enc_s = time.time() - t
    d_back = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    nb = eng.decode_device(d_ids.data_ptr(), n_ids, m, d_back.data_ptr(), n + 64)
    # a token that straddles two shards is emitted by the left one, so compare the concatenation: total bytes must
    # match and every rank's bytes must equal the corpus at its decoded offset
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([nb], dtype=torch.int64, device="cuda"))
    sizes = [int(x) for x in sizes]
    off = sum(sizes[:rank])
    ref = sc.generate(nb, SEED, sc.BYTE, offset=off, nthreads=max(1, (os.cpu_count() or 8) // world))
    ok = bool(torch.equal(d_back[:nb].cpu(), torch.from_numpy(ref)))
    okt = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    enc = {"encode_s": round(enc_s, 2), "input_GBps": round(total / 1e9 / enc_s, 3), "ids_rank0": int(n_ids), "decoded_total_bytes": sum(sizes),
           "round_trip_ok_all_ranks": bool(int(okt)) and sum(sizes) == total}
if rank == 0:
    print(json.dumps({"config": cfg, "bytes": total, "vocab": vocab, "gpus": world, "gen_s": round(gen_s, 1), "train_s": round(train_s, 2),
                      "merges_per_s": round(len(m) / train_s, 1), "device_ms": round(st["device_ms"], 1), "tie_steps": st["tie_steps"],
                      "tie_slow_steps": st["tie_slow_steps"], "compactions": st["compactions"], "launches": st["kernel_launches"],
                      "last_merges": ma[-2:].tolist(), "last_counts": c[-2:].tolist(), **props, "encode": enc}), flush=True)
dist.barrier(); dist.destroy_process_group()
