"""tools/dist_check.py — multi-GPU parity on real GPUs (run under torchrun): sharded training over NCCL
must return the merges a single GPU learns from the whole corpus.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py [bytes] [vocab]"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
zb = importlib.import_module("zig-bpe_b200")
from tools import synthcorpus as sc

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
n_total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
vocab = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
box = [zb.Engine.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
eng = zb.Engine(device=lr, rank=rank, world=world, nccl_unique_id=box[0])
for kv in filter(None, os.environ.get("BPE_OPTS", "").split(",")):  # e.g. BPE_OPTS=pdl=0,xchg_impl=1
    eng.set_option(kv.split("=")[0], int(kv.split("=")[1]))
lo, hi = n_total * rank // world, n_total * (rank + 1) // world
shard = sc.generate(hi - lo, sc.SEED_C3, sc.BYTE, offset=lo)
for rep in range(2):
    torch.cuda.synchronize(); dist.barrier()
    t = time.time()
    m, c = eng.train(shard, vocab)
    torch.cuda.synchronize(); dist.barrier()
    dt = time.time() - t
st = eng.last_stats
# sharded encode: the concatenation over ranks must equal the single-GPU encoding
ids = eng.encode(shard, m)
sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
dist.all_gather(sizes, torch.tensor([len(ids)], dtype=torch.int64, device="cuda"))
sizes = [int(x) for x in sizes]
pad = torch.zeros(max(sizes), dtype=torch.int32, device="cuda")
pad[: len(ids)] = torch.from_numpy(ids.astype(np.int32)).cuda()
allids = [torch.zeros(max(sizes), dtype=torch.int32, device="cuda") for _ in range(world)]
dist.all_gather(allids, pad)
ok = None
if rank == 0:
    single = zb.Engine(device=lr)
    full = sc.generate(n_total, sc.SEED_C3, sc.BYTE)
    t = time.time(); ms, cs = single.train(full, vocab); dt1 = time.time() - t
    ok = bool(np.array_equal(m, ms) and np.array_equal(c, cs))
    ids_single = single.encode(full, ms)
    cat = np.concatenate([allids[r][: sizes[r]].cpu().numpy().astype(np.uint16) for r in range(world)])
    enc_ok = bool(np.array_equal(cat, ids_single))
    ok = ok and enc_ok
    print(json.dumps({"world": world, "encode_identical": enc_ok, "bytes": n_total, "vocab": vocab, "merges": len(m), "sharded_s": round(dt, 3), "single_gpu_s": round(dt1, 3),
                      "identical_to_single_gpu": ok, "tie_steps": st["tie_steps"], "tie_slow": st["tie_slow_steps"], "launches": st["kernel_launches"]}), flush=True)
flag = torch.tensor([1 if (ok is None or ok) else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 1)
