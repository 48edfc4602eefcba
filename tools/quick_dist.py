"""tools/quick_dist.py — per-kernel-class profile of sharded training (run under torchrun): weak scaling, bytes per GPU."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
zb = importlib.import_module("zig-bpe_b200")
from tools import synthcorpus as sc
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
n = int(float(sys.argv[1])); vocab = int(sys.argv[2]); prof = int(sys.argv[3]) if len(sys.argv) > 3 else 1
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
box = [zb.Engine.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
eng = zb.Engine(device=lr, rank=rank, world=world, nccl_unique_id=box[0])
for kv in filter(None, os.environ.get("BPE_OPTS", "").split(",")):  # e.g. BPE_OPTS=pdl=0,xchg_impl=1
    eng.set_option(kv.split("=")[0], int(kv.split("=")[1]))
lo, hi = n * rank // world, n * (rank + 1) // world  # strong scaling: n is the whole corpus
shard = sc.generate(hi - lo, sc.SEED_C3, sc.BYTE, offset=lo)
n = hi - lo
d = torch.from_numpy(shard).cuda()
eng.set_option("profile", prof)
if len(sys.argv) > 4:
    eng.set_option("xchg_impl", int(sys.argv[4]))
if rank == 0:
    print("create note:", eng.lib.bpe_last_error(None).decode() or "(peer exchange ready)", flush=True)
for rep in range(2):
    torch.cuda.synchronize(); dist.barrier()
    m, c = eng.train(None, vocab, device_ptr=d.data_ptr(), n=n)
st = eng.last_stats
if rank == 0:
    print(json.dumps({"world": world, "n_per_gpu": n, "device_ms": round(st["device_ms"], 1), "launches": st["kernel_launches"],
                      "kernel_ms": {k: round(t, 1) for k, t in zip(zb.KERNEL_CLASSES, st["kernel_ms"]) if t}}), flush=True)
dist.barrier(); dist.destroy_process_group()
