#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 BPE engine (contract: see the task prompt / DESIGN.md).

Metric (BASELINE.json): train merges/sec on the 1 GB synthetic byte corpus, vocab 8192 (config C3,
SURVEY.md §8d). One "step" = one complete training run (7,936 merges) over the corpus.

  value : merges/sec with the corpus already resident in HBM (bpe_train_device), whole job
  e2e   : the same through the reference-facing call with HOST buffers (bpe_train: H2D of the
          corpus inside the timed region, merge list read back)
  roofline: the merge kernel (dominant) — algorithmic bytes = 2 B x token slots scanned per launch
          (u16 tokens) / its CUDA-event time inside the library's own stream
  cpu_baseline: the oracle's verbatim restatement of the reference, single thread, first K merge
          steps of the same corpus on this box's host cores (a reported baseline, not the target)

python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--bytes B] [--vocab V]
Multi-GPU: torchrun launches one rank per GPU; the corpus is sharded by contiguous chunks.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

GB = 1_000_000_000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bytes", type=int, default=GB, help="corpus bytes of the whole job (BASELINE config 3: 1 GB); N GPUs hold contiguous shards of bytes/N")
    ap.add_argument("--vocab", type=int, default=8192)
    ap.add_argument("--cpu-steps", type=int, default=2, help="merge steps timed for the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-encode", action="store_true")
    ap.add_argument("--max-steps", type=int, default=0, help="profiling aid: stop every training run after this many merges")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "500", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU algorithm (oracle restatement; the Zig reference cannot be
    built here: no zig toolchain) on the host cores, bounded sample = first `cpu_steps` merge steps per step."""
    if rank != 0:
        return
    from oracle import oracle_py as ora
    from tools import synthcorpus as sc
    n = args.bytes
    data = sc.generate(n, sc.SEED_C3, sc.BYTE)
    ksteps = args.cpu_steps
    times = []
    for i in range(args.warmup + args.steps):
        t = time.perf_counter()
        m, _ = ora.train(data, args.vocab, max_steps=ksteps, fast=False)
        dt = time.perf_counter() - t
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    val = args.steps * ksteps / total
    sample = f"first {ksteps} merge steps (full recount each, as the reference does) of the {n}-byte corpus per step"
    line = {"impl": "reference", "metric": "train_merges_per_sec", "value": val, "unit": "merges/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u16", "data": "synthetic (synthcorpus-v1 byte variant, seed 0x5EED0003)",
            "config": {"workload": f"C3: {n}-byte synthetic byte corpus, vocab {args.vocab} (train)", "bytes": n, "vocab": args.vocab},
            "cpu_baseline": {"value": val, "unit": "merges/s", "cores": 1, "kind": "port", "sample": sample,
                             "host_cores_available": os.cpu_count()},
            "e2e": {"value": val, "unit": "merges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def emit(line: dict):
    """Print the one JSON line on the real stdout (libraries such as NCCL write banners to fd 1, which is
    pointed at stderr for the duration of the run)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    import torch.distributed as dist
    zb = importlib.import_module("zig-bpe_b200")
    from tools import synthcorpus as sc

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        box = [zb.Engine.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    eng = zb.Engine(device=local_rank, rank=rank, world=world, nccl_unique_id=uid)

    # strong scaling (BASELINE config 3: the same 1 GB corpus on 1/2/4/8 GPUs): rank r holds the r-th contiguous shard
    total_bytes = args.bytes
    lo, hi = total_bytes * rank // world, total_bytes * (rank + 1) // world
    n = hi - lo
    pinned = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host = pinned.numpy()
    sc.generate(n, sc.SEED_C3, sc.BYTE, offset=lo, out=host)
    d_text = pinned.to(dev, non_blocking=False)
    n_merges_target = args.vocab - 256

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        m, c = eng.train(None, args.vocab, device_ptr=d_text.data_ptr(), n=n)
        return m, c, dict(eng.last_stats)

    def step_host():
        m, c = eng.train(host, args.vocab)
        return m, c, dict(eng.last_stats)

    if args.max_steps:
        eng.set_option("max_steps", args.max_steps)
    if os.environ.get("BPE_DEBUG"):
        eng.set_option("debug", int(os.environ["BPE_DEBUG"]))
    eng.set_option("profile", 3)  # CUDA-event marks around the merge kernel of every 8th step (sampled: ~1 us/step of overhead)
    for _ in range(args.warmup):
        merges, counts, st = step_device()
    clocks = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    launches = 0; merge_ms = 0.0; merge_calls = 0; scanned = 0; sampled_slots = 0.0; dev_ms = 0.0; got = 0
    for _ in range(args.steps):
        merges, counts, st = step_device()
        launches += st["kernel_launches"]; merge_ms += st["kernel_ms"][5]; merge_calls += st["kernel_calls"][5]
        scanned += st["scanned_slots"]; sampled_slots += st["kernel_ms"][10]; dev_ms += st["device_ms"]; got += len(merges)
    ev1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1000
    clk = clocks.stop() if rank == 0 else None
    # the library runs on its own stream; its CUDA-event time (device_ms) is the device-side clock of the
    # step, the host wall clock around the barriers is the fallback when they disagree
    t = torch.tensor([max(dev_ms, 0.0), wall_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max = float(t[0]), float(t[1])
    timed_ms = wall_ms_max
    value = got / (timed_ms / 1000.0)

    # e2e through the host-buffer entry point
    e2e = None
    if not args.no_e2e:
        eng.set_option("profile", 0)
        step_host()
        barrier()
        t1 = time.perf_counter()
        got_e = 0
        ksteps = max(1, min(args.steps, 2))
        for _ in range(ksteps):
            m2, c2, st2 = step_host()
            got_e += len(m2)
        barrier()
        e_ms = (time.perf_counter() - t1) * 1000
        te = torch.tensor([e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": got_e / (float(te[0]) / 1000.0), "unit": "merges/s", "h2d_bytes_per_step": total_bytes,
               "d2h_bytes_per_step": int(len(m2)) * (6 + 8), "steps": ksteps}
        assert np.array_equal(m2, merges), "host-buffer and device-resident runs disagree"

    # encode (the other half of BASELINE.json's metric): every GPU encodes its own shard with the merges just learned
    enc = None
    if not args.no_encode:
        eng.set_option("profile", 0)
        d_out = torch.empty(n, dtype=torch.int16, device=dev)
        n_ids = eng.encode_device(d_text.data_ptr(), n, merges, d_out.data_ptr())  # warm-up
        barrier()
        t3 = time.perf_counter()
        n_ids = eng.encode_device(d_text.data_ptr(), n, merges, d_out.data_ptr())
        est = dict(eng.last_stats)
        barrier()
        enc_ms = (time.perf_counter() - t3) * 1000
        te = torch.tensor([enc_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        enc = {"metric": "encode_input_GB_per_sec", "value": total_bytes / 1e9 / (float(te[0]) / 1000.0), "unit": "GB/s", "merges": int(len(merges)),
               "bytes_rank0": n, "ids_out_rank0": int(n_ids), "gpu_launches": int(est["kernel_launches"]),
               "scan_GBps_per_gpu": est["scanned_slots"] * 2 / 1e9 / (est["device_ms"] / 1000.0),
               "note": "level-scheduled passes: merges that commute share one pass (exact for any list, same ids as one pass per merge); algorithmic bytes n + 2*n_out"}
        del d_out
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    alg_bytes = scanned * 2  # u16 slots, all launches
    samp_bytes = sampled_slots * 2  # the launches whose duration was measured (every 8th merge step)
    traffic = None  # DRAM bytes per launch from the committed ncu --set full capture (ratio to algorithmic bytes)
    tp = os.path.join(ROOT, "profiles", "r01_merge_traffic.json")
    if os.path.exists(tp) and merge_calls:
        traffic = json.load(open(tp))["traffic_over_algorithmic"] * samp_bytes / merge_calls
    achieved = samp_bytes / 1e9 / (merge_ms / 1000.0) if merge_ms > 0 else None
    est_merge_ms_all = alg_bytes / 1e9 / achieved * 1000.0 if achieved else None  # all launches at the sampled rate
    line = {
        "metric": "train_merges_per_sec", "value": value, "unit": "merges/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": timed_ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u16", "data": "synthetic (synthcorpus-v1 byte variant, seed 0x5EED0003)",
        "config": {"workload": f"C3: {total_bytes}-byte synthetic byte corpus, vocab {args.vocab} (train, {n_merges_target} merges/step), "
                               f"contiguous shards over {world} GPU(s)",
                   "bytes": total_bytes, "bytes_rank0": n, "vocab": args.vocab, "merges_per_step": int(len(merges)),
                   "l2": "the resident sequence (2 B x token slots) is larger than the 126 MB L2 for all but the last steps at N>=4; consecutive "
                         "passes alternate direction so the tail of one pass is reused from L2 by the next; no explicit flush",
                   "timing": "host clock between barrier+synchronize pairs, max over ranks; the library's CUDA-event device_ms is reported beside it"},
        "device_ms_per_step": dev_ms_max / args.steps,
        "gpu_launches": int(launches),
        "clocks": clk,
        "e2e": e2e,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                     "traffic": traffic, "kernel": "merge_kernel<u16>", "peak_source": peak_src,
                     "bytes_per_launch": samp_bytes / max(merge_calls, 1), "avg_launch_ms": merge_ms / max(merge_calls, 1),
                     "launches_timed": int(merge_calls), "sampling": "every 8th merge step, CUDA events on the library's stream",
                     "kernel_share_of_step": est_merge_ms_all / dev_ms if dev_ms and est_merge_ms_all else None,
                     "whole_step_scan_GBps": alg_bytes / 1e9 / (dev_ms / 1000.0) if dev_ms else None},
        "tie_steps": int(st["tie_steps"]), "tie_slow_steps": int(st["tie_slow_steps"]), "compactions": int(st["compactions"]),
        "encode": enc,
    }
    if not args.no_cpu_baseline and world == 1:
        from oracle import oracle_py as ora
        t2 = time.perf_counter()
        om, oc = ora.train(host, args.vocab, max_steps=args.cpu_steps, fast=False)
        cpu_s = time.perf_counter() - t2
        k = len(om)
        ok = bool(np.array_equal(np.stack([merges["first"], merges["second"], merges["new_token"]], axis=1)[:k], om)
                  and np.array_equal(counts[:k], oc))
        line["cpu_baseline"] = {"value": k / cpu_s, "unit": "merges/s", "cores": 1, "kind": "port",
                                "sample": f"first {k} merge steps (verbatim per-step recount) of the same {n}-byte corpus, {cpu_s:.1f} s",
                                "host_cores_available": os.cpu_count(), "matches_gpu_merges": ok}
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
