#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 BPE engine (contract: see the task prompt / DESIGN.md).

Metric (BASELINE.json): train merges/sec on the 1 GB synthetic byte corpus, vocab 8192 (config C3,
SURVEY.md §8d). One "step" = one complete training run (7,936 merges) over the corpus.

  value : merges/sec with the corpus already resident in HBM (bpe_train_device), whole job
  e2e   : the same through the reference-facing call with HOST buffers (bpe_train: H2D of the
          corpus inside the timed region, merge list read back)
  parity: sha256 of the learned merges as merges.txt bytes and of the encoded ids (concatenated over the ranks in
          shard order), the oracle-derived prefix hash of this configuration when tests/golden holds one, and at
          N > 1 rank 0's own single-GPU training + encoding of the whole corpus compared with the sharded result
  gpu_same_k_steps: the GPU's time for the same first K merge steps the CPU baseline is timed on
  encode / decode: input GB/s device-resident, their rooflines (n + 2*n_out, 2*n_tok + n_bytes) and e2e from host buffers
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
    ap.add_argument("--no-single-check", action="store_true", help="N > 1: skip rank 0's single-GPU retrain / re-encode of the whole corpus")
    ap.add_argument("--max-steps", type=int, default=0, help="profiling aid: stop every training run after this many merges")
    return ap.parse_args()


class ClockSampler:
    """SM clocks / throttle reasons during the timed region (B200_PROFILING.md recipe). NVML is polled in-process once a
    second (round 1 ran `nvidia-smi -lms 500` beside the timed loop; on some boxes its driver queries delayed the kernel
    launches of the loop by ~10 %); `nvidia-smi -lms 2000` is the fallback when NVML cannot be loaded."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        self.stop_flag = False
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = None
            try:
                import torch
                uuid = "GPU-" + str(torch.cuda.get_device_properties(self.gpu).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}

            def poll():
                while not self.stop_flag:
                    try:
                        sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                        self.rows.append([str(self.gpu), str(sm), str(self.mx), str(pw)] + ["Active" if (r & b) else "Not Active" for b in bits.values()])
                    except Exception:
                        pass
                    time.sleep(1.0)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            self.source = "nvml, 1 s"
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "2000", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            self.source = "nvidia-smi -lms 2000"
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        self.stop_flag = True
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
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm),
                "source": getattr(self, "source", None)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload(total_bytes, vocab, world):
    """config.workload — the same string on both arms so the driver can see that they ran the same configuration"""
    return (f"C3: {total_bytes}-byte synthetic byte corpus, vocab {vocab} (train, {vocab - 256} merges/step), "
            f"contiguous shards over {world} GPU(s)")


def pin_to_one_core():
    """CPU baseline = the reference's single thread, pinned (`taskset -c 0`). Returns the old mask."""
    try:
        old = os.sched_getaffinity(0)
        os.sched_setaffinity(0, {min(old)})
        return old
    except Exception:
        return None


def unpin(old):
    if old:
        try:
            os.sched_setaffinity(0, old)
        except Exception:
            pass


def merges_text(m):
    """the merges.txt bytes the reference would write (basic_tokenizer.zig:328) — what merges_sha256 hashes"""
    return "".join(f"{int(a)},{int(b)},{int(c)}\n" for a, b, c in zip(m["first"], m["second"], m["new_token"])).encode()


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
    old = pin_to_one_core()
    for i in range(args.warmup + args.steps):
        t = time.perf_counter()
        m, _ = ora.train(data, args.vocab, max_steps=ksteps, fast=False)
        dt = time.perf_counter() - t
        if i >= args.warmup:
            times.append(dt)
    unpin(old)
    total = sum(times)
    val = args.steps * ksteps / total
    sample = (f"first {ksteps} merge steps (full recount each, as the reference does) of the {n}-byte corpus per step, 1 thread pinned to one core; "
              f"these are the most expensive steps of a run (the sequence only shrinks), see gpu_same_k_steps in the other arm")
    line = {"impl": "reference", "metric": "train_merges_per_sec", "value": val, "unit": "merges/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u16", "data": "synthetic (synthcorpus-v1 byte variant, seed 0x5EED0003)",
            "config": {"workload": workload(n, args.vocab, args.gpus), "bytes": n, "vocab": args.vocab},
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
    import hashlib
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
    for kv in filter(None, os.environ.get("BPE_OPTS", "").split(",")):  # measurement aid: BPE_OPTS=merge_impl=1,...
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    eng.set_option("profile", 3)  # CUDA-event marks around the merge kernel of every 32nd step (sampled: ~1 us/step of overhead)
    for _ in range(args.warmup):
        merges, counts, st = step_device()
    clocks = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        clocks.start()
    t0 = time.perf_counter()
    launches = 0; merge_ms = 0.0; merge_calls = 0; scanned = 0; sampled_slots = 0.0; dev_ms = 0.0; got = 0; lib_ms = 0.0
    for _ in range(args.steps):
        merges, counts, st = step_device()
        lib_ms += st["total_ms"]
        launches += st["kernel_launches"]; merge_ms += st["kernel_ms"][5]; merge_calls += st["kernel_calls"][5]
        scanned += st["scanned_slots"]; sampled_slots += st["kernel_ms"][10]; dev_ms += st["device_ms"]; got += len(merges)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1000
    clk = clocks.stop() if rank == 0 else None
    # the library runs on its own stream; its CUDA-event time (device_ms) is the device-side clock of the
    # step, the host wall clock around the barriers is what `value` uses (it can only be larger)
    t = torch.tensor([max(dev_ms, 0.0), wall_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max = float(t[0]), float(t[1])
    timed_ms = wall_ms_max
    value = got / (timed_ms / 1000.0)

    # GPU time for the same first K merge steps the CPU baseline is timed on (like-for-like beside cpu_baseline)
    same_k = None
    if not args.max_steps:
        eng.set_option("max_steps", args.cpu_steps)
        eng.set_option("profile", 0)
        step_device()
        barrier()
        tk = time.perf_counter()
        mk, ck, stk = step_device()
        barrier()
        same_k = {"steps": int(len(mk)), "wall_ms": (time.perf_counter() - tk) * 1000, "device_ms": stk["device_ms"],
                  "note": "initial pair count + the first K merge steps at full length, corpus resident in HBM"}
        eng.set_option("max_steps", 0)

    # e2e through the host-buffer entry point
    e2e = None
    if not args.no_e2e:
        eng.set_option("profile", 0)
        step_host()
        barrier()
        t1 = time.perf_counter()
        got_e = 0
        ksteps = max(1, min(args.steps, 3))
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

    # ---- parity evidence in the record itself: the learned merge list as the reference would write it ----
    mtext = merges_text(merges)
    merges_sha = hashlib.sha256(mtext).hexdigest()
    parity = {"merges_sha256": merges_sha, "merges": int(len(merges))}
    if world > 1:
        shas = [None] * world
        dist.all_gather_object(shas, merges_sha)
        parity["identical_on_all_ranks"] = len(set(shas)) == 1
    gold = os.path.join(ROOT, "tests", "golden", "big_sha256.json")
    if os.path.exists(gold) and total_bytes == GB and args.vocab == 8192 and not args.max_steps:
        g = json.load(open(gold)).get("c3")
        if g:  # oracle-derived prefix hashes of this very configuration (tools/make_golden_big.py)
            lines = mtext.decode().splitlines(keepends=True)
            ks = sorted(int(k) for k in g["prefix_sha256"] if int(k) <= len(lines))
            if ks:
                k = ks[-1]
                parity["oracle_prefix_merges"] = k
                parity["matches_oracle_prefix"] = hashlib.sha256("".join(lines[:k]).encode()).hexdigest() == g["prefix_sha256"][str(k)]

    # encode (the other half of BASELINE.json's metric): every GPU encodes its own shard with the merges just learned
    enc = None
    dec = None
    if not args.no_encode:
        eng.set_option("profile", 0)
        d_out = torch.empty(n, dtype=torch.int16, device=dev)
        n_ids = eng.encode_device(d_text.data_ptr(), n, merges, d_out.data_ptr())  # warm-up
        enc_times = []
        for _ in range(3):
            barrier()
            t3 = time.perf_counter()
            n_ids = eng.encode_device(d_text.data_ptr(), n, merges, d_out.data_ptr())
            est = dict(eng.last_stats)
            barrier()
            enc_times.append((time.perf_counter() - t3) * 1000)
        te = torch.tensor([float(np.median(enc_times)), float(n_ids)], dtype=torch.float64, device=dev)
        tsum = te.clone()
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        enc_s = float(te[0]) / 1000.0
        ids_total = int(tsum[1])
        peak_e, _ = peaks()
        alg = total_bytes + 2 * ids_total  # read every input byte once, write every output id once (u16 ids)
        enc = {"metric": "encode_input_GB_per_sec", "value": total_bytes / 1e9 / enc_s, "unit": "GB/s", "merges": int(len(merges)),
               "bytes_rank0": n, "ids_out_rank0": int(n_ids), "ids_out_total": ids_total, "gpu_launches": int(est["kernel_launches"]),
               "ms": enc_s * 1000, "runs": 3, "path": {0: "level passes", 1: "segment/tile-resident kernel", 2: "level passes after a seam fallback"}.get(int(est["kernel_calls"][11]), "?"),
               "roofline": {"bound": "hbm", "achieved": alg / 1e9 / enc_s / world, "peak": peak_e, "unit": "GB/s", "frac": alg / 1e9 / enc_s / world / peak_e,
                            "algorithmic_bytes": alg, "note": "n + 2*n_out bytes per GPU-second, per GPU"}}
        # e2e encode: host bytes in, host ids out (H2D + D2H inside)
        if not args.no_e2e:
            # (ids land in a page-locked buffer allocated before the timed region, like the corpus: a fresh pageable 2 GB
            # array would put its page faults into the measurement)
            try:
                ids_pinned = torch.empty(max(n, 1), dtype=torch.int16, pin_memory=True).numpy().view(np.uint16)
            except Exception:  # no page-locked memory left: let the mirror allocate a pageable array
                ids_pinned = None
            barrier()
            t4 = time.perf_counter()
            ids_h = eng.encode(host, merges, out=ids_pinned)
            barrier()
            te2 = torch.tensor([(time.perf_counter() - t4) * 1000], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te2, op=dist.ReduceOp.MAX)
            enc["e2e"] = {"value": total_bytes / 1e9 / (float(te2[0]) / 1000.0), "unit": "GB/s", "h2d_bytes_per_step": total_bytes, "d2h_bytes_per_step": 2 * ids_total}
            assert len(ids_h) == n_ids
        # The fused level groups of encode (bpe_groupenc.cuh, "encode_fuse") were written after the round's last GPU lease and are
        # off by default; this is their first run on a GPU, in a process of its own so that nothing it does can touch the numbers
        # above: same corpus, same merges, device-resident, level passes vs groups of 4 / 8 levels, identical ids required.
        if world == 1 and not args.max_steps and os.environ.get("BPE_BENCH_FUSED", "1") != "0":
            try:
                fr = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "encode_gpu.py"), str(total_bytes), str(args.vocab),
                                    "2,2:encode_fuse=4,2:encode_fuse=8"], capture_output=True, text=True, timeout=240)
                last = [ln for ln in fr.stdout.splitlines() if ln.startswith("{")]
                enc["fused_groups_experimental"] = json.loads(last[-1]) if (fr.returncode == 0 and last) else {
                    "error": f"rc={fr.returncode}", "stderr_tail": fr.stderr[-300:]}
            except Exception as e:  # noqa: BLE001 — never let the experiment break the bench line
                enc["fused_groups_experimental"] = {"error": repr(e)[:300]}
        # N-invariant hash of the whole encoding: rank 0 hashes the ranks' ids in shard order
        ids_dev = d_out[:n_ids]
        if rank == 0:
            h = hashlib.sha256(ids_dev.cpu().numpy().tobytes())
            for r in range(1, world):
                cnt = torch.zeros(1, dtype=torch.int64, device=dev)
                dist.recv(cnt, src=r)
                buf = torch.empty(2 * int(cnt[0]), dtype=torch.uint8, device=dev)  # (NCCL has no 16-bit integer type)
                dist.recv(buf, src=r)
                h.update(buf.cpu().numpy().tobytes())
                del buf
            parity["encode_ids_sha256"] = h.hexdigest()
            parity["encode_ids"] = ids_total
        else:
            dist.send(torch.tensor([n_ids], dtype=torch.int64, device=dev), dst=0)
            dist.send(ids_dev.contiguous().view(torch.uint8), dst=0)
        # decode of this rank's ids back to its shard (round trip) + throughput
        # (a token that straddles two shards belongs to the left one, so a rank's ids decode to its shard shifted by a few bytes at N > 1)
        cap = n + (1 << 16)
        d_back = torch.empty(cap, dtype=torch.uint8, device=dev)
        nb = eng.decode_device(d_out.data_ptr(), n_ids, merges, d_back.data_ptr(), cap)
        barrier()
        t5 = time.perf_counter()
        nb = eng.decode_device(d_out.data_ptr(), n_ids, merges, d_back.data_ptr(), cap)
        barrier()
        td = torch.tensor([(time.perf_counter() - t5) * 1000, 0.0], dtype=torch.float64, device=dev)
        tb = torch.tensor([float(nb)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
            dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        rt_ok = bool(int(tb[0]) == total_bytes) and (world > 1 or bool(nb == n and torch.equal(d_back[:n], d_text)))
        dec_s = float(td[0]) / 1000.0
        dalg = 2 * ids_total + total_bytes
        dec = {"metric": "decode_output_GB_per_sec", "value": total_bytes / 1e9 / dec_s, "unit": "GB/s", "ms": dec_s * 1000,
               "round_trip": rt_ok, "round_trip_check": "bytes equal the corpus" if world == 1 else "decoded byte total equals the corpus size (the byte comparison runs at N = 1)",
               "roofline": {"bound": "hbm", "achieved": dalg / 1e9 / dec_s / world, "peak": peak_e, "unit": "GB/s", "frac": dalg / 1e9 / dec_s / world / peak_e,
                            "algorithmic_bytes": dalg, "note": "2*n_tok + n_bytes"}}
        del d_out, d_back

    # N > 1: rank 0 retrains and re-encodes the WHOLE corpus on its own GPU and compares (N-invariance in the record)
    if world > 1 and not args.no_single_check:
        if rank == 0:
            full = sc.generate(total_bytes, sc.SEED_C3, sc.BYTE)
            d_full = torch.from_numpy(full).to(dev)
            eng1 = zb.Engine(device=local_rank)
            if args.max_steps:
                eng1.set_option("max_steps", args.max_steps)
            m1, c1 = eng1.train(None, args.vocab, device_ptr=d_full.data_ptr(), n=total_bytes)
            parity["single_gpu_merges_sha256"] = hashlib.sha256(merges_text(m1)).hexdigest()
            parity["matches_single_gpu_train"] = bool(np.array_equal(m1, merges) and np.array_equal(c1, counts))
            if not args.no_encode:
                d_o1 = torch.empty(total_bytes, dtype=torch.int16, device=dev)
                k1 = eng1.encode_device(d_full.data_ptr(), total_bytes, m1, d_o1.data_ptr())
                parity["single_gpu_encode_ids_sha256"] = hashlib.sha256(d_o1[:k1].cpu().numpy().tobytes()).hexdigest()
                parity["matches_single_gpu_encode"] = parity["single_gpu_encode_ids_sha256"] == parity.get("encode_ids_sha256")
                del d_o1
            eng1.close()
            del d_full, full
        dist.barrier()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    alg_bytes = scanned * 2  # u16 slots, all launches
    samp_bytes = sampled_slots * 2  # the launches whose duration was measured (every 32nd merge step)
    traffic = None  # DRAM bytes per launch from the committed ncu --set full capture (ratio to algorithmic bytes)
    for tp in (os.path.join(ROOT, "profiles", "r02_merge_traffic.json"), os.path.join(ROOT, "profiles", "r01_merge_traffic.json")):
        if os.path.exists(tp) and merge_calls:
            traffic = json.load(open(tp))["traffic_over_algorithmic"] * samp_bytes / merge_calls
            break
    achieved = samp_bytes / 1e9 / (merge_ms / 1000.0) if merge_ms > 0 else None
    est_merge_ms_all = alg_bytes / 1e9 / achieved * 1000.0 if achieved else None  # all launches at the sampled rate
    line = {
        "metric": "train_merges_per_sec", "value": value, "unit": "merges/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": timed_ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u16", "data": "synthetic (synthcorpus-v1 byte variant, seed 0x5EED0003)",
        "config": {"workload": workload(total_bytes, args.vocab, world),
                   "bytes": total_bytes, "bytes_rank0": n, "vocab": args.vocab, "merges_per_step": int(len(merges)),
                   "l2": "the resident sequence (2 B x token slots) is larger than the 126 MB L2 for all but the last steps at N>=4; consecutive "
                         "passes alternate direction so the tail of one pass is reused from L2 by the next; no explicit flush",
                   "timing": "host clock between barrier+synchronize pairs, max over ranks; the library's CUDA-event device_ms is reported beside it"},
        "device_ms_per_step": dev_ms_max / args.steps,
        "library_call_ms_per_step": lib_ms / args.steps,  # host clock inside bpe_train_device (rank 0); the rest of ms_per_step is the Python binding + barriers
        "aeqb_steps": int(st.get("aeqb_steps", 0)),
        "gpu_launches": int(launches),
        "clocks": clk,
        "e2e": e2e,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                     "traffic": traffic, "kernel": "merge_loop_kernel (u16)", "peak_source": peak_src,
                     "bytes_per_launch": samp_bytes / max(merge_calls, 1), "avg_launch_ms": merge_ms / max(merge_calls, 1),
                     "launches_timed": int(merge_calls), "sampling": "every 32nd merge step, CUDA events on the library's stream",
                     "kernel_share_of_step": est_merge_ms_all / dev_ms if dev_ms and est_merge_ms_all else None,
                     "whole_step_scan_GBps": alg_bytes / 1e9 / (dev_ms / 1000.0) if dev_ms else None,
                     "frac_of_nominal_8TBps": (achieved / 8000.0) if achieved else None},
        "tie_steps": int(st["tie_steps"]), "tie_slow_steps": int(st["tie_slow_steps"]), "compactions": int(st["compactions"]),
        "gpu_same_k_steps": same_k,
        "parity": parity,
        "encode": enc,
        "decode": dec,
    }
    if not args.no_cpu_baseline and world == 1:
        from oracle import oracle_py as ora
        old = pin_to_one_core()
        t2 = time.perf_counter()
        om, oc = ora.train(host, args.vocab, max_steps=args.cpu_steps, fast=False)
        cpu_s = time.perf_counter() - t2
        unpin(old)
        k = len(om)
        ok = bool(np.array_equal(np.stack([merges["first"], merges["second"], merges["new_token"]], axis=1)[:k], om)
                  and np.array_equal(counts[:k], oc))
        line["cpu_baseline"] = {"value": k / cpu_s, "unit": "merges/s", "cores": 1, "kind": "port",
                                "sample": f"first {k} merge steps (verbatim per-step recount) of the same {n}-byte corpus, {cpu_s:.1f} s, 1 thread pinned to one core; "
                                          f"the GPU's time for the same {k} steps is gpu_same_k_steps",
                                "host_cores_available": os.cpu_count(), "matches_gpu_merges": ok,
                                "like_for_like_speedup_same_k_steps": (cpu_s * 1000 / same_k["wall_ms"]) if same_k else None}
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
