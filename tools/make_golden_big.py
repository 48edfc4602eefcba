#!/usr/bin/env python
"""tools/make_golden_big.py — full-size golden merge lists from the CPU oracle (test infrastructure).

Runs the oracle's `fast` mode (same results as the verbatim mode, checked in tests/test_oracle.py) over a
whole BASELINE.json config and commits what the GPU parity tests compare against:

  tests/golden/<cfg>_merges.txt   the learned merges in the reference's merges.txt format (:328)
  tests/golden/<cfg>_counts.txt   count of the winning pair per merge (the verbose line, :309)
  tests/golden/big_sha256.json    sha256 of the corpus, of the merges file and of its prefixes

  python tools/make_golden_big.py c2            # 100 MB UTF-8 corpus, vocab 4096: ~15 min on one core
  python tools/make_golden_big.py c3            # 1 GB byte corpus, vocab 8192: a few hours on one core
  python tools/make_golden_big.py c3 --collect  # only collect what a (still running / stopped) job has printed

The oracle CLI prints every merge as it is learned (verbose line on stderr), so a stopped C3 job still
leaves a usable prefix: the JSON then records `complete: false` and the prefix hashes that exist.
"""
import argparse
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synthcorpus as sc  # noqa: E402

CFG = {
    "c2": dict(bytes=100_000_000, seed=sc.SEED_C2, variant=sc.UTF8, vocab=4096),
    "c3": dict(bytes=1_000_000_000, seed=sc.SEED_C3, variant=sc.BYTE, vocab=8192),
}
PREFIXES = (1, 10, 40, 100, 250, 500, 1000, 2000, 3000, 3840, 4000, 4971, 5000, 5500, 6000, 6500, 7000, 7500, 7936)
LINE = re.compile(r"merge (\d+)/(\d+): \((\d+),(\d+)\) -> (\d+) had (\d+) occurrences")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cfg", choices=sorted(CFG))
    ap.add_argument("--workdir", default="/tmp/gold")
    ap.add_argument("--collect", action="store_true", help="do not run the oracle, parse <workdir>/<cfg>.log")
    ap.add_argument("--gpu-verified", type=int, default=None,
                    help="record that a B200 training was compared with the first N merges of this list (tools/check_c3_golden.py); "
                         "the GPU test asserts that many, bench.py reports the comparison with every prefix hash")
    a = ap.parse_args()
    c = CFG[a.cfg]
    os.makedirs(a.workdir, exist_ok=True)
    corpus = os.path.join(a.workdir, a.cfg + ".bin")
    log = os.path.join(a.workdir, a.cfg + ".log")
    if not os.path.exists(corpus):
        sc.generate(c["bytes"], c["seed"], c["variant"]).tofile(corpus)
    if not a.collect:
        exe = os.path.join(ROOT, "oracle", "bpe_oracle")
        with open(log, "w") as lf:
            subprocess.run([exe, "train", corpus, str(c["vocab"]), os.path.join(a.workdir, a.cfg + "_merges.txt"), "-1", "1", "1"],
                           stderr=lf, check=True)
    merges, counts = [], []
    for line in open(log):
        m = LINE.match(line)
        if m:
            assert int(m[1]) == len(merges) + 1
            merges.append((int(m[3]), int(m[4]), int(m[5])))
            counts.append(int(m[6]))
    want = c["vocab"] - 256
    text = "".join(f"{x},{y},{z}\n" for x, y, z in merges)
    gold = os.path.join(ROOT, "tests", "golden")
    open(os.path.join(gold, a.cfg + "_merges.txt"), "w").write(text)
    open(os.path.join(gold, a.cfg + "_counts.txt"), "w").write("".join(f"{v}\n" for v in counts))
    h = hashlib.sha256()
    with open(corpus, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
    lines = text.splitlines(keepends=True)
    rec = {"bytes": c["bytes"], "seed": hex(c["seed"]), "variant": c["variant"], "vocab": c["vocab"], "corpus_sha256": h.hexdigest(),
           "merges": len(merges), "complete": len(merges) == want, "merges_sha256": hashlib.sha256(text.encode()).hexdigest(),
           "prefix_sha256": {str(k): hashlib.sha256("".join(lines[:k]).encode()).hexdigest() for k in PREFIXES if k <= len(merges)},
           "how": "oracle/bpe_oracle train <corpus> <vocab> <out> -1 1 1 (fast mode, verbose), tools/make_golden_big.py"}
    p = os.path.join(gold, "big_sha256.json")
    allrec = json.load(open(p)) if os.path.exists(p) else {}
    gv = a.gpu_verified if a.gpu_verified is not None else allrec.get(a.cfg, {}).get("gpu_verified_merges")
    if gv is not None:
        rec["gpu_verified_merges"] = min(int(gv), len(merges))
    allrec[a.cfg] = rec
    json.dump(allrec, open(p, "w"), indent=1, sort_keys=True)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
