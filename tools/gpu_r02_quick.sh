#!/bin/sh
# tools/gpu_r02_quick.sh — one gpurun call: A/B of the merge kernel's candidate filter on the bench workload, then the
# encode variants with the tile encoder's per-window counters
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n1.json"))
print({k: d[k] for k in ("value", "ms_per_step", "device_ms_per_step", "library_call_ms_per_step", "aeqb_steps")}, d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["e2e"]["value"], d["encode"]["value"], d["decode"]["value"])
PY
BPE_OPTS=${AB_OPTS:-batch_steps=16} timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-encode --no-cpu-baseline > gpurun_out/bench_n1_ab.json 2> gpurun_out/bench_n1_ab.err; echo "bench (A/B options) rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n1_ab.json"))
print({k: d[k] for k in ("value", "ms_per_step", "device_ms_per_step", "library_call_ms_per_step")}, d["roofline"]["frac"], d["roofline"]["avg_launch_ms"])
PY
ENC_DEBUG=1 timeout 600 python tools/encode_gpu.py 1e9 8192 ${VARIANTS:-4,4:encode_tile=5632,2} > gpurun_out/tilenc_variants.json 2> gpurun_out/tilenc_variants.err; echo "variants rc=$?"; cat gpurun_out/tilenc_variants.json; grep "tile encode" gpurun_out/tilenc_variants.err | tail -4
