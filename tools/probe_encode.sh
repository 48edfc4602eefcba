#!/bin/sh
# tools/probe_encode.sh — one gpurun call that measures the encode variants and captures the level pass:
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'sh tools/probe_encode.sh'
# 1. plain timing of the variants (level schedule with/without the successor filter, CTA-per-tile, one pass per merge)
# 2. only after that exited 0: the launch list of one encode and one full capture of three level_kernel launches
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BYTES=${BYTES:-1e9}
VOCAB=${VOCAB:-8192}
timeout 200 python tools/encode_gpu.py "$BYTES" "$VOCAB" 0:6:0,0:6:1,0:0:0,0:0:1,1 > gpurun_out/encode_variants.json 2> gpurun_out/encode_variants.err
cat gpurun_out/encode_variants.json
# launch list: the training launches come first (4 per step), so count from the end instead of skipping
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/encode_launches.csv \
    -k regex:'level_kernel|halo_kernel|compact|tile_' python tools/encode_gpu.py 2e8 4096 0:6:0 > gpurun_out/ncu_launches.log 2>&1 || true
timeout 300 ncu --set full --clock-control none --import-source on -k regex:level_kernel -s 2 -c 3 -o gpurun_out/level_kernel \
    python tools/encode_gpu.py 2e8 4096 0:6:0 > gpurun_out/ncu_full.log 2>&1 || true
ls -la gpurun_out | tail -8
