#!/bin/sh
# tools/probe_segenc.sh — one gpurun call for the segment-resident encoder:
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'sh tools/probe_segenc.sh'
# 1. GPU parity tests that touch encode; 2. timing of the window geometries against the level schedule on C3;
# 3. only after that exited 0: launch list + one full ncu capture of segenc_kernel on a 200 MB run
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "encode or c2 or smoke or ref_" > gpurun_out/pytest_encode.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_encode.log
BYTES=${BYTES:-1e9}
VOCAB=${VOCAB:-8192}
timeout 300 python tools/encode_gpu.py "$BYTES" "$VOCAB" 0:6:0:0,0:6:0:1,0:6:0:2,2:6 > gpurun_out/segenc_variants.json 2> gpurun_out/segenc_variants.err || { tail -20 gpurun_out/segenc_variants.err; exit 1; }
cat gpurun_out/segenc_variants.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/segenc_launches.csv \
    -k regex:'segenc|tile_|compact|fill_holes' python tools/encode_gpu.py 2e8 4096 0:6:0:0 > gpurun_out/ncu_launches.log 2>&1 || true
timeout 400 ncu --set full --clock-control none --import-source on -k regex:segenc -c 1 -o gpurun_out/segenc_kernel \
    python tools/encode_gpu.py 2e8 4096 0:6:0:0 > gpurun_out/ncu_full.log 2>&1 || true
ls -la gpurun_out | tail -8
