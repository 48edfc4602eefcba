#!/bin/sh
# tools/gpu_r02_enc.sh — one gpurun call for the tile-resident encoder: GPU parity tests that touch encode, timing of the
# variants on C3 (1 GB, 7,936 merges), then ncu: launch list + full capture of tilenc_kernel and the compaction kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "encode or tile or c2 or smoke or ref_ or decode" > gpurun_out/pytest_encode.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_encode.log
timeout 600 python tools/encode_gpu.py 1e9 8192 ${VARIANTS:-4,4:encode_tile=4096,4:encode_tile=6144,4:encode_tile=2048,2} > gpurun_out/tilenc_variants.json 2> gpurun_out/tilenc_variants.err || { tail -20 gpurun_out/tilenc_variants.err; exit 1; }
cat gpurun_out/tilenc_variants.json
E="python tools/encode_gpu.py 2e8 4096 4"
timeout 200 $E > gpurun_out/plain_enc.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tilenc|compact_scatter|tile_count|tile_scan' -s 12 -c 8 -o gpurun_out/r02_tilenc $E > gpurun_out/ncu_enc.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_enc.log
ls -la gpurun_out | tail -6
