#!/bin/sh
# tools/gpu_r02_c5.sh — one gpurun call on 1 GPU: all GPU parity tests, BASELINE config 5 at N = 1 (10 GB encode with the
# 65,279-merge list), then ncu captures of the level pass, the tile encoder and the decode kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
ENC_DEBUG=1 timeout 900 python tools/c5_run.py ${C5_BYTES:-1e10} 1e9 65535 > gpurun_out/c5_n1.json 2> gpurun_out/c5_n1.err; echo "c5 rc=$?"; cat gpurun_out/c5_n1.json; grep "tile encode" gpurun_out/c5_n1.err | tail -2; tail -2 gpurun_out/c5_n1.err
if [ -z "$NO_NCU" ]; then
E="python tools/encode_gpu.py 2e8 4096 2,4"
timeout 200 $E > gpurun_out/plain_enc.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'level_kernel|tilenc_kernel' -s 20 -c 5 -o gpurun_out/r02_level_tilenc $E > gpurun_out/ncu_enc.log 2>&1
echo "ncu enc rc=$?"; tail -2 gpurun_out/ncu_enc.log
fi
ls -la gpurun_out | tail -5
