#!/bin/sh
# tools/gpu_r02_all.sh — one gpurun call: all GPU parity tests, the bench line, the encode variants, then ncu captures
# (step kernels in steady state, the tile encoder and the compaction kernels)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 600 python tools/encode_gpu.py 1e9 8192 ${VARIANTS:-4,4:encode_tile=4096,4:encode_tile=6144,4:encode_tile=2048,2} > gpurun_out/tilenc_variants.json 2> gpurun_out/tilenc_variants.err; echo "variants rc=$?"; cat gpurun_out/tilenc_variants.json; tail -3 gpurun_out/tilenc_variants.err
if [ -z "$NO_NCU" ]; then
B="python bench.py --steps 1 --warmup 0 --max-steps 2100 --no-e2e --no-encode --no-cpu-baseline"
timeout 200 $B > gpurun_out/plain_short.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'apply_kernel|merge_kernel' -s 3000 -c 6 -o gpurun_out/r02_step_kernels_b $B > gpurun_out/ncu_steps.log 2>&1
echo "ncu steps rc=$?"
E="python tools/encode_gpu.py 2e8 4096 4"
timeout 200 $E > gpurun_out/plain_enc.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tilenc|compact_scatter|tile_count|tile_scan' -s 12 -c 8 -o gpurun_out/r02_tilenc $E > gpurun_out/ncu_enc.log 2>&1
echo "ncu enc rc=$?"; tail -2 gpurun_out/ncu_enc.log
fi
ls -la gpurun_out | tail -8
