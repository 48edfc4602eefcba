#!/bin/sh
# tools/gpu_r02_base.sh — one gpurun call: GPU parity tests, the bench line, the encode variants and ncu captures of the
# step kernels besides merge_kernel (apply / halo / level / segment encoder / compaction):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'sh tools/gpu_r02_base.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_n1.json
timeout 300 python tools/encode_gpu.py 1e9 8192 3:6:1:4,2:6:1,2:6:0 > gpurun_out/encode_variants.json 2> gpurun_out/encode_variants.err; echo "encode rc=$?"; cat gpurun_out/encode_variants.json
B="python bench.py --steps 1 --warmup 0 --max-steps 700 --no-e2e --no-encode --no-cpu-baseline"
timeout 200 $B > gpurun_out/plain_short.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'apply_kernel|halo_kernel|merge_kernel' -s 1500 -c 9 -o gpurun_out/r02_step_kernels $B > gpurun_out/ncu_steps.log 2>&1
echo "ncu steps rc=$?"
E="python tools/encode_gpu.py 2e8 4096 3:6:1:4,2:6:1"
timeout 200 $E > gpurun_out/plain_enc.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'segenc|level_kernel|compact_scatter|tile_count' -c 8 -o gpurun_out/r02_encode_kernels $E > gpurun_out/ncu_enc.log 2>&1
echo "ncu enc rc=$?"
ls -la gpurun_out | tail -12
