#!/bin/sh
# tools/gpu_r02_final.sh — the round's last single-GPU numbers: bench.py as the driver runs it (fewer steps), both arms
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_n1_final.json; tail -2 gpurun_out/bench_n1_final.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; echo "reference arm rc=$?"; cut -c1-300 gpurun_out/bench_reference_arm.json
python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
