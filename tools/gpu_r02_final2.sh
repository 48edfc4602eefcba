#!/bin/sh
# tools/gpu_r02_final2.sh — one gpurun call: GPU parity tests, bench (both arms), smoke, then the ncu launch list and a full
# capture of the step kernels of the same short bench command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_n1_final.json; tail -2 gpurun_out/bench_n1_final.err
timeout 120 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; echo "reference arm rc=$?"; cut -c1-200 gpurun_out/bench_reference_arm.json
python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
B="python bench.py --steps 1 --warmup 0 --max-steps 1600 --no-e2e --no-encode --no-cpu-baseline"
timeout 100 $B > gpurun_out/plain_short.log 2>&1 &&
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 400 --csv --log-file gpurun_out/r02_launches_final.csv $B > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'apply_kernel|merge_loop_kernel' -s 3000 -c 4 -o gpurun_out/r02_step_kernels_final $B > gpurun_out/ncu_steps.log 2>&1
echo "ncu steps rc=$?"; tail -2 gpurun_out/ncu_steps.log
