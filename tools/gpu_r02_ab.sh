#!/bin/sh
# tools/gpu_r02_ab.sh — one gpurun call: A/B of merge-pass options on the C3 training, then an ncu capture of the step
# kernels at two depths of the run (dense steps ~300, sparse steps ~1500)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/ab_gpu.py 1e9 8192 ${SETS:-merge_prestage=0,merge_prestage=1} > gpurun_out/ab.json 2> gpurun_out/ab.err; echo "ab rc=$?"; cat gpurun_out/ab.json; tail -3 gpurun_out/ab.err
B="python bench.py --steps 1 --warmup 0 --max-steps ${MAXSTEPS:-1600} --no-e2e --no-encode --no-cpu-baseline"
timeout 200 $B > gpurun_out/plain_short.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'apply_kernel|merge_kernel' -s ${SKIP:-3000} -c 4 -o gpurun_out/${REP:-r02_step_kernels_c} $B > gpurun_out/ncu_steps.log 2>&1
echo "ncu steps rc=$?"; tail -2 gpurun_out/ncu_steps.log
