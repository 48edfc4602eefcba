#!/bin/sh
# tools/gpu_r02_n2.sh — gpurun --gpus 2: sharded train + encode against one GPU, then the bench line at N = 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 100 $TR --master-port 29533 tools/dist_check.py 5e7 2048 > gpurun_out/dist_check_n2.json 2> gpurun_out/dist_check_n2.err; echo "dist_check rc=$?"; tail -1 gpurun_out/dist_check_n2.json
timeout 150 $TR --master-port 29536 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
