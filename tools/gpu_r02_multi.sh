#!/bin/sh
# tools/gpu_r02_multi.sh — one `gpurun --gpus N` call (N = 2, 4, 8; default 2): multi-GPU parity against a single GPU (train +
# encode), the same with one rank forced to fail its peer-mailbox set-up (every rank must fall back to NCCL together),
# the per-kernel-class breakdown of a sharded training, and the bench line at N
N=${N:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29533 tools/dist_check.py 5e7 2048 > gpurun_out/dist_check_n$N.json 2> gpurun_out/dist_check_n$N.err; echo "dist_check rc=$?"; tail -1 gpurun_out/dist_check_n$N.json
[ -z "$SKIP_PEERFAIL" ] && BPE_TEST_PEER_FAIL_RANK=1 timeout 300 $TR --master-port 29534 tools/dist_check.py 2e7 1024 > gpurun_out/dist_check_peerfail_n$N.json 2> gpurun_out/dist_check_peerfail_n$N.err; echo "dist_check (peer set-up fails on rank 1) rc=$?"; tail -1 gpurun_out/dist_check_peerfail_n$N.json
timeout 300 $TR --master-port 29535 tools/quick_dist.py 1e9 8192 1 > gpurun_out/quick_dist_n$N.json 2> gpurun_out/quick_dist_n$N.err; echo "quick_dist rc=$?"; tail -2 gpurun_out/quick_dist_n$N.json
timeout 600 $TR --master-port 29536 bench.py --gpus $N --steps ${STEPS:-3} --warmup 2 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
if [ -n "$C5" ]; then
timeout 600 $TR --master-port 29541 tools/c5_run.py ${C5_BYTES:-1e10} 1e9 65535 > gpurun_out/c5_n$N.json 2> gpurun_out/c5_n$N.err; echo "c5 rc=$?"; tail -1 gpurun_out/c5_n$N.json | cut -c1-900
fi
