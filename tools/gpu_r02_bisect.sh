#!/bin/sh
# tools/gpu_r02_bisect.sh — N-GPU dist_check with library variants (BPE_B200_LIB) to bisect the exchange
N=${N:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29533
for v in default; do
  port=$((port+1))
  BPE_OPTS=${BPE_OPTS:-pdl=1} timeout 120 $TR --master-port $port tools/dist_check.py 5e7 2048 > gpurun_out/bisect_$v.json 2> gpurun_out/bisect_$v.err; echo "$v rc=$?"; tail -1 gpurun_out/bisect_$v.json | cut -c1-200
done
