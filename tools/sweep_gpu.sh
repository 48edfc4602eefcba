#!/bin/sh
# tools/sweep_gpu.sh — A/B of merge-kernel geometry variants built as libbpe_b200_<tag>.so
for f in zig-bpe_b200/lib/libbpe_b200*.so; do
  echo "== $f"
  export BPE_B200_LIB=$f
  timeout 60 python tools/stream_gpu.py 3e8 60 | tail -1 | cut -c1-110
  timeout 60 python tools/stream_gpu.py 3e8 60 122 | tail -1 | cut -c1-110
  timeout 60 python tools/stream_gpu.py 3e8 60 101 | tail -1 | cut -c1-110
  timeout 120 python tools/quick_gpu.py 1e9 8192 2 2>&1 | tail -2 | head -1 | cut -c1-160
done
