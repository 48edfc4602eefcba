#!/bin/sh
# tests/emul/build.sh — builds the CPU emulation of the CUDA kernels (TEST INFRASTRUCTURE ONLY).
# Small tiles / blocks so that tile-boundary logic is exercised by tiny inputs.
set -e
cd "$(dirname "$0")"
g++ -O2 -g -std=c++17 -fPIC -shared -DBPE_EMUL -DBPE_TILE=${BPE_TILE:-256} -DBPE_THREADS=${BPE_THREADS:-32} \
    -I. -I../../zig-bpe_b200/csrc -x c++ ../../zig-bpe_b200/csrc/bpe_engine.cu cuda_emul.cpp \
    -o libbpe_emul.so
