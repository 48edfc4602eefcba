cd "$(dirname "$0")/.." 2>/dev/null
mkdir -p gpurun_out
timeout 300 python tools/encode_gpu.py 1e9 8192 0:6:0:0,0:6:0:4,0:6:0:5,0:6:0:2,2:6:0,2:6:1 > gpurun_out/segenc_variants2.json 2> gpurun_out/segenc_variants2.err || { tail -20 gpurun_out/segenc_variants2.err; exit 1; }
cat gpurun_out/segenc_variants2.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:segenc -c 1 -o gpurun_out/segenc_kernel2 \
    python tools/encode_gpu.py 2e8 4096 0:6:0:0 > gpurun_out/ncu_full2.log 2>&1 || true
