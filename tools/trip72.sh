#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_split_gpu.py -m gpu -q -s -k "split_conv_family" 2>&1 | grep -E "fp16x2\] (128|192)|passed|failed|Error|error" | head -30
for g in 1 0; do
  STB_UMMA_KGROUP=$g STB_UMMA_VERBOSE=1 timeout 300 python tools/layer_bench.py --precision fp16x2 --reps 7 --only "128->64" 2> gpurun_out/t72_v$g.err | head -3
  grep -E "stb_conv3d_umma" gpurun_out/t72_v$g.err | sort | uniq | head -6
done
