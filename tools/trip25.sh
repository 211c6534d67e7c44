#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_igev_stereo_gpu.py -m gpu -q -x -s -k "tensor_cores" 2>&1 | grep -vE "^$|warn" | tail -12
for extra in "--update umma --cuda-graph --channels-last" "--update umma --channels-last"; do
  echo "--- raft 512x1024 x32 $extra"
  timeout 600 python tools/model_bench.py --model raft --height 512 --width 1024 --iters 32 $extra 2>&1 | tail -1 | cut -c1-330
done
for extra in "--cuda-graph" "--update umma" "--update umma --cuda-graph" "--update umma --cuda-graph --precision fp16"; do
  echo "--- igev 1152x1920 D=256 x32 $extra"
  timeout 600 python tools/model_bench.py --model igev --height 1152 --width 1920 --maxdisp 256 --iters 32 $extra 2>&1 | tail -1 | cut -c1-330
done
