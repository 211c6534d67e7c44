#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_update_umma_gpu.py -m gpu -q -x -s 2>&1 | grep -vE "^$|warn" | tail -30
for extra in "--update torch" "--update torch --exact-glue" "--update umma" "--update umma --cuda-graph" "--update torch --cuda-graph"; do
  echo "--- raft 512x1024 x32 $extra"
  timeout 600 python tools/model_bench.py --model raft --height 512 --width 1024 --iters 32 $extra 2>&1 | tail -1 | cut -c1-400
done
