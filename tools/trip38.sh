#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_split_gpu.py tests/test_gpu_umma.py -m gpu -q -x 2>&1 | tail -3
echo "--- pair merge on (default)"
timeout 600 python tools/layer_bench.py --precision fp16x2 --only "s2" 2>&1 | tail -6
echo "--- STB_UMMA_PAIRMERGE=0"
STB_UMMA_PAIRMERGE=0 timeout 600 python tools/layer_bench.py --precision fp16x2 --only "k3 s2" 2>&1 | tail -6
echo "--- fp16 on / off"
timeout 600 python tools/layer_bench.py --precision fp16 --only "k3 s2" 2>&1 | tail -5
STB_UMMA_PAIRMERGE=0 timeout 600 python tools/layer_bench.py --precision fp16 --only "k3 s2" 2>&1 | tail -5
