#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_split_gpu.py -m gpu -q -x -k "conv_family" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_umma.py -m gpu -q -x 2>&1 | tail -3
echo "--- cluster on (default)"
timeout 600 python tools/layer_bench.py --precision fp16x2 --json gpurun_out/t36_layers_cluster.json 2>&1 | tail -13
echo "--- STB_UMMA_CLUSTER=2 (also 8-CTA clusters)"
STB_UMMA_CLUSTER=2 timeout 600 python tools/layer_bench.py --precision fp16x2 --only "128" 2>&1 | tail -5
echo "--- fp16 cluster on"
timeout 600 python tools/layer_bench.py --precision fp16 2>&1 | tail -13
echo "--- fp16 STB_UMMA_CLUSTER=0"
STB_UMMA_CLUSTER=0 timeout 600 python tools/layer_bench.py --precision fp16 2>&1 | tail -13
