#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train16_gpu.py -m gpu -q -x -s 2>&1 | grep -E "wgrad|passed|failed|Error|error" | head -40
echo "--- train_step bf16 576x960 (tensor-core wgrad)"
timeout 600 python tools/train_step.py --precision bf16 --height 576 --width 960 --batch 1 --steps 3 --warmup 2 2>&1 | tail -2 | cut -c1-600
echo "--- train_step bf16 576x960 (STB_WGRAD_TC=0)"
STB_WGRAD_TC=0 timeout 600 python tools/train_step.py --precision bf16 --height 576 --width 960 --batch 1 --steps 3 --warmup 2 2>&1 | tail -2 | cut -c1-600
