#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train16_gpu.py -m gpu -q -x -s -k "training_step or raw_conv" 2>&1 | grep -E "cos|passed|failed|Error|error|assert" | head -60
echo "--- train_step bf16 576x960 amp features"
timeout 600 python tools/train_step.py --precision bf16 --features amp --height 576 --width 960 --batch 1 --steps 3 --warmup 2 2>&1 | tail -1 | cut -c1-700
echo "--- fp32 features"
timeout 600 python tools/train_step.py --precision bf16 --height 576 --width 960 --batch 1 --steps 3 --warmup 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
timeout 500 python tools/train_profile.py 2>&1 | tail -42
