#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train16_gpu.py tests/test_split_gpu.py -m gpu -q -s -k "training_step or psmnet_golden" 2>&1 | grep -E "cos|PSMNet|passed|failed|Error|error|assert" | head -70
for f in tf32 amp; do
echo "--- train_step bf16 576x960 features=$f"
timeout 600 python tools/train_step.py --precision bf16 --features $f --height 576 --width 960 --batch 1 --steps 3 --warmup 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
done
for fm in "" "--features fp32"; do
echo "--- psmnet fwd 576x960 B=4 fp16x2 $fm"
timeout 600 python tools/model_bench.py --model psmnet --height 576 --width 960 --batch 4 --precision fp16x2 $fm 2>&1 | tail -1 | cut -c1-300
done
