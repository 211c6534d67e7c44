#!/bin/bash
timeout 600 python tools/train_step.py --precision bf16 --features tf32 --height 576 --width 960 --batch 1 --steps 5 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['gpu_launches'], d['loss'])"
