#!/bin/bash
STB_UMMA_VERBOSE=0 timeout 600 python -m pytest tests/test_gpu_blocks.py tests/test_igev_stereo_gpu.py -m gpu -q -x -s -k "igev" 2>&1 | grep -vE "^$|warn" | tail -25
for P in fp16 fp16x2; do echo "--- igev 1152x1920 $P"; timeout 400 python tools/model_bench.py --model igev --height 1152 --width 1920 --maxdisp 256 --iters 32 --precision $P --channels-last 2>&1 | tail -1 | cut -c1-260; done
