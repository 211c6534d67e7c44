#!/bin/bash
timeout 300 python -m pytest tests/test_refine2d_gpu.py -m gpu -q 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q -k "pcwnet or cfnet" 2>&1 | tail -3
timeout 300 python tools/model_bench.py --model pcwnet_gc --height 384 --width 1248 --precision fp32 2>&1 | tail -1 | cut -c1-250
