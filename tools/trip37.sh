#!/bin/bash
STB_UMMA_VERBOSE=1 timeout 120 python -m pytest tests/test_split_gpu.py -m gpu -q -x -s -k "conv_family and 32-64-3-1" 2>&1 | grep -E "smem base|stb_conv3d|passed|failed|Error|error" | head -20
