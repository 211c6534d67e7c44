#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/layer_bench.py --precision fp16x2 --reps 5 > gpurun_out/t13_layers_x2.log 2>&1; grep -v "^\[stb" gpurun_out/t13_layers_x2.log
timeout 600 python tools/layer_bench.py --precision fp16 --reps 5 > gpurun_out/t13_layers_f16.log 2>&1; grep -v "^\[stb" gpurun_out/t13_layers_f16.log
timeout 900 python -m pytest tests/test_split_gpu.py tests/test_gpu_umma.py -m gpu -q --timeout 600 > gpurun_out/t13_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t13_tests.log | cut -c1-200
