#!/bin/bash
mkdir -p gpurun_out
timeout 300 bash stereo_toolbox_b200/csrc/probe/run_tmabw2.sh > gpurun_out/t4_tmabw2.txt 2>&1; cat gpurun_out/t4_tmabw2.txt
STB_UMMA_VERBOSE=1 timeout 600 python tools/layer_bench.py --precision fp16x2 --reps 5 > gpurun_out/t4_layers_x2.log 2>&1; cat gpurun_out/t4_layers_x2.log | cut -c1-330
STB_UMMA_VERBOSE=1 timeout 600 python tools/layer_bench.py --precision fp16 --reps 5 > gpurun_out/t4_layers_f16.log 2>&1; grep -v "^\[stb" gpurun_out/t4_layers_f16.log
