#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_igev_stereo_gpu.py tests/test_update_umma_gpu.py tests/test_raft_train_gpu.py tests/test_igev_train_gpu.py tests/test_gpu_blocks.py -m gpu -q 2>&1 | tail -4
echo "--- raft"; timeout 400 python tools/model_bench.py --model raft --height 512 --width 1024 --iters 32 --cuda-graph 2>&1 | tail -1 | cut -c1-330
for P in fp16 fp16x2; do echo "--- igev $P"; timeout 400 python tools/model_bench.py --model igev --height 1152 --width 1920 --maxdisp 256 --iters 32 --precision $P 2>&1 | tail -1 | cut -c1-330; done
