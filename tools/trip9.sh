#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py tests/test_igev_stereo_gpu.py -m gpu -q --timeout 500 > gpurun_out/t9_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t9_tests.log | cut -c1-300
{
  timeout 300 python tools/model_bench.py --model raft   --height 512  --width 1024 --iters 32
  timeout 300 python tools/model_bench.py --model raft   --height 512  --width 1024 --iters 32 --cuda-graph
  timeout 300 python tools/model_bench.py --model raft   --height 512  --width 1024 --iters 32 --cuda-graph --channels-last
  timeout 300 python tools/model_bench.py --model acvnet --height 1152 --width 1920 --maxdisp 256 --precision fp16x2
  timeout 300 python tools/model_bench.py --model acvnet --height 1152 --width 1920 --maxdisp 256 --precision fp16
  timeout 300 python tools/model_bench.py --model igev   --height 1152 --width 1920 --maxdisp 256 --iters 32 --precision fp32
  timeout 300 python tools/model_bench.py --model igev   --height 1152 --width 1920 --maxdisp 256 --iters 32 --precision fp16
  timeout 300 python tools/model_bench.py --model psmnet --height 576  --width 960  --batch 4 --precision fp16x2
  timeout 300 python tools/model_bench.py --model cfnet --height 384 --width 1248 --precision fp32
  timeout 300 python tools/model_bench.py --model pcwnet_gc --height 384 --width 1248 --precision fp32
} > gpurun_out/t9_models.jsonl 2> gpurun_out/t9_models.err
cut -c1-330 gpurun_out/t9_models.jsonl; tail -5 gpurun_out/t9_models.err
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/t9_bench.json 2> gpurun_out/t9_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/t9_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/t9_bench.json'))
for k in ('value','ms_per_step','e2e','epe_e2e_px','epe_hot_path_px','train_step','reference_gpu_eager','sceneflow'):
    print(k, str(d.get(k))[:600])
PY
