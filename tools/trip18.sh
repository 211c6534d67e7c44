#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_split_gpu.py -m gpu -q -x -k "conv_family" > gpurun_out/t18_split.log 2>&1; echo "split tests rc=$?"; tail -5 gpurun_out/t18_split.log
STB_UMMA_VERBOSE=1 timeout 300 python tools/layer_bench.py --precision fp16x2 --only "64->32 k3 s2T" --reps 3 2>&1 | grep -v "^$" | tail -6
timeout 600 python tools/layer_bench.py --precision fp16x2 --json gpurun_out/t18_layers_fp16x2.json 2>&1 | tail -16
echo "--- KDEPTH3D=0"
STB_UMMA_KDEPTH3D=0 timeout 300 python tools/layer_bench.py --precision fp16x2 --only "64->" 2>&1 | tail -8
timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-train --no-cpu-baseline > gpurun_out/t18_bench.json 2> gpurun_out/t18_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t18_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','epe_hot_path_px','epe_e2e_px','gpu_launches','roofline'):
    print(k, d.get(k))
print({k:(v.get('ms'),v.get('launches')) for k,v in d.get('kernels',{}).items()})
PY
