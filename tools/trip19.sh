#!/bin/bash
mkdir -p gpurun_out
echo "--- promo = chunk bytes (default)"
timeout 600 python tools/layer_bench.py --precision fp16x2 --json gpurun_out/t19_layers_fp16x2.json 2>&1 | tail -13
echo "--- STB_TMA_PROMO=256 (old)"
STB_TMA_PROMO=256 timeout 600 python tools/layer_bench.py --precision fp16x2 2>&1 | tail -13
echo "--- fp16 default"
timeout 600 python tools/layer_bench.py --precision fp16 2>&1 | tail -3
STB_CUDA_PROFILER=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:volume_cl2 --launch-count 1 --profile-from-start off \
  -f -o gpurun_out/r2_ncu_volume_cl2 python bench.py --steps 1 --warmup 2 --no-extras --no-train --no-cpu-baseline > gpurun_out/t19_ncu_vol.log 2>&1; echo "ncu vol rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-train --no-cpu-baseline > gpurun_out/t19_bench.json 2> gpurun_out/t19_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t19_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e'):
    print(k, d.get(k))
print(json.dumps(d.get('kernels',{}))[:1500])
PY
