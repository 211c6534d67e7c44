#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_split_gpu.py tests/test_gpu_umma.py tests/test_update_umma_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python tools/layer_bench.py --precision fp16x2 2>&1 | tail -13
timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-train --no-cpu-baseline > gpurun_out/t41_bench.json 2> gpurun_out/t41_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t41_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e'):
    print(k, d.get(k))
print({k:(round(v['ms_total']/5,2), v['launches']//5) for k,v in d.get('kernels',{}).items()})
for n,v in d['layers'].items():
    if n.startswith('2d'): print(n, v)
PY
timeout 300 python tools/layer_bench.py --precision fp16 2>&1 | tail -13
