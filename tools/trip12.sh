#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_split_gpu.py tests/test_gpu_umma.py -m gpu -q --timeout 600 > gpurun_out/t12_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t12_tests.log | cut -c1-200
timeout 900 python bench.py --steps 5 --warmup 3 --no-train --no-extras > gpurun_out/t12_bench.json 2> gpurun_out/t12_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/t12_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/t12_bench.json'))
for k in ('value','ms_per_step','epe_e2e_px','epe_hot_path_px'):
    print(k, d.get(k))
for k,v in d['kernels'].items(): print(k, v)
for k,v in d['layers'].items():
    if k.startswith('2d'): print(k, v)
PY
