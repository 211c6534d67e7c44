#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_fullshape_parity_gpu.py -m gpu -q -s -rA --timeout 1500 > gpurun_out/t6_full.log 2>&1; echo "fullshape rc=$?"
grep -E "EPE|err|passed|failed|Error|PASSED|FAILED" gpurun_out/t6_full.log | cut -c1-250 | tail -30
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/t6_bench.json 2> gpurun_out/t6_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/t6_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/t6_bench.json'))
for k in ('value','ms_per_step','e2e','epe_e2e_px','epe_hot_path_px','gpu_launches','cpu_baseline','fast_fp16','reference_gpu_eager','sceneflow'):
    print(k, d.get(k))
PY
