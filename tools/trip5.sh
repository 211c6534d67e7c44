#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_split_gpu.py tests/test_gpu_umma.py -m gpu -q -s --timeout 600 > gpurun_out/t5_tests.log 2>&1; echo "tests rc=$?"
grep -E "GwcNet|PSMNet|passed|failed|Error" gpurun_out/t5_tests.log | tail -12
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/t5_bench.json 2> gpurun_out/t5_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/t5_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/t5_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','epe_e2e_px','epe_hot_path_px','gpu_launches','cpu_baseline','fast_fp16','reference_gpu_eager','sceneflow') if k in d})
print(d['roofline'])
for k,v in d['kernels'].items(): print(k, v)
for k,v in d['layers'].items(): print(k, v)
PY
