#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_split_gpu.py -m gpu -q -s -rA --timeout 600 > gpurun_out/t3_split.log 2>&1; echo "split rc=$?"
grep -E "GwcNet|PSMNet|passed|failed|Error|error" gpurun_out/t3_split.log | tail -20
timeout 900 python bench.py --precision fp16x2 --steps 5 --warmup 3 > gpurun_out/t3_bench_x2.json 2> gpurun_out/t3_bench_x2.err; echo "bench rc=$?"
tail -5 gpurun_out/t3_bench_x2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/t3_bench_x2.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','epe_e2e_px','epe_hot_path_px','gpu_launches') if k in d})
print(json.dumps(d['kernels'],indent=0))
for k,v in d['layers'].items(): print(k, v)
PY
