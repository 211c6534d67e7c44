#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/t15_bench.json 2> gpurun_out/t15_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/t15_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/t15_bench.json'))
for k in ('value','ms_per_step','e2e','epe_e2e_px','epe_hot_path_px','gpu_launches','clocks'):
    print(k, d.get(k))
for k in ('fast_fp16','reference_gpu_eager','sceneflow','train_step'):
    print(k, str(d.get(k))[:330])
print({k:d['roofline'][k] for k in ('kernel','achieved','frac','frac_issued')})
for k,v in d['kernels'].items(): print(k, v)
PY
