#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_models.py tests/test_igev_stereo_gpu.py tests/test_update_umma_gpu.py tests/test_raft_train_gpu.py tests/test_igev_train_gpu.py -m gpu -q 2>&1 | tail -4
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/t30_bench.json 2> gpurun_out/t30_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/t30_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t30_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','epe_hot_path_px','epe_e2e_px','gpu_launches','cpu_baseline','raft_stereo','sceneflow','fast_fp16','reference_gpu_eager'):
    print(k, json.dumps(d.get(k))[:600])
t=d.get('train_step',{})
print('train_step', {k:t.get(k) for k in ('value','ms_per_step','forward_backward_ms','error')})
PY
