#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fullshape_parity_gpu.py -m gpu -q -s -k "igev_stereo_update" 2>&1 | grep -E "IGEV|passed|failed|Error"
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/t68_bench.json 2> gpurun_out/t68_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t68_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','epe_hot_path_px','epe_e2e_px','config5'):
    print(k, json.dumps(d.get(k))[:500])
print("raft", {k:d["raft_stereo"].get(k) for k in ("ms_per_forward","epe_vs_torch_fp32_px","error")}, "train", d["train_step"].get("ms_per_step"), "sceneflow", d["sceneflow"].get("value"))
PY
