#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/t71_bench.json 2> gpurun_out/t71_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t71_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','epe_hot_path_px','epe_e2e_px','fast_fp16','clocks'):
    print(k, json.dumps(d.get(k))[:300])
print("raft", d["raft_stereo"].get("ms_per_forward"), "train", d["train_step"].get("ms_per_step"), "sceneflow", d["sceneflow"].get("value"), "cfg5", json.dumps(d["config5"])[:200])
PY
