#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/t73_bench.json 2> gpurun_out/t73_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t73_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','epe_hot_path_px','epe_e2e_px','clocks'):
    print(k, json.dumps(d.get(k))[:300])
print("fast", d["fast_fp16"].get("value"), "raft", d["raft_stereo"].get("ms_per_forward"), "train", d["train_step"].get("ms_per_step"), "sceneflow", d["sceneflow"].get("value"), "cfg5", json.dumps(d["config5"])[:260])
print(json.dumps(d["layers"].get("128->64 k3 s2T @12x24x78")))
PY
