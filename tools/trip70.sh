#!/bin/bash
# tile-height sweep of the fp16x2 3-D layers (host-side tiling knobs only; results are bit-identical across tilings)
mkdir -p gpurun_out
STB_UMMA_VERBOSE=1 timeout 300 python tools/layer_bench.py --precision fp16x2 --reps 1 2> gpurun_out/t70_verbose.err > /dev/null
for th in 0 4 8 12 16; do
  STB_UMMA_TH=$th timeout 300 python tools/layer_bench.py --precision fp16x2 --reps 7 --json gpurun_out/t70_th$th.json > gpurun_out/t70_th$th.log 2>&1
  echo "TH=$th rc=$?"
done
python - <<'PY'
import json
tabs = {th: {r["layer"]: r for r in json.load(open(f"gpurun_out/t70_th{th}.json"))} for th in (0, 4, 8, 12, 16)}
names = list(tabs[0].keys())
print("%-18s" % "layer", *("TH=%-6d" % t for t in tabs))
for n in names:
    print("%-18s" % n, *("%-9.1f" % tabs[t][n]["us"] for t in tabs))
PY
grep -E "stb_conv3d_umma" gpurun_out/t70_verbose.err | sort | uniq | head -60
