#!/bin/bash
# Gray order of the class blocks of the merged transposed convs: parity first, then A/B per layer, then the full suite + smoke +
# bench under the setting that wins (one call: the round's GPU budget is nearly spent)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_split_gpu.py -m gpu -q -x -k "split_conv_family" 2>&1 | tail -3 > gpurun_out/t77_split.log
cat gpurun_out/t77_split.log
if ! grep -q " passed" gpurun_out/t77_split.log || grep -q "failed" gpurun_out/t77_split.log; then echo "GRAY PARITY FAILED"; export STB_UMMA_GRAYCLS=0; fi
for g in 1 0; do
  STB_UMMA_GRAYCLS=$g timeout 300 python tools/layer_bench.py --precision fp16x2 --reps 9 --only "s2T" --json gpurun_out/t77_g$g.json > /dev/null 2>&1
done
python - <<'PY'
import json
a = {r["layer"]: r["us"] for r in json.load(open("gpurun_out/t77_g1.json"))}
b = {r["layer"]: r["us"] for r in json.load(open("gpurun_out/t77_g0.json"))}
for k in a: print(k, "gray", a[k], "binary", b[k])
open("gpurun_out/t77_choice", "w").write("1" if sum(a.values()) < 0.98 * sum(b.values()) else "0")
PY
if [ "$STB_UMMA_GRAYCLS" != "0" ]; then export STB_UMMA_GRAYCLS=$(cat gpurun_out/t77_choice); fi
echo "running the suite with STB_UMMA_GRAYCLS=$STB_UMMA_GRAYCLS"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/t77_bench.json 2> gpurun_out/t77_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t77_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','epe_hot_path_px','epe_e2e_px','clocks'):
    print(k, json.dumps(d.get(k))[:200])
for k in ("128->64 k3 s2T @12x24x78", "64->32 k3 s2T @24x48x156", "32->32 k3 s1 @48x96x312"): print(k, json.dumps(d["layers"].get(k)))
PY
