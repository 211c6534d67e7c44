#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_split_gpu.py tests/test_gpu_umma.py -m gpu -q -x 2>&1 | tail -2
for rep in 1 2; do
echo "--- A (csplit) rep $rep"
timeout 600 python tools/layer_bench.py --precision fp16x2 --only "k3 s1" --reps 9 2>&1 | tail -6
echo "--- B (no csplit) rep $rep"
STB200_LIB=$PWD/stereo_toolbox_b200/libstb200_nocsplit.so timeout 600 python tools/layer_bench.py --precision fp16x2 --only "k3 s1" --reps 9 2>&1 | tail -6
done
echo "--- bench A"
timeout 900 python bench.py --steps 8 --warmup 3 --no-extras --no-train --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], {k:round(v['ms_total']/8,2) for k,v in d['kernels'].items()})"
echo "--- bench B"
STB200_LIB=$PWD/stereo_toolbox_b200/libstb200_nocsplit.so timeout 900 python bench.py --steps 8 --warmup 3 --no-extras --no-train --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], {k:round(v['ms_total']/8,2) for k,v in d['kernels'].items()})"
