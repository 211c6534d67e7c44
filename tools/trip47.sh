#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train16_gpu.py -m gpu -q -k "wgrad or raw_conv" 2>&1 | tail -2
bash tools/trip31.sh
timeout 500 python tools/train_profile.py 2>&1 | grep -E "GPU time|wgrad"
for L in "32->32 k3 s1" "64->32 k3 s2T"; do
  tag=$(echo "fp16x2_$L" | tr -c 'A-Za-z0-9' '_')
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3d_umma_kernel --launch-skip 1 --launch-count 1 \
    -f -o gpurun_out/r2b_ncu_$tag python tools/layer_bench.py --precision fp16x2 --only "$L" --reps 1 > gpurun_out/r2b_ncu_$tag.log 2>&1
  echo "== $L rc=$?"
  python tools/ncu_summary.py gpurun_out/r2b_ncu_$tag.ncu-rep --md 2>/dev/null | tail -1
done
