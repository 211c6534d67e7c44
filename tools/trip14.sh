#!/bin/bash
mkdir -p gpurun_out
for P in fp16x2 fp16; do
for L in "64->32 k3 s2T" "32->64 k3 s2"; do
  tag=$(echo "${P}_$L" | tr -c 'A-Za-z0-9' '_')
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3d_umma_kernel --launch-skip 3 --launch-count 2 \
    -f -o gpurun_out/r2_ncu_$tag python tools/layer_bench.py --precision $P --only "$L" --reps 1 > gpurun_out/r2_ncu_$tag.log 2>&1
  echo "== $P $L rc=$?"
  python tools/ncu_summary.py gpurun_out/r2_ncu_$tag.ncu-rep --md 2>/dev/null | tail -2
done; done
ls -la gpurun_out/*.ncu-rep | tail -5
