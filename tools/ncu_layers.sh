#!/bin/bash
# One `ncu --set full` capture per slow conv flavour (profiles/next_round_plan.md section 3), through tools/layer_bench.py.
# Launch order inside one layer_bench run of ONE layer: #0 the kernel instance's one-off smem-base query (grid 1),
# #1 the warm-up conv, #2.. the timed convs  ->  skip 2, capture 2 (a K-split layer launches one pass per K-chunk).   Run under gpurun (1 GPU); reports land in gpurun_out/.
mkdir -p gpurun_out
for L in "64->32 k3 s2T" "128->64 k3 s2T" "32->64 k3 s2" "64->128 k3 s2" "32->32 k3 s1" "32->1 k3 s1"; do
  tag=$(echo "$L" | tr -c 'A-Za-z0-9' '_')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3d_umma_kernel --launch-skip 2 --launch-count 2 \
    -f -o gpurun_out/ncu_layer_$tag python tools/layer_bench.py --only "$L" --reps 1 > gpurun_out/ncu_layer_$tag.log 2>&1
  echo "== $L rc=$?"
  python tools/ncu_summary.py gpurun_out/ncu_layer_$tag.ncu-rep --md 2>/dev/null | tail -3
done
