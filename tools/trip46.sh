#!/bin/bash
mkdir -p gpurun_out
for P in fp16x2 fp16; do
STB_CUDA_PROFILER=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --profile-from-start off \
  --log-file gpurun_out/r2b_launches_$P.csv python bench.py --precision $P --steps 1 --warmup 2 --no-extras --no-train --no-cpu-baseline > gpurun_out/r2b_ncu_bench_$P.log 2>&1; echo "ncu $P rc=$?"
wc -l gpurun_out/r2b_launches_$P.csv
done
for L in "32->32 k3 s1" "64->32 k3 s2T" "32->64 k3 s2"; do
  tag=$(echo "fp16x2_$L" | tr -c 'A-Za-z0-9' '_')
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3d_umma_kernel --launch-skip 3 --launch-count 1 \
    -f -o gpurun_out/r2b_ncu_$tag python tools/layer_bench.py --precision fp16x2 --only "$L" --reps 1 > gpurun_out/r2b_ncu_$tag.log 2>&1
  echo "== $L rc=$?"
  python tools/ncu_summary.py gpurun_out/r2b_ncu_$tag.ncu-rep --md 2>/dev/null | tail -1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel --launch-skip 30 --launch-count 2 \
  -f -o gpurun_out/r2b_ncu_wgrad python tools/train_step.py --precision bf16 --features tf32 --height 576 --width 960 --batch 1 --steps 1 --warmup 1 > gpurun_out/r2b_ncu_wgrad.log 2>&1
echo "== wgrad rc=$?"; python tools/ncu_summary.py gpurun_out/r2b_ncu_wgrad.ncu-rep --md 2>/dev/null | tail -2
ls -la gpurun_out/*.ncu-rep | tail -6
