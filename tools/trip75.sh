#!/bin/bash
# ncu --set full of the two transposed-conv flavours at the last commit: 64->32 s2T (LEAN 6, one pass) and 128->64 s2T (LEAN 9, two passes)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3d_umma_kernel --launch-skip 1 --launch-count 1 \
  -f -o gpurun_out/r2c_ncu_s2T_64_32 python tools/layer_bench.py --precision fp16x2 --only "64->32 k3 s2T" --reps 1 > gpurun_out/r2c_ncu_s2T_64_32.log 2>&1
echo "rc=$?"; python tools/ncu_summary.py gpurun_out/r2c_ncu_s2T_64_32.ncu-rep --md 2>/dev/null | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3d_umma_kernel --launch-skip 2 --launch-count 2 \
  -f -o gpurun_out/r2c_ncu_s2T_128_64 python tools/layer_bench.py --precision fp16x2 --only "128->64 k3 s2T" --reps 1 > gpurun_out/r2c_ncu_s2T_128_64.log 2>&1
echo "rc=$?"; python tools/ncu_summary.py gpurun_out/r2c_ncu_s2T_128_64.ncu-rep --md 2>/dev/null | tail -2
ls -la gpurun_out/r2c_ncu_s2T_*.ncu-rep
