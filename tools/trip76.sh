#!/bin/bash
# ncu --set full of 64->32 s2T WITH residual (launch 0 = smem probe, 1 = first call without residual, 2 = the timed call); pages exported, report dropped
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3d_umma_kernel --launch-skip 2 --launch-count 1 \
  -f -o /tmp/r2c_s2T python tools/layer_bench.py --precision fp16x2 --only "64->32 k3 s2T" --reps 1 > gpurun_out/r2c_ncu_s2T_64_32.log 2>&1
echo "rc=$?"; python tools/ncu_summary.py /tmp/r2c_s2T.ncu-rep --md 2>/dev/null | tail -1
ncu -i /tmp/r2c_s2T.ncu-rep --page raw --csv > gpurun_out/r2c_s2T_64_32_raw.csv 2>/dev/null
ncu -i /tmp/r2c_s2T.ncu-rep --page source --csv > gpurun_out/r2c_s2T_64_32_source.csv 2>/dev/null
ncu -i /tmp/r2c_s2T.ncu-rep --page details > gpurun_out/r2c_s2T_64_32_details.txt 2>/dev/null
ls -la gpurun_out/r2c_s2T_64_32_*
