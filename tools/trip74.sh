#!/bin/bash
# ncu launch list (time + DRAM bytes) of one timed bench step at the last commit of the round
mkdir -p gpurun_out
P=fp16x2
STB_CUDA_PROFILER=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --profile-from-start off \
  --log-file gpurun_out/r2c_launches_$P.csv python bench.py --precision $P --steps 1 --warmup 2 --no-extras --no-train --no-cpu-baseline > gpurun_out/r2c_ncu_bench_$P.log 2>&1; echo "ncu $P rc=$?"
wc -l gpurun_out/r2c_launches_$P.csv
