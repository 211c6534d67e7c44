#!/bin/bash
mkdir -p gpurun_out
for P in fp16x2 fp16; do
STB_CUDA_PROFILER=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --profile-from-start off \
  --log-file gpurun_out/r2_launches_$P.csv python bench.py --precision $P --steps 1 --warmup 2 --no-extras --no-train --no-cpu-baseline > gpurun_out/r2_ncu_bench_$P.log 2>&1; echo "ncu $P rc=$?"
wc -l gpurun_out/r2_launches_$P.csv
done
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "head" > gpurun_out/t16_head.log 2>&1; tail -2 gpurun_out/t16_head.log
