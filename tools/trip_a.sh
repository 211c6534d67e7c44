#!/bin/bash
# GPU trip A: tests, bench, per-op bench, ncu captures (run under gpurun from the repo root)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/bench.json
timeout 400 python tools/op_bench.py --json gpurun_out/op_bench.json > gpurun_out/op_bench.log 2>&1; echo "op_bench rc=$?"
cat gpurun_out/op_bench.log | cut -c1-200
STB_VOLUME_V1=1 timeout 200 python tools/op_bench.py --only volume_cl16 > gpurun_out/op_bench_v1.log 2>&1
cat gpurun_out/op_bench_v1.log | cut -c1-200
STB_OPBENCH_NCU=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/ops_full python tools/op_bench.py > gpurun_out/ncu_ops.log 2>&1; echo "ncu ops rc=$?"
STB_CUDA_PROFILER=1 STB_BENCH_NOPROF=1 timeout 900 ncu --profile-from-start off \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none --csv --log-file gpurun_out/launches_metrics.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu bench rc=$?"
ls -la gpurun_out
