#!/bin/bash
mkdir -p gpurun_out
STB_UMMA_VERBOSE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-extras --no-train --no-cpu-baseline > gpurun_out/t10_bench.json 2> gpurun_out/t10_verbose.log; echo "rc=$?"
grep "^\[stb" gpurun_out/t10_verbose.log | sort | uniq -c | sort -rn | cut -c1-330 | head -70
