#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/t28_gpu_tests.log 2>&1; echo "pytest -m gpu rc=$?"; tail -15 gpurun_out/t28_gpu_tests.log | cut -c1-250
