#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/t26_gpu_tests.log 2>&1; echo "pytest -m gpu rc=$?"; tail -8 gpurun_out/t26_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t26_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/t26_smoke.log
