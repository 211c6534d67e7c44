#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rxXs -x --timeout 1500 > gpurun_out/t7_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/t7_pytest.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t7_smoke.log 2>&1; echo "smoke rc=$?"; tail -12 gpurun_out/t7_smoke.log
