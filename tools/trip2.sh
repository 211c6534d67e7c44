#!/bin/bash
# round-2 trip 2: the split (fp16x2) path for the first time + train16 (never run) + the whole gpu suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_split_gpu.py -m gpu -q -s -rA --timeout 600 > gpurun_out/t2_split.log 2>&1; echo "split rc=$?"
grep -E "fp16x2|passed|failed|Error|error" gpurun_out/t2_split.log | tail -40
timeout 900 python -m pytest tests/test_train16_gpu.py -m gpu -q -s -rA > gpurun_out/t2_train16.log 2>&1; echo "train16 rc=$?"
grep -iE "cos|passed|failed|error" gpurun_out/t2_train16.log | tail -20
timeout 1500 python -m pytest tests -m gpu -q -rxXs --deselect tests/test_split_gpu.py --deselect tests/test_train16_gpu.py > gpurun_out/t2_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/t2_pytest.log
