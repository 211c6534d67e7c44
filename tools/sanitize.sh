#!/bin/bash
# compute-sanitizer passes over the small golden-fixture GPU tests (race detection / sanitizers).
# Run under gpurun from the repo root; writes gpurun_out/sanitize_<tool>.log.  The tests are the tiny fixture cases, so
# each pass stays within a few minutes even at the sanitizer's ~50x slowdown.
mkdir -p gpurun_out
SEL='test_volumes_golden or test_head_golden or test_corr_golden or test_geo_golden or test_patch_dw_golden_and_slices or test_feature_gate_layouts or test_align_corners_head_and_variance'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --target-processes all --error-exitcode 86 \
    python -m pytest tests/test_gpu_ops.py tests/test_gpu_blocks.py -m gpu -q -x -k "$SEL" \
    > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_$tool.log | tail -3
done
# the tcgen05 / TMA kernel: one small 16-bit conv case under memcheck only (racecheck does not model TMA / UMMA async proxies)
timeout 900 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 86 \
  python -m pytest tests/test_gpu_umma.py -m gpu -q -x -k "test_layout_roundtrip or (test_conv_family_16bit and fp16)" --maxfail=1 \
  > gpurun_out/sanitize_umma_memcheck.log 2>&1
echo "umma memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_umma_memcheck.log | tail -3
