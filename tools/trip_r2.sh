#!/bin/bash
# GPU trips prepared at the end of round 1 (run under gpurun from the repo root; everything lands in gpurun_out/):
#   bash tools/trip_r2.sh a    confirm what was written without hardware + the headline bench            (~10 min)
#   bash tools/trip_r2.sh b    whole-model numbers for the other BASELINE configs, Table 3, training      (~15 min)
#   bash tools/trip_r2.sh c    probes / sweeps / ncu captures / sanitizers for the kernel work            (~40 min)
# e.g.  gpurun --timeout 1500 -- 'bash tools/trip_r2.sh a'     (no argument = a, b and c in that order)
PARTS=${1:-abc}
mkdir -p gpurun_out
if [[ $PARTS == *a* ]]; then
timeout 1500 python -m pytest tests -m gpu -x -q -rxXs > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest.log
export STB200_RUN_UNCONFIRMED=1      # from here on also run the tests that launch not-yet-confirmed kernels (skipped by default)
timeout 600 python -m pytest tests/test_igev_stereo_gpu.py tests/test_lowp_model_gpu.py tests/test_raft_train_gpu.py tests/test_igev_train_gpu.py tests/test_cascade_train_gpu.py tests/test_sampled_volume_gpu.py tests/test_properties_fullsize_gpu.py -m gpu -q -s --runxfail -rA > gpurun_out/r2_igev.log 2>&1; echo "igev rc=$?"
grep -i "EPE\|passed\|failed\|storage model" gpurun_out/r2_igev.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/r2_bench.json
fi
if [[ $PARTS == *b* ]]; then
export STB200_RUN_UNCONFIRMED=1
{
  timeout 300 python tools/model_bench.py --model igev   --height 1152 --width 1920 --maxdisp 256 --iters 32 --precision fp16
  timeout 300 python tools/model_bench.py --model igev   --height 1152 --width 1920 --maxdisp 256 --iters 32 --precision fp32
  timeout 300 python tools/model_bench.py --model acvnet --height 1152 --width 1920 --maxdisp 256 --precision fp16
  timeout 300 python tools/model_bench.py --model raft   --height 512  --width 1024 --iters 32 --cuda-graph
  timeout 300 python tools/model_bench.py --model raft   --height 512  --width 1024 --iters 32 --cuda-graph --channels-last
  timeout 300 python tools/model_bench.py --model igev   --height 1152 --width 1920 --maxdisp 256 --iters 32 --precision fp16 --channels-last
  timeout 300 python tools/model_bench.py --model igev   --height 480 --width 640 --iters 32 --precision fp16
  timeout 300 python tools/model_bench.py --model igev   --height 480 --width 640 --iters 32 --precision fp16 --cuda-graph
  STB_CFNET_SAMPLED=1 timeout 300 python tools/model_bench.py --model cfnet --height 384 --width 1248 --precision fp16
  for cl in "" "--channels-last"; do
    timeout 300 python tools/model_bench.py --model pcwnet_gc --height 384 --width 1248 --precision fp16 $cl
    timeout 300 python tools/model_bench.py --model cfnet     --height 384 --width 1248 --precision fp16 $cl
  done
  timeout 300 python tools/model_bench.py --model psmnet --height 576  --width 960  --batch 4 --precision fp16
  timeout 300 python tools/model_bench.py --model gwcnet_gc --height 576 --width 960 --batch 4 --precision fp16
} > gpurun_out/r2_models.jsonl 2> gpurun_out/r2_models.err
cat gpurun_out/r2_models.jsonl | cut -c1-300
# second north_star shape through the full bench contract
timeout 600 python bench.py --workload sceneflow --steps 10 --warmup 3 > gpurun_out/r2_bench_sceneflow.json 2> gpurun_out/r2_bench_sceneflow.err; echo "bench sceneflow rc=$?"
cut -c1-300 gpurun_out/r2_bench_sceneflow.json
STB_HEAD_X4=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_headx4.json 2> gpurun_out/r2_bench_headx4.err; echo "bench head-x4 rc=$?"
cut -c1-300 gpurun_out/r2_bench_headx4.json
STB_HEAD_X4=1 STB_UMMA_CLS1=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_headx4_cls1.json 2> gpurun_out/r2_bench_headx4_cls1.err; echo "bench head-x4 + cls1 rc=$?"
cut -c1-300 gpurun_out/r2_bench_headx4_cls1.json
STB_HEAD_X4=1 STB_UMMA_CLS1=1 STB_UMMA_T2PAIR=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_all_optins.json 2> gpurun_out/r2_bench_all_optins.err; echo "bench all opt-ins rc=$?"
cut -c1-300 gpurun_out/r2_bench_all_optins.json
STB_UMMA_CLS1=1 timeout 200 python tools/layer_bench.py --only "32->1 k3" > gpurun_out/r2_layer_cls1.log 2>&1; python tools/layer_bench.py --only "32->1 k3" >> gpurun_out/r2_layer_cls1.log 2>&1; cat gpurun_out/r2_layer_cls1.log
STB_UMMA_T2PAIR=1 timeout 200 python tools/layer_bench.py --only "64->32 k3 s2T" > gpurun_out/r2_layer_t2pair.log 2>&1; python tools/layer_bench.py --only "64->32 k3 s2T" >> gpurun_out/r2_layer_t2pair.log 2>&1; cat gpurun_out/r2_layer_t2pair.log
# BASELINE config 3 shape: exact fp32 training path, then the mixed-precision backend (train16.py, unconfirmed)
timeout 600 python tools/train_step.py --height 576 --width 960 --batch 1 --steps 3 --warmup 1 > gpurun_out/r2_train_sceneflow.json 2> gpurun_out/r2_train_sceneflow.err; echo "train rc=$?"
cut -c1-300 gpurun_out/r2_train_sceneflow.json
timeout 600 python -m pytest tests/test_train16_gpu.py -m gpu -q -s --runxfail > gpurun_out/r2_train16.log 2>&1; echo "train16 rc=$?"
grep -i "cos\|passed\|failed\|error" gpurun_out/r2_train16.log | tail -20
timeout 600 python tools/train_step.py --height 576 --width 960 --batch 1 --steps 3 --warmup 1 --precision bf16 > gpurun_out/r2_train_sceneflow_bf16.json 2> gpurun_out/r2_train_sceneflow_bf16.err; echo "train bf16 rc=$?"
cut -c1-300 gpurun_out/r2_train_sceneflow_bf16.json
# the reference's published Table 3 protocol (RTX 4090 numbers in BASELINE.md) on the drop-in models
timeout 1500 python tools/table3.py --iters 10 > gpurun_out/r2_table3.md 2> gpurun_out/r2_table3.err; echo "table3 rc=$?"
cat gpurun_out/r2_table3.md | cut -c1-260
fi
if [[ $PARTS == *c* ]]; then
# does TMA elementStrides throttle the stride-2 convs' plane loads? (unit stride vs elementStrides vs parity-folded map)
timeout 300 bash stereo_toolbox_b200/csrc/probe/run_tmabw2.sh > gpurun_out/r2_tmabw2.txt 2>&1; cat gpurun_out/r2_tmabw2.txt
# tiling sweep of the slow conv flavours (TH=0 / RING=0 = the built-in choice)
timeout 1800 bash tools/sweep_tiling.sh > gpurun_out/r2_sweep_tiling.log 2>&1; tail -90 gpurun_out/r2_sweep_tiling.log
# per-layer ncu captures of the slow conv flavours (reports come back in gpurun_out/)
timeout 2400 bash tools/ncu_layers.sh > gpurun_out/r2_ncu_layers.log 2>&1; tail -60 gpurun_out/r2_ncu_layers.log
# sanitizers last (slow)
timeout 2400 bash tools/sanitize.sh > gpurun_out/r2_sanitize.log 2>&1; tail -12 gpurun_out/r2_sanitize.log
fi
