#!/bin/bash
mkdir -p gpurun_out
for prec in bf16 fp32; do
timeout 900 python tools/train_step.py --height 576 --width 960 --batch 1 --steps 3 --warmup 1 --precision $prec > gpurun_out/t8_train_$prec.json 2> gpurun_out/t8_train_$prec.err; echo "train $prec rc=$?"
cut -c1-500 gpurun_out/t8_train_$prec.json; tail -3 gpurun_out/t8_train_$prec.err
done
timeout 900 python tools/train_step.py --height 576 --width 960 --batch 2 --steps 3 --warmup 1 --precision bf16 > gpurun_out/t8_train_bf16_b2.json 2> gpurun_out/t8_train_bf16_b2.err; echo "train bf16 b2 rc=$?"
cut -c1-500 gpurun_out/t8_train_bf16_b2.json; tail -3 gpurun_out/t8_train_bf16_b2.err
