#!/bin/bash
# Is the stride-2 conv's plane load slow because of TMA elementStrides?  (profiles/next_round_plan.md section 3)
cd "$(dirname "$0")"
for C in 32 64; do for R in 2 4; do for mode in 0 1 2; do
  timeout 60 ./umma_probe tmabw2 $C 10 $R $mode || true
done; done; done
