#!/bin/bash
# Runs every probe variant in its own process (a faulting variant must not poison the rest).
cd "$(dirname "$0")"
out=${1:-/dev/stdout}
{
timeout 60 ./umma_probe halo
timeout 60 ./umma_probe halo2 0
timeout 60 ./umma_probe halo2 1
for sw in 128 64 32; do
  for bo in 0 1; do
    for sh in 0 1 2 3 5 8 9; do
      timeout 60 ./umma_probe $sw $sh $bo 32 || true
    done
  done
done
timeout 60 ./umma_probe 128 0 0 64
timeout 60 ./umma_probe 128 3 0 128
timeout 60 ./umma_probe 64 3 0 64
} > $out 2>&1
