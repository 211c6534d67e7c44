#!/bin/bash
cd "$(dirname "$0")"
for N in 32 96 192 256; do for rowb in 64 128; do for sh in 0 3; do for ni in 1 2; do
  timeout 60 ./umma_probe mmarate $N $rowb $sh $ni || true
done; done; done; done
