#!/bin/bash
cd "$(dirname "$0")"
for C in 32 64; do for rows in 10 18; do for R in 1 2 4 8; do
  timeout 60 ./umma_probe tmabw $C $rows $R || true
done; done; done
