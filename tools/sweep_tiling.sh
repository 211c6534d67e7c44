#!/bin/bash
# Tile-height / ring-depth sweep of the tcgen05 conv on the slow layer flavours (host-side overrides STB_UMMA_TH,
# STB_UMMA_RING of stb_conv3d_umma; no rebuild).  Run under gpurun; prints one line per (layer, TH, ring).
for L in "64->32 k3 s2T" "128->64 k3 s2T" "32->64 k3 s2" "64->128 k3 s2" "32->32 k3 s1" "64->64 k3 s1"; do
  for TH in 0 4 8 16; do
    for RING in 0 4; do
      out=$(STB_UMMA_TH=$TH STB_UMMA_RING=$RING timeout 120 python tools/layer_bench.py --only "$L" --reps 5 2>/dev/null | head -1)
      echo "TH=$TH RING=$RING $out"
    done
  done
done
