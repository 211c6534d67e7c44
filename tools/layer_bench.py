#!/usr/bin/env python
"""Per-layer timing of the tcgen05 conv family on the layer shapes of GwcNet_GC at the KITTI benchmark shape
(B=8, 1/4 res 48x96x312) -- the iteration loop for kernel work, much shorter than a full bench.py run.

  python tools/layer_bench.py [--only SUBSTR] [--reps 7] [--json out.json]

CUDA events on the launching stream, 512 MB L2 flush before every timed launch, median of `reps`.
STB_UMMA_VERBOSE=1 makes the C host entry print the tiling it chose for every launch."""
import argparse
import json
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stereo_toolbox_b200.aggregation_umma import UmmaBackend  # noqa: E402

B = 8
# (name, cin, cout, k, stride, transposed, (D,H,W) of the INPUT, residual?, act, count per forward)
LAYERS = [
    ("64->32 k3 s1", 64, 32, 3, 1, False, (48, 96, 312), False, "relu", 1),
    ("32->32 k3 s1", 32, 32, 3, 1, False, (48, 96, 312), False, "relu", 4),
    ("32->64 k3 s2", 32, 64, 3, 2, False, (48, 96, 312), False, "relu", 3),
    ("64->64 k3 s1", 64, 64, 3, 1, False, (24, 48, 156), False, "relu", 3),
    ("64->128 k3 s2", 64, 128, 3, 2, False, (24, 48, 156), False, "relu", 3),
    ("128->128 k3 s1", 128, 128, 3, 1, False, (12, 24, 78), False, "relu", 3),
    ("64->64 k1 s1", 64, 64, 1, 1, False, (24, 48, 156), False, "none", 3),
    ("128->64 k3 s2T", 128, 64, 3, 2, True, (12, 24, 78), True, "relu", 3),
    ("32->32 k1 s1", 32, 32, 1, 1, False, (48, 96, 312), False, "none", 3),
    ("64->32 k3 s2T", 64, 32, 3, 2, True, (24, 48, 156), True, "relu", 3),
    ("32->1 k3 s1", 32, 1, 3, 1, False, (48, 96, 312), False, "none", 1),
]


def make_layer(cin, cout, k, stride, tr, bn=True):
    if tr:
        conv = nn.ConvTranspose3d(cin, cout, k, stride, 1, output_padding=1, bias=False)
    else:
        conv = nn.Conv3d(cin, cout, k, stride, k // 2, bias=False)
    mods = [conv] + ([nn.BatchNorm3d(cout)] if bn else [])
    return nn.Sequential(*mods).cuda().eval()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--json", default="")
    ap.add_argument("--precision", default="fp16")
    args = ap.parse_args()
    torch.manual_seed(0)
    be = UmmaBackend(args.precision)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    rows, total = [], 0.0
    for name, cin, cout, k, s, tr, (D, H, W), has_res, act, count in LAYERS:
        if args.only and args.only not in name:
            continue
        x = torch.randn(B, D, H, W, cin * getattr(be, "cmul", 1), device="cuda").to(be.dtype)
        layer = make_layer(cin, cout, k, s, tr, bn=cout > 1)
        out = be.conv(layer, x, act)
        res = torch.randn_like(out) if has_res else None
        ts = []
        for _ in range(args.reps):
            flush.zero_()
            torch.cuda._sleep(1_000_000)      # keep the GPU busy while the host prepares the launch (no idle gap in the bracket)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            be.conv(layer, x, act, res)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        vox_out = out.numel() // out.shape[-1]
        taps = k ** 3 if not tr else k ** 3 / 8.0
        flops = 2.0 * taps * cin * cout * vox_out
        row = {"layer": name, "us": round(us, 1), "tflops": round(flops / us * 1e-6, 1), "per_forward_us": round(us * count, 1)}
        total += us * count
        rows.append(row)
        print(json.dumps(row), flush=True)
        del x, out, res, layer
    print(json.dumps({"aggregation_total_ms": round(total * 1e-3, 3)}))
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
