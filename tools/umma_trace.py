#!/usr/bin/env python
"""In-kernel timeline of one tcgen05 conv layer (debug aid): prints per accumulator round the clocks between
issue start -> commit -> epilogue wake -> buffer release for CTA 0.   python tools/umma_trace.py [cin cout k]"""
import ctypes
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stereo_toolbox_b200 import _lib
from stereo_toolbox_b200.aggregation_umma import UmmaBackend

cin, cout, k = (int(v) for v in (sys.argv[1:4] + ["32", "32", "3"][len(sys.argv) - 1:]))
B, D, H, W = 8, 48, 96, 312
x = torch.randn(B, D, H, W, cin, device="cuda").half()
layer = nn.Sequential(nn.Conv3d(cin, cout, k, 1, k // 2, bias=False), nn.BatchNorm3d(cout)).cuda().eval()
be = UmmaBackend("fp16")
be.conv(layer, x, "relu")
R = 24
buf = torch.zeros(R, 8, dtype=torch.int64, device="cuda")
_lib.check(_lib.lib().stb_conv3d_umma_set_trace(ctypes.c_void_p(buf.data_ptr()), R), "set_trace")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
be.conv(layer, x, "relu")
e1.record()
torch.cuda.synchronize()
_lib.check(_lib.lib().stb_conv3d_umma_set_trace(ctypes.c_void_p(0), 0), "set_trace")
t = buf.cpu()
print(f"layer {cin}->{cout} k{k} on [{B},{D},{H},{W}]: {e0.elapsed_time(e1) * 1e3:.1f} us")
print("round   step_top  wait_tmem  wait_plane0  issue_len  commit->epi_wake  epi: tmem_ld  math  stores+rest   (clocks; step_top relative to round 0)")
t0 = int(t[0, 0])
for r in range(R):
    top, tmem, start, commit, wake, end = (int(v) for v in t[r][:6])
    if top == 0:
        break
    ld, mt = int(t[r][6]), int(t[r][7])
    print(f"{r:5d} {top - t0:10d} {tmem - top:10d} {start - tmem:12d} {commit - start:10d} {wake - commit:17d} {ld - wake:12d} {mt - ld:6d} {end - mt:10d}")
