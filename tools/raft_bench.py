#!/usr/bin/env python
"""RAFT-Stereo on the CUDA hot path (BASELINE config 4: 32 iterations, 1024x512 synthetic pair): eager loop vs the
CUDA-graph-replayed iteration, and the share of the CorrBlock1D kernels.   python tools/raft_bench.py [--iters 32]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=32)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_pair
    torch.manual_seed(0)
    net = S.RAFTStereo().cuda().eval()
    left, right = synth_pair(args.batch, args.height, args.width, seed=4, shift=9)
    left, right = left.cuda(), right.cuda()
    out = {}
    ref = None
    for mode in ("eager", "cuda_graph"):
        net.cuda_graph = mode == "cuda_graph"
        with torch.no_grad():
            for _ in range(2):
                d = net(left, right, iters=args.iters)
            torch.cuda.synchronize()
            ts = []
            for _ in range(args.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                d = net(left, right, iters=args.iters)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
        ts.sort()
        out[mode + "_ms"] = ts[len(ts) // 2]
        if ref is None:
            ref = d
        else:
            out["max_abs_diff_vs_eager"] = (d - ref).abs().max().item()
    out["speedup"] = out["eager_ms"] / out["cuda_graph_ms"]
    out["maps_per_s_graph"] = args.batch / (out["cuda_graph_ms"] * 1e-3)
    out["config"] = f"RAFTStereo {args.iters} iters, {args.height}x{args.width}, batch {args.batch}"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
