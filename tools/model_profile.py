#!/usr/bin/env python
"""GPU time of one forward of a drop-in model by kernel (torch.profiler); same model arguments as tools/model_bench.py.
   python tools/model_profile.py --model igev --height 1152 --width 1920 --maxdisp 256 --iters 32 [--precision fp16]"""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    import argparse
    import model_bench as MB
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="igev")
    ap.add_argument("--height", type=int, default=1152)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--maxdisp", type=int, default=256)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--iters", type=int, default=32)
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--top", type=int, default=30)
    args = ap.parse_args()
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_pair, synth_state_dict
    net = MB.build(args, S)
    net.load_state_dict(synth_state_dict(net.state_dict(), 0), strict=True)
    net = net.cuda().eval()
    left, right = synth_pair(args.batch, args.height, args.width, seed=4, shift=9)
    gl, gr = left.cuda(), right.cuda()
    fwd = dict(iters=args.iters) if args.model in ("raft", "igev") else {}
    with torch.no_grad():
        for _ in range(2):
            net(gl, gr, **fwd)
        torch.cuda.synchronize()
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            net(gl, gr, **fwd)
            torch.cuda.synchronize()
    names = collections.defaultdict(lambda: [0.0, 0])
    total = 0.0
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        us = ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        total += us
        names[ev.name[:110]][0] += us
        names[ev.name[:110]][1] += 1
    print(f"GPU time of one forward: {total / 1e3:.1f} ms over {sum(v[1] for v in names.values())} kernels")
    for k, (us, n) in sorted(names.items(), key=lambda kv: -kv[1][0])[:args.top]:
        print(f"  {us / 1e3:8.2f} ms {n:5d}x  {us / n:8.1f} us  {k}")


if __name__ == "__main__":
    main()
