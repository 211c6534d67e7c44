#!/usr/bin/env python
"""GPU time of one RAFT-Stereo forward by kernel (torch.profiler): python tools/raft_profile.py [--update umma] [--iters 32]"""
import argparse
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--update", default="umma")
    ap.add_argument("--iters", type=int, default=32)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--top", type=int, default=30)
    args = ap.parse_args()
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_pair, synth_state_dict
    net = S.RAFTStereo()
    net.load_state_dict(synth_state_dict(net.state_dict(), 0), strict=True)
    net = net.cuda().eval()
    net.update_mode = args.update
    left, right = synth_pair(1, args.height, args.width, seed=4, shift=9)
    gl, gr = left.cuda(), right.cuda()
    with torch.no_grad():
        for _ in range(2):
            net(gl, gr, iters=args.iters)
        torch.cuda.synchronize()
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            net(gl, gr, iters=args.iters)
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
