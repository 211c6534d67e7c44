#!/usr/bin/env python
"""Where one PSMNet training step (BASELINE config 3 shape) spends its GPU time: torch.profiler (kineto) over one step,
kernels grouped by family.  python tools/train_profile.py [--precision bf16] [--height 576 --width 960]"""
import argparse
import collections
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FAMILIES = [("conv3d_umma", "conv3d_umma_kernel"), ("wgrad_tc", "wgrad_kernel"), ("wgrad_f32", "wgrad_f32_kernel"),
            ("volume fwd/bwd", "volume"), ("concat/gwc bwd", "_bwd_kernel"), ("head", "softargmin"), ("layout", "cl16"),
            ("layout", "ncdhw"), ("cudnn / cublas (2-D extractor)", "cudnn"), ("cudnn / cublas (2-D extractor)", "cutlass"),
            ("cudnn / cublas (2-D extractor)", "gemm"), ("cudnn / cublas (2-D extractor)", "conv"), ("batch_norm", "batch_norm"),
            ("batch_norm", "bn_"), ("elementwise (torch)", "elementwise"), ("reduce (torch)", "reduce"), ("copy / cat", "copy"),
            ("copy / cat", "Cat")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--height", type=int, default=576)
    ap.add_argument("--width", type=int, default=960)
    ap.add_argument("--maxdisp", type=int, default=192)
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--features", default="tf32")
    ap.add_argument("--cpu", action="store_true", help="also: host issue time of a step and the host-side top list")
    args = ap.parse_args()
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_state_dict, synth_pair, synth_gt
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "models.json")))["psmnet"]
    tmpl = {k: torch.zeros(s, dtype=torch.int64 if k.endswith("num_batches_tracked") else torch.float32)
            for k, s in meta["keys"].items()}
    z = np.load(os.path.join(ROOT, "tests", "golden", "bn_calib_psmnet.npz"))
    net = S.PSMNet(args.maxdisp)
    net.load_state_dict(synth_state_dict(tmpl, 0, {k: z[k] for k in z.files}))
    net = net.cuda().train()
    net.train_precision = args.precision
    net.train_features = args.features
    left, right = synth_pair(1, args.height, args.width, seed=1000, shift=11)
    left, right = left.cuda(), right.cuda()
    gt = synth_gt(1, args.height, args.width).cuda() * (args.maxdisp / 32.0)
    mask = (gt > 0) & (gt < args.maxdisp)
    maskf, nvalid = mask.float(), mask.float().sum().clamp_min(1.0)

    def step():
        for p in net.parameters():
            p.grad = None
        preds = net(left, right)
        loss = sum(w * (F.smooth_l1_loss(p.squeeze(1), gt, reduction="none") * maskf).sum() / nvalid
                   for w, p in zip((0.5, 0.7, 1.0), preds))       # masked mean without boolean indexing (no device sync)
        loss.backward()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    fam = collections.defaultdict(lambda: [0.0, 0])
    names = collections.defaultdict(lambda: [0.0, 0])
    total = 0.0
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        us = ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        name = ev.name
        total += us
        key = "other"
        for f, pat in FAMILIES:
            if pat in name:
                key = f
                break
        fam[key][0] += us; fam[key][1] += 1
        names[name[:90]][0] += us; names[name[:90]][1] += 1
    print(f"GPU time of one step: {total / 1e3:.1f} ms over {sum(v[1] for v in fam.values())} kernels")
    for k, (us, n) in sorted(fam.items(), key=lambda kv: -kv[1][0]):
        print(f"  {k:38s} {us / 1e3:8.2f} ms  {100 * us / total:5.1f} %  {n:5d} launches")
    if args.cpu:
        import time
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()
        t_host = time.perf_counter() - t0          # host time to ISSUE a step (no sync inside unless the step syncs itself)
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        print(f"host issue time of one step: {t_host * 1e3:.1f} ms, until the GPU is done: {t_all * 1e3:.1f} ms")
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof2:
            step()
            torch.cuda.synchronize()
        print(prof2.key_averages().table(sort_by="self_cpu_time_total", row_limit=30, max_name_column_width=60))
    print("top kernels:")
    for k, (us, n) in sorted(names.items(), key=lambda kv: -kv[1][0])[:args.top]:
        print(f"  {us / 1e3:8.2f} ms {n:5d}x  {k}")


if __name__ == "__main__":
    main()
