#!/usr/bin/env python
"""Whole-model throughput of any drop-in model on the CUDA hot path, for the BASELINE configs bench.py does not carry
(bench.py is config 2; the others are parity-test cases there):

  config 4   python tools/model_bench.py --model raft --height 512 --width 1024 --iters 32
  config 5   python tools/model_bench.py --model acvnet --height 1152 --width 1920 --maxdisp 256 --precision fp16
             python tools/model_bench.py --model igev   --height 1152 --width 1920 --maxdisp 256 --iters 32
  SceneFlow  python tools/model_bench.py --model psmnet --height 576 --width 960 --batch 4 --precision fp16

Random-init name-keyed weights (synth.synth_state_dict), synthetic pair, eval + no_grad, CUDA events, median of --reps.
Prints one JSON line: maps/s, ms per forward, peak memory.  (Parity is the tests' job; the oracle is not imported here.)
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def build(args, S):
    p = args.precision
    m = args.model
    if m == "gwcnet_gc":
        return S.GwcNet_GC(args.maxdisp, precision=p)
    if m == "gwcnet_g":
        return S.GwcNet_G(args.maxdisp, precision=p)
    if m == "psmnet":
        return S.PSMNet(args.maxdisp, precision=p)
    if m == "acvnet":
        return S.ACVNet(args.maxdisp, precision=p)
    if m == "cfnet":
        return S.CFNet(args.maxdisp, precision=p)
    if m == "pcwnet_gc":
        return S.PCWNet_GC(args.maxdisp, precision=p)
    if m == "raft":
        return S.RAFTStereo()
    if m == "igev":
        return S.IGEVStereo({"max_disp": args.maxdisp}, precision=p)
    raise SystemExit(f"unknown model {m}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", required=True,
                    choices=["gwcnet_gc", "gwcnet_g", "psmnet", "acvnet", "cfnet", "pcwnet_gc", "raft", "igev"])
    ap.add_argument("--height", type=int, default=384)
    ap.add_argument("--width", type=int, default=1248)
    ap.add_argument("--maxdisp", type=int, default=192)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--iters", type=int, default=32, help="GRU iterations (raft / igev)")
    ap.add_argument("--precision", default="fp16", choices=["fp32", "fp16x2", "fp16", "bf16"])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--cuda-graph", action="store_true", help="raft / igev: replay one captured GRU iteration")
    ap.add_argument("--update", default="", choices=["", "torch", "umma"],
                    help="raft / igev: update block in torch/cuDNN or on the tcgen05 2-D conv path (update_umma.py, exact fp16x2 format)")
    ap.add_argument("--features", default="", help="2-D extractor mode of the 3-D-conv models (fp32 | tf32 | tf32_cl | fp16 | umma); default: the model's")
    ap.add_argument("--exact-glue", action="store_true", help="torch glue in true fp32 (no TF32): the arithmetic the parity tests pin")
    ap.add_argument("--channels-last", action="store_true", help="raft / igev / cfnet / pcwnet_gc: NHWC torch glue (model.channels_last)")
    args = ap.parse_args()

    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_pair, synth_state_dict
    net = build(args, S)
    net.load_state_dict(synth_state_dict(net.state_dict(), 0), strict=True)
    net = net.cuda().eval()
    if args.cuda_graph:
        net.cuda_graph = True
    if args.channels_last:
        net.channels_last = True
    if args.update:
        net.update_mode = args.update
    if args.features:
        net.feature_mode = args.features
    if args.exact_glue:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    left, right = synth_pair(args.batch, args.height, args.width, seed=4, shift=9)
    gl, gr = left.cuda(), right.cuda()
    fwd = dict(iters=args.iters) if args.model in ("raft", "igev") else {}
    torch.cuda.reset_peak_memory_stats()
    with torch.no_grad():
        for _ in range(args.warmup):
            out = net(gl, gr, **fwd)
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = net(gl, gr, **fwd)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    res = dict(model=args.model, precision=args.precision if args.model != "raft" else "fp32",
               shape=[args.batch, args.height, args.width], maxdisp=args.maxdisp,
               iters=fwd.get("iters"), cuda_graph=args.cuda_graph, channels_last=args.channels_last, update=args.update or "auto", features=args.features or "default", exact_glue=args.exact_glue, ms_per_forward=ms, maps_per_s=args.batch / (ms * 1e-3),
               peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30, out_shape=list(out.shape),
               finite=bool(torch.isfinite(out.float()).all().item()))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
