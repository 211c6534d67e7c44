#!/usr/bin/env python
"""One data-parallel training step of PSMNet on the CUDA hot path (BASELINE config 3 family), timed.

    python tools/train_step.py --steps 5 --warmup 2                       # 1 GPU
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/train_step.py ...

Every rank holds a replica and its own shard of the batch (synthetic pairs of the reference's training crop,
384x512, trainer data_augmentation/__init__.py:22); forward + loss (smooth-L1 on the three outputs, weights 0.5/0.7/1.0,
mask 0 < gt < maxdisp: trainer/trainer_torchrun.py:272-278) + backward run through libstb200.so (autograd.py), then the
ONE collective of the path: the gradient all-reduce over NCCL (distrib.FlatGradAllReduce: a single flat fp32 buffer),
then an SGD-style update so that the step is complete.  Prints one JSON line (rank 0): pairs/s over all ranks (max-over-ranks
device time), all-reduce time and bus bandwidth, and the largest gradient disagreement between ranks after the exchange
(must be 0).  fp32 exact path: this tool demonstrates and measures the exchange step, it is not the headline benchmark.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--batch", type=int, default=2, help="pairs per GPU per step")
    ap.add_argument("--height", type=int, default=384)
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--maxdisp", type=int, default=192)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16", "fp16"],
                    help="fp32 = exact training path; bf16 / fp16 = train16.Umma16TrainBackend (tcgen05 forward + dgrad)")
    args = ap.parse_args()
    import torch.distributed as dist
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.distrib import env_rank, FlatGradAllReduce, reduce_stats
    from stereo_toolbox_b200.synth import synth_state_dict, synth_pair, synth_gt
    rank, world, local = env_rank()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "models.json")))["psmnet"]
    tmpl = {k: torch.zeros(s, dtype=torch.int64 if k.endswith("num_batches_tracked") else torch.float32)
            for k, s in meta["keys"].items()}
    z = np.load(os.path.join(ROOT, "tests", "golden", "bn_calib_psmnet.npz"))
    net = S.PSMNet(args.maxdisp)
    net.load_state_dict(synth_state_dict(tmpl, 0, {k: z[k] for k in z.files}))
    net = net.cuda().train()
    net.train_precision = args.precision
    bucket = FlatGradAllReduce(net.parameters())
    left, right = synth_pair(args.batch, args.height, args.width, seed=1000 + rank, shift=11)
    left, right = left.cuda(), right.cuda()
    gt = synth_gt(args.batch, args.height, args.width).cuda() * (args.maxdisp / 32.0)
    mask = (gt > 0) & (gt < args.maxdisp)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    t_step, t_ar = [], []
    for it in range(args.warmup + args.steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        e0.record()
        bucket.zero_()
        preds = net(left, right)
        loss = sum(w * F.smooth_l1_loss(p.squeeze(1)[mask], gt[mask], reduction="mean") for w, p in zip((0.5, 0.7, 1.0), preds))
        loss.backward()
        e1.record()
        bucket.allreduce_()
        e2.record()
        with torch.no_grad():
            for p in bucket.params:
                p.add_(p.grad, alpha=-args.lr)
        e3.record()
        torch.cuda.synchronize()
        if it >= args.warmup:
            t_step.append(e0.elapsed_time(e3))
            t_ar.append(e1.elapsed_time(e2))
    # every rank must hold identical gradients after the exchange
    disagree = 0.0
    if world > 1:
        lo, hi = bucket.flat.clone(), bucket.flat.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        disagree = (hi - lo).abs().max().item()
    (step_ms, ar_ms), (pairs,) = reduce_stats([sum(t_step) / len(t_step), sum(t_ar) / len(t_ar)], [float(args.batch)], device="cuda")
    if rank == 0:
        bus = 2.0 * (world - 1) / world * bucket.nbytes / (ar_ms * 1e-3) / 1e9 if world > 1 else 0.0
        print(json.dumps({
            "metric": "PSMNet training pairs/sec (fp32 exact path, forward+backward in libstb200.so, flat NCCL grad all-reduce)",
            "value": pairs / (step_ms * 1e-3), "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "allreduce_ms": ar_ms, "allreduce_bytes": bucket.nbytes, "allreduce_busbw_gbs": bus,
            "grad_disagreement_after_allreduce": disagree, "loss": loss.item(), "scaling": "weak", "dtype": "f32",
            "config": {"workload": f"PSMNet train step {args.height}x{args.width} D={args.maxdisp}", "batch_per_gpu": args.batch,
                       "precision": args.precision},
            "gpu_launches": __import__("stereo_toolbox_b200._lib", fromlist=["x"]).LAUNCH_COUNT}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
