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


def run(height=384, width=512, batch=2, steps=5, warmup=2, maxdisp=192, lr=1e-4, precision="fp32", features="fp32"):
    """Times `steps` complete training steps on every rank of the (already initialised, if world > 1) process group and
    returns the JSON-able result on rank 0 (None elsewhere).  Called by main() below and by bench.py's ``train_step`` leg."""
    import torch.distributed as dist
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200 import _lib
    from stereo_toolbox_b200.distrib import env_rank, FlatGradAllReduce, reduce_stats
    from stereo_toolbox_b200.synth import synth_state_dict, synth_pair, synth_gt
    rank, world, local = env_rank()
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "models.json")))["psmnet"]
    tmpl = {k: torch.zeros(s, dtype=torch.int64 if k.endswith("num_batches_tracked") else torch.float32)
            for k, s in meta["keys"].items()}
    z = np.load(os.path.join(ROOT, "tests", "golden", "bn_calib_psmnet.npz"))
    net = S.PSMNet(maxdisp)
    net.load_state_dict(synth_state_dict(tmpl, 0, {k: z[k] for k in z.files}))
    net = net.cuda().train()
    net.train_precision = precision
    net.train_features = features          # "amp": the torch 2-D extractor under autocast in the training dtype (the reference's amp recipe)
    bucket = FlatGradAllReduce(net.parameters())
    left, right = synth_pair(batch, height, width, seed=1000 + rank, shift=11)
    left, right = left.cuda(), right.cuda()
    gt = synth_gt(batch, height, width).cuda() * (maxdisp / 32.0)
    mask = (gt > 0) & (gt < maxdisp)
    maskf, nvalid = mask.float(), mask.float().sum().clamp_min(1.0)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    t_step, t_ar, t_fb = [], [], []
    launches0 = _lib.LAUNCH_COUNT
    for it in range(warmup + steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        e0.record()
        bucket.zero_()
        preds = net(left, right)
        # mean smooth-L1 over the masked pixels (trainer_torchrun.py:272-278) written as a masked reduction: the same value as
        # F.smooth_l1_loss(p[mask], gt[mask]) without boolean indexing, whose nonzero() forces a device sync per prediction
        loss = sum(w * (F.smooth_l1_loss(p.squeeze(1), gt, reduction="none") * maskf).sum() / nvalid
                   for w, p in zip((0.5, 0.7, 1.0), preds))
        loss.backward()
        e1.record()
        bucket.allreduce_()
        e2.record()
        with torch.no_grad():       # SGD update as one multi-tensor launch sequence instead of one launch per parameter
            torch._foreach_add_(list(bucket.params), [p.grad for p in bucket.params], alpha=-lr)
        e3.record()
        torch.cuda.synchronize()
        if it >= warmup:
            t_step.append(e0.elapsed_time(e3))
            t_ar.append(e1.elapsed_time(e2))
            t_fb.append(e0.elapsed_time(e1))
    # every rank must hold identical gradients after the exchange
    disagree = 0.0
    if world > 1:
        lo, hi = bucket.flat.clone(), bucket.flat.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        disagree = (hi - lo).abs().max().item()
    # the collective alone (no rank skew in front of it): the in-step figure above includes waiting for the slower rank
    ar_pure = 0.0
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        ea, eb = ev(), ev()
        ea.record()
        for _ in range(10):
            dist.all_reduce(bucket.flat, op=dist.ReduceOp.SUM)
        eb.record()
        torch.cuda.synchronize()
        ar_pure = ea.elapsed_time(eb) / 10
    avg = lambda v: sum(v) / len(v)
    (step_ms, ar_ms, fb_ms, ar_pure), (pairs,) = reduce_stats([avg(t_step), avg(t_ar), avg(t_fb), ar_pure], [float(batch)], device="cuda")
    out = None
    if rank == 0:
        bus = 2.0 * (world - 1) / world * bucket.nbytes / (ar_pure * 1e-3) / 1e9 if world > 1 else 0.0
        kind = "fp32 exact path" if precision == "fp32" else \
            f"{precision}: tcgen05 forward + dgrad and tensor-core (mma.sync) wgrad on 16-bit channels-last activations, fp32 BatchNorm statistics / volumes / head; 2-D extractor (torch): {'autocast ' + precision if features == 'amp' else features}"
        out = {
            "metric": f"PSMNet training pairs/sec ({kind}; forward+backward in libstb200.so, one flat NCCL gradient all-reduce)",
            "value": pairs / (step_ms * 1e-3), "unit": "pairs/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": step_ms, "forward_backward_ms": fb_ms, "allreduce_ms": ar_ms, "allreduce_ms_note": "inside the step: includes waiting for the slower rank's backward",
            "allreduce_alone_ms": ar_pure, "allreduce_bytes": bucket.nbytes,
            "allreduce_busbw_gbs": bus, "allreduce_share_of_step": ar_ms / step_ms,
            "overlap": "off: one all-reduce of the flat fp32 gradient buffer after backward (no bucketing, no overlap with backward)",
            "grad_disagreement_after_allreduce": disagree, "loss": loss.item(), "scaling": "weak",
            "dtype": {"fp32": "f32"}.get(precision, precision),
            "config": {"workload": f"PSMNet train step {height}x{width} D={maxdisp}", "batch_per_gpu": batch,
                       "precision": precision, "features": features, "loss": "smooth-L1 on the 3 outputs, weights 0.5/0.7/1.0, mask 0 < gt < maxdisp"},
            "gpu_launches": _lib.LAUNCH_COUNT - launches0}
    del net, bucket
    torch.cuda.empty_cache()
    return out


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
    ap.add_argument("--features", default="fp32", choices=["fp32", "tf32", "amp"],
                    help="torch 2-D extractor: exact fp32, or autocast in the training dtype (the reference's amp recipe)")
    args = ap.parse_args()
    import torch.distributed as dist
    from stereo_toolbox_b200.distrib import env_rank
    rank, world, local = env_rank()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = run(args.height, args.width, args.batch, args.steps, args.warmup, args.maxdisp, args.lr, args.precision, args.features)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
