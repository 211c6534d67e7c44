"""The reference's whole-model latency / memory protocol (evaluation/speed_and_memory_test.py:11-79), for drop-in models.

Same call and return as the reference's ``speed_and_memory_test`` -- the function behind its published Table 3
(README.md:182-203; RTX 4090 rows in BASELINE.md): eval + no_grad, cudnn.benchmark, ONE random tensor fed as both views,
20 warm-up + ``num_iterations`` timed forwards bracketed by CUDA events, peak allocated memory per forward, at
(480,640), (736,1280), (1088,1920) (+ an optional extra resolution).  Measurement harness, not part of the hot path.
"""
from __future__ import annotations

import torch

RESOLUTIONS = ((480, 640), (736, 1280), (1088, 1920))
WARMUP = 20


def speed_and_memory_test(model, resolution=None, batch_size=1, num_iterations=100, device="cuda:0", verbose=True):
    """Returns (resolutions, mean seconds per forward, mean peak MB per forward) like the reference."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise ValueError("speed_and_memory_test times CUDA events: a CUDA device is required")
    torch.backends.cudnn.benchmark = True                    # speed_and_memory_test.py:8
    model = model.to(dev).eval()
    if verbose:
        n_all = sum(p.numel() for p in model.parameters())
        n_train = sum(p.numel() for p in model.parameters() if p.requires_grad)
        print(f"Total number of parameters: {n_all / 1e6:.2f}M")
        print(f"Learnable parameters: {n_train / 1e6:.2f}M")
    todo = list(RESOLUTIONS) + ([tuple(resolution)] if resolution is not None else [])
    mean_s, mean_mb = [], []
    with torch.no_grad():
        for res in todo:
            view = torch.randn(batch_size, 3, *res, device=dev)       # the reference feeds the same tensor twice (:44-45)
            for _ in range(WARMUP):
                model(view, view)
            secs, mbs = 0.0, 0.0
            for _ in range(num_iterations):
                torch.cuda.reset_peak_memory_stats(dev)
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                model(view, view)
                t1.record()
                torch.cuda.synchronize(dev)
                secs += t0.elapsed_time(t1) * 1e-3
                mbs += torch.cuda.max_memory_allocated(dev) / 2 ** 20
            mean_s.append(secs / num_iterations)
            mean_mb.append(mbs / num_iterations)
            if verbose:
                print(f"Resolution: {res}, Avg Time: {mean_s[-1]:.4f} s, Avg Frequency: {1 / mean_s[-1]:.4f} Hz,"
                      f"Avg Memory: {mean_mb[-1]:.2f} MB")
    return todo, mean_s, mean_mb
