"""Deterministic synthetic weights and inputs (there is no network for checkpoints or datasets).

``synth_state_dict`` fills a state-dict *template* (names -> tensors giving shape/dtype) with values
drawn from a generator seeded by the parameter NAME, so the reference model, the oracle and the
CUDA-backed model receive bit-identical weights without sharing constructor order or init code.
BatchNorm running statistics are non-trivial on purpose (exercises the BN fold), conv weights are
He-scaled so activations neither vanish nor explode through ~25 stacked 3-D convs.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Mapping

import torch


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    return g


def synth_state_dict(template: Mapping[str, torch.Tensor], seed: int = 0,
                     overrides: Mapping[str, object] | None = None) -> Dict[str, torch.Tensor]:
    """``overrides``: name -> array, e.g. the calibrated BatchNorm statistics in
    ``tests/golden/bn_calib_<model>.npz`` (see tests/golden/make_golden.py)."""
    out: Dict[str, torch.Tensor] = {}
    for name in template:
        ref = template[name]
        shape = tuple(ref.shape)
        g = _gen(name, seed)
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            t = torch.zeros(shape, dtype=ref.dtype)
        elif leaf == "running_mean":
            t = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "running_var":
            t = 0.5 + torch.rand(shape, generator=g)
        elif len(shape) <= 1 and leaf == "weight":
            t = 0.75 + 0.5 * torch.rand(shape, generator=g)
            if name.endswith(".conv2.1.weight"):
                # last BN of a residual BasicBlock: damp it, otherwise the 25 un-normalised
                # residual adds of the 2-D extractor double the variance per block (x 2^25)
                t = 0.25 * t
        elif len(shape) <= 1:
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = max(1, math.prod(shape[1:]))
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        out[name] = t.to(ref.dtype)
    if overrides is not None:
        for name in overrides:
            if name in out:
                out[name] = torch.as_tensor(overrides[name]).to(out[name].dtype).reshape(out[name].shape).clone()
    return out


def state_checksum(sd: Mapping[str, torch.Tensor]) -> float:
    """Order-independent fingerprint used by the golden fixtures to detect RNG drift."""
    return float(sum(v.double().abs().sum().item() for v in sd.values()))


def synth_pair(batch: int, height: int, width: int, seed: int = 0, shift: int | None = None):
    """Left/right image pair, ImageNet-normalised scale.
    ``shift=None``: independent N(0,1) images; else right = roll(left, -shift) + 0.1*noise."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1234567 + seed)
    left = torch.randn(batch, 3, height, width, generator=g)
    if shift is None:
        right = torch.randn(batch, 3, height, width, generator=g)
    else:
        right = torch.roll(left, -shift, dims=3) + 0.1 * torch.randn(batch, 3, height, width, generator=g)
    return left, right


def synth_gt(batch: int, height: int, width: int) -> torch.Tensor:
    """Deterministic synthetic ground-truth disparity in [1, 31): an RNG-free pattern so that fixtures and tests agree."""
    b = torch.arange(batch).view(-1, 1, 1)
    y = torch.arange(height).view(1, -1, 1)
    x = torch.arange(width).view(1, 1, -1)
    return 1.0 + ((x * 7 + y * 13 + b * 5) % 300).float() / 10.0
