"""Storage model of the 16-bit tensor-core path, inside the CPU oracle.

TEST INFRASTRUCTURE -- see oracle/__init__.py for who may import this.

The tcgen05 convolution path keeps the reference's ARITHMETIC (every product and sum of the Conv3d / ConvTranspose3d +
BatchNorm3d + activation chain, accumulated in fp32) but stores operands in 16 bits.  ``storage_16bit(dtype)`` makes the
oracle round exactly where that path rounds, and nowhere else:

* BatchNorm scale folded into the conv weight, THEN rounded to 16 bits (ops.ConvPlan / aggregation_umma.py);
* the conv input and the residual are read as 16-bit values (the cost volume is therefore rounded once, when it is
  written channels-last);
* accumulation, BatchNorm shift, residual add and activation in fp32; the result is rounded to 16 bits on store --
  except the 1-channel classifier output (the pre-softmax cost), which stays fp32 all the way into the head.

Used by the tests to separate "error of the 16-bit storage format" (a property of the chosen precision: what this model
predicts) from "error of the kernel" (whatever the CUDA path differs from this model by).  Measured with the bench
weights at the full KITTI shape (1x3x384x1248, D=192): model 0.1159 px EPE vs the fp32 reference, CUDA path 0.1165 px.
"""
from __future__ import annotations

import contextlib

import torch
import torch.nn.functional as F

from . import ref_ops as R


def _round(dtype: torch.dtype, split: bool):
    """Storage rounding: one 16-bit value, or (``split``) a hi + lo pair of 16-bit values -- x ~= hi + lo with
    hi = rn16(x), lo = rn16(x - hi): ~22 mantissa bits for fp16 (profiles/next_round_plan.md section 6)."""
    if not split:
        return lambda t: t.to(dtype).float()

    def r(t):
        hi = t.to(dtype).float()
        return hi + (t - hi).to(dtype).float()
    return r


def _conv_16bit(dtype: torch.dtype, round_weights: bool = True, round_activations: bool = True, split: bool = False):
    r = _round(dtype, split)

    def conv3d_bn_act(x, weight, bn=None, stride=1, padding=1, act="none", residual=None, transposed=False,
                      output_padding=0):
        c_out = weight.shape[1] if transposed else weight.shape[0]
        scale, shift = R.fold_bn(bn, c_out)
        w = weight * (scale.view(1, -1, 1, 1, 1) if transposed else scale.view(-1, 1, 1, 1, 1))
        if round_weights:
            w = r(w)
        if round_activations:
            x = r(x)
        if transposed:
            y = F.conv_transpose3d(x, w, stride=stride, padding=padding, output_padding=output_padding)
        else:
            y = F.conv3d(x, w, stride=stride, padding=padding)
        y = y + shift.view(1, -1, 1, 1, 1)
        if residual is not None:
            y = y + (r(residual) if round_activations else residual)
        y = R.activation(y, act)
        return y if (c_out == 1 or not round_activations) else r(y)

    return conv3d_bn_act


@contextlib.contextmanager
def storage_16bit(dtype: torch.dtype = torch.float16, round_weights: bool = True, round_activations: bool = True,
                  split: bool = False):
    """Inside the block, every ``ref_ops.conv3d_bn_act`` call (hence every 3-D layer of oracle/ref_models.py) rounds its
    operands / result to ``dtype`` as described in the module docstring."""
    orig = R.conv3d_bn_act
    R.conv3d_bn_act = _conv_16bit(dtype, round_weights, round_activations, split)
    try:
        yield
    finally:
        R.conv3d_bn_act = orig
