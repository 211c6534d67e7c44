"""Differentiable forms of the hot-path operators (training path, exact fp32): ``torch.autograd.Function`` wrappers whose
forward AND backward run in libstb200.so.

  conv3d / conv_transpose3d : forward stb_conv3d_taps_f32; data gradient = the adjoint convolution of the same family
                              (Conv3d <-> ConvTranspose3d with the same weight tensor) through the same kernel; weight
                              gradient stb_conv3d_wgrad_f32            (PSMNet/submodule.py:16-19, stackhourglass.py:25-29)
  concat_volume / gwc_volume: stb_*_volume_f32 / stb_*_volume_bwd_f32  (GwcNet/submodule.py:30-63)
  upsample_softargmin       : fused head and its adjoint               (PSMNet/stackhourglass.py:139-156)

BatchNorm3d (batch statistics in train mode) and the activations are applied by the caller with ordinary torch modules,
so their autograd is torch's; what the reference trains through cuDNN conv kernels and ~6 materialised [B,192,H,W]
tensors per head goes through the kernels above.  Gradients are checked against torch autograd of the oracle
restatement in tests/test_gpu_train.py.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from .ops import _f32c, _p, _stream


def _same(vals, what):
    v = set(vals)
    if len(v) != 1:
        raise ValueError(f"{what} must be equal along D, H and W (got {tuple(vals)})")
    return vals[0]


class _Conv3dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, transposed, output_padding):
        plan = ops.ConvPlan(weight, None, stride, padding, transposed, output_padding, bias=bias)
        y = ops.conv3d_plan_apply(plan, x)
        ctx.save_for_backward(x, weight)
        ctx.cfg = (stride, padding, transposed, bias is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        stride, padding, transposed, has_bias = ctx.cfg
        gy = _f32c(gy)
        k = w.shape[2]
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            if not transposed:
                # adjoint of Conv3d(k,s,p) = ConvTranspose3d(k,s,p,op) with the SAME weight tensor [Cout,Cin,k,k,k]
                op = _same([x.shape[i] - ((gy.shape[i] - 1) * stride - 2 * padding + k) for i in (2, 3, 4)], "output_padding")
                plan = ops.ConvPlan(w, None, stride, padding, True, op)
            else:
                # adjoint of ConvTranspose3d = Conv3d with the same weight tensor [Cin,Cout,k,k,k] read as [out,in,...]
                plan = ops.ConvPlan(w, None, stride, padding, False)
            gx = ops.conv3d_plan_apply(plan, gy)
            assert gx.shape == x.shape
        if ctx.needs_input_grad[1]:
            P, Q = (gy, _f32c(x)) if transposed else (_f32c(x), gy)
            dw = torch.zeros(k, k, k, P.shape[1], Q.shape[1], device=x.device, dtype=torch.float32)
            _lib.call("stb_conv3d_wgrad_f32", _p(P), _p(Q), _p(dw), P.shape[0], P.shape[1], P.shape[2], P.shape[3],
                      P.shape[4], Q.shape[1], Q.shape[2], Q.shape[3], Q.shape[4], k, padding, stride, _stream())
            gw = dw.permute(4, 3, 0, 1, 2).contiguous()       # both layouts: weight[cq, cp, kd, kh, kw]
        if has_bias and ctx.needs_input_grad[2]:
            gb = gy.sum((0, 2, 3, 4))
        return gx, gw, gb, None, None, None, None


def conv3d(x, conv: torch.nn.Module):
    """``conv`` is an nn.Conv3d or nn.ConvTranspose3d (cubic kernel, equal strides / paddings); returns conv(x)."""
    tr = isinstance(conv, torch.nn.ConvTranspose3d)
    return _Conv3dFn.apply(x, conv.weight, conv.bias, conv.stride[0], conv.padding[0], tr,
                           conv.output_padding[0] if tr else 0)


def concat_volume_backward(gvol, shape, maxdisp, mask_left):
    B, C, H, W = shape
    gvol = _f32c(gvol)
    gl = torch.empty(B, C, H, W, device=gvol.device, dtype=torch.float32)
    gr = torch.empty_like(gl)
    _lib.call("stb_concat_volume_bwd_f32", _p(gvol), _p(gl), _p(gr), B, C, H, W, maxdisp, int(mask_left), 2 * C, 0, _stream())
    return gl, gr


def gwc_volume_backward(gvol, left, right, maxdisp, groups):
    B, C, H, W = left.shape
    gvol, left, right = _f32c(gvol), _f32c(left), _f32c(right)
    gl, gr = torch.empty_like(left), torch.empty_like(right)
    _lib.call("stb_gwc_volume_bwd_f32", _p(gvol), _p(left), _p(right), _p(gl), _p(gr), B, C, H, W, maxdisp, groups,
              groups, 0, _stream())
    return gl, gr


def upsample_softargmin_backward(gdisp, cost4, maxdisp, out_h, out_w, align_corners):
    B, D, H, W = cost4.shape
    gdisp, cost4 = _f32c(gdisp), _f32c(cost4)
    gcost = torch.zeros_like(cost4)
    _lib.call("stb_upsample_softargmin_bwd_f32", _p(cost4), _p(gdisp), _p(gcost), B, D, H, W, maxdisp, out_h, out_w,
              int(align_corners), _stream())
    return gcost


class _ConcatVolumeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, left, right, maxdisp, mask_left):
        ctx.cfg = (maxdisp, mask_left, tuple(left.shape))
        return ops.concat_volume(left, right, maxdisp, mask_left)

    @staticmethod
    def backward(ctx, gvol):
        maxdisp, mask_left, shape = ctx.cfg
        gl, gr = concat_volume_backward(gvol, shape, maxdisp, mask_left)
        return gl, gr, None, None


def concat_volume(left, right, maxdisp, mask_left=True):
    return _ConcatVolumeFn.apply(left, right, maxdisp, mask_left)


class _GwcVolumeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, left, right, maxdisp, groups):
        left, right = _f32c(left), _f32c(right)
        ctx.save_for_backward(left, right)
        ctx.cfg = (maxdisp, groups)
        return ops.gwc_volume(left, right, maxdisp, groups)

    @staticmethod
    def backward(ctx, gvol):
        left, right = ctx.saved_tensors
        maxdisp, groups = ctx.cfg
        gl, gr = gwc_volume_backward(gvol, left, right, maxdisp, groups)
        return gl, gr, None, None


def gwc_volume(left, right, maxdisp, groups):
    return _GwcVolumeFn.apply(left, right, maxdisp, groups)


class _HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cost, maxdisp, out_h, out_w, align_corners):
        if cost.dim() == 5:
            assert cost.shape[1] == 1
            ctx.squeeze = True
            cost4 = cost[:, 0]
        else:
            ctx.squeeze = False
            cost4 = cost
        cost4 = _f32c(cost4)
        ctx.save_for_backward(cost4)
        ctx.cfg = (maxdisp, out_h, out_w, align_corners)
        return ops.upsample_softargmin(cost4, maxdisp, out_h, out_w, align_corners)

    @staticmethod
    def backward(ctx, gdisp):
        (cost4,) = ctx.saved_tensors
        maxdisp, out_h, out_w, align = ctx.cfg
        gcost = upsample_softargmin_backward(gdisp, cost4, maxdisp, out_h, out_w, align)
        return (gcost.unsqueeze(1) if ctx.squeeze else gcost), None, None, None, None


def upsample_softargmin(cost, maxdisp, out_h, out_w, align_corners=False):
    """[B,1,D,h,w] or [B,D,h,w] -> [B,out_h,out_w]; differentiable w.r.t. cost."""
    return _HeadFn.apply(cost, maxdisp, out_h, out_w, align_corners)


# ------------------------------------------------------------------------------------------ CorrBlock1D (RAFT-Stereo training)
# Forward = the CUDA kernels of csrc/corr1d.cu.  The adjoints are composed from library calls for now (cuBLAS einsum for the
# all-pairs correlation, scatter-add for the lookup): they are plain dense / gather adjoints, run once per iteration, and
# are not on the inference path.  Checked against torch autograd of the oracle restatement (tests/test_raft_train_cpu.py).
class _Corr1dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fmap1, fmap2, scale):
        ctx.save_for_backward(fmap1, fmap2)
        ctx.s = (fmap1.shape[1] ** -0.5) if scale else 1.0
        return ops.corr1d(fmap1, fmap2, scale)

    @staticmethod
    def backward(ctx, g):                         # g [B,H,W1,W2]
        f1, f2 = ctx.saved_tensors
        g = g * ctx.s
        g1 = torch.einsum("bhij,bchj->bchi", g, f2.float()) if ctx.needs_input_grad[0] else None
        g2 = torch.einsum("bhij,bchi->bchj", g, f1.float()) if ctx.needs_input_grad[1] else None
        return g1, g2, None


def corr1d(fmap1, fmap2, scale=True):
    return _Corr1dFn.apply(fmap1, fmap2, scale)


class _AvgPoolLastFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.w = x.shape[-1]
        return ops.avgpool_last(x)

    @staticmethod
    def backward(ctx, g):                         # y[..., j] = (x[..., 2j] + x[..., 2j+1]) / 2; an odd last column is dropped
        gx = (0.5 * g).repeat_interleave(2, dim=-1)
        if gx.shape[-1] != ctx.w:
            gx = torch.nn.functional.pad(gx, (0, ctx.w - gx.shape[-1]))
        return gx


def avgpool_last(x):
    return _AvgPoolLastFn.apply(x)


class _Corr1dLookupFn(torch.autograd.Function):
    """out[b, l*(2r+1)+k, h, w] = linear interpolation of level l at x = coords[b,0,h,w] / 2^l + (k - r), taps outside the
    row contribute zero (RAFTStereo/corr.py:127-146).  Differentiable w.r.t. the pyramid levels only: the reference detaches
    the coordinates before every lookup (raft_stereo.py:154)."""

    @staticmethod
    def forward(ctx, coords, radius, num_levels, *levels):
        ctx.save_for_backward(coords)
        ctx.cfg = (radius, num_levels, [tuple(l.shape) for l in levels])
        return ops.corr1d_lookup(list(levels), coords, radius, num_levels)

    @staticmethod
    def backward(ctx, g):                         # g [B, L*(2r+1), H, W1]
        (coords,) = ctx.saved_tensors
        radius, num_levels, shapes = ctx.cfg
        K = 2 * radius + 1
        B, _, H, W1 = g.shape
        g = g.view(B, num_levels, K, H, W1).permute(0, 1, 3, 4, 2)            # [B,L,H,W1,K]
        dx = torch.arange(-radius, radius + 1, device=g.device, dtype=torch.float32)
        x0c = coords[:, 0].float()
        grads = []
        for l, shape in enumerate(shapes):
            if l >= num_levels or not ctx.needs_input_grad[3 + l]:
                grads.append(None)
                continue
            W2 = shape[-1]
            x = x0c[..., None] / (2 ** l) + dx                                # [B,H,W1,K]
            i0 = torch.floor(x)
            f = x - i0
            i0 = i0.long()
            gl = torch.zeros(shape, device=g.device, dtype=torch.float32)
            for idx, wgt in ((i0, 1.0 - f), (i0 + 1, f)):
                ok = (idx >= 0) & (idx < W2)
                gl.scatter_add_(3, idx.clamp(0, W2 - 1), torch.where(ok, wgt * g[:, l], torch.zeros_like(f)))
            grads.append(gl)
        return (None, None, None) + tuple(grads)


def corr1d_lookup(levels, coords, radius, num_levels):
    return _Corr1dLookupFn.apply(coords, radius, num_levels, *levels)


# ------------------------------------------------------------------------------------------ geometry encoding (IGEV training)
class _GeoLookupFn(torch.autograd.Function):
    """Combined_Geo_Encoding_Volume.__call__ (IGEVStereo/geometry.py:35-59): forward = stb_geo_lookup_f32.  Differentiable
    w.r.t. the geometry-volume and correlation pyramids; the disparity is detached before every lookup by the reference
    (igev_stereo.py:238).  Output channel order per level: C*(2r+1) geometry taps (channel-major), then 2r+1 correlation taps."""

    @staticmethod
    def forward(ctx, disp, coords, radius, n_levels, *pyr):
        geos, corrs = list(pyr[:n_levels]), list(pyr[n_levels:])
        ctx.save_for_backward(disp, coords)
        ctx.cfg = (radius, n_levels, [tuple(t.shape) for t in geos], [tuple(t.shape) for t in corrs])
        return ops.geo_lookup(geos, corrs, disp, coords, radius)

    @staticmethod
    def backward(ctx, g):
        disp, coords = ctx.saved_tensors
        radius, L, gshapes, cshapes = ctx.cfg
        K = 2 * radius + 1
        B, _, H, W = g.shape
        C = gshapes[0][3]
        g = g.permute(0, 2, 3, 1).reshape(B, H, W, L, (C + 1) * K)
        d = disp.reshape(B, H, W).float()
        xc0 = coords.reshape(B, H, W).float()
        dx = torch.arange(-radius, radius + 1, device=g.device, dtype=torch.float32)

        def scatter(shape, x, grad):                  # x, grad: [..., K] taps along the last axis of `shape`
            n = shape[-1]
            i0 = torch.floor(x)
            f = x - i0
            i0 = i0.long()
            out = torch.zeros(shape, device=grad.device, dtype=torch.float32)
            for idx, wgt in ((i0, 1.0 - f), (i0 + 1, f)):
                ok = (idx >= 0) & (idx < n)
                out.scatter_add_(out.dim() - 1, idx.clamp(0, n - 1), torch.where(ok, wgt * grad, torch.zeros_like(grad)))
            return out

        ggeo, gcorr = [], []
        for l in range(L):
            gl = g[:, :, :, l]
            x = d[..., None] / (2 ** l) + dx                                                   # [B,H,W,K]
            if ctx.needs_input_grad[4 + l]:
                ggeo.append(scatter(gshapes[l], x[..., None, :].expand(B, H, W, C, K), gl[..., :C * K].reshape(B, H, W, C, K)))
            else:
                ggeo.append(None)
            if ctx.needs_input_grad[4 + L + l]:
                gcorr.append(scatter(cshapes[l], xc0[..., None] / (2 ** l) - d[..., None] / (2 ** l) + dx, gl[..., C * K:]))
            else:
                gcorr.append(None)
        return (None, None, None, None) + tuple(ggeo) + tuple(gcorr)


def geo_lookup(geos, corrs, disp, coords, radius):
    return _GeoLookupFn.apply(disp, coords, radius, len(geos), *geos, *corrs)
