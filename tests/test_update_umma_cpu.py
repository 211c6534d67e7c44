"""Host logic of update_umma.UmmaRaftUpdate (fused z|r weights, convolution over a concatenation as a residual chain with
per-piece input-channel ranges, motion-encoder padding, level wiring of BasicMultiUpdateBlock) pinned on CPU: the kernel-backed
primitives are replaced by torch stand-ins with the SAME call contract and the result is compared with the torch update block
(the reference-named module, itself pinned to the reference by tests/test_host_mirror_cpu.py::test_raft_stereo_mirror).
The kernels behind the primitives are pinned by tests/test_update_umma_gpu.py."""
import pytest
import torch
import torch.nn.functional as F

import stereo_toolbox_b200.raft_stereo as rs
from stereo_toolbox_b200.update_umma import UmmaIgevUpdate, UmmaRaftUpdate


class _TorchFx:
    """features_umma.UmmaGwcFeatures.conv by torch: x [1,N,H,W,Cpad] fp64 'kernel layout' stand-in."""

    def __init__(self):
        self._plans = {}

    def conv(self, conv, bn, x, act="none", residual=None, cin_range=None, with_shift=True):
        assert bn is None
        w = conv.weight.detach().double()
        if cin_range is not None:
            w = w[:, cin_range[0]:cin_range[1]]
        ct = x.shape[-1]
        assert w.shape[1] <= ct
        w = F.pad(w, (0, 0, 0, 0, 0, ct - w.shape[1]))                      # zero weights for the tensor's padding channels
        y = F.conv2d(x[0].permute(0, 3, 1, 2), w, None, conv.stride, conv.padding, conv.dilation)
        cpad = (y.shape[1] + 15) // 16 * 16
        y = F.pad(y, (0, 0, 0, 0, 0, cpad - y.shape[1]))
        if with_shift and conv.bias is not None:
            y[:, :conv.bias.numel()] += conv.bias.detach().double().view(1, -1, 1, 1)
        y = y.permute(0, 2, 3, 1)[None]
        if residual is not None:
            assert residual.shape == y.shape
            y = y + residual
        return {"none": lambda t: t, "relu": torch.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh}[act](y)


class _CpuMixin:
    def __init__(self, ub, args):
        self.ub, self.args = ub, args
        self.fx = _TorchFx()
        self._v, self._ver = {}, None

    def to_cl(self, x, cpad=None):
        cpad = cpad or (x.shape[1] + 15) // 16 * 16
        return F.pad(x.double(), (0, 0, 0, 0, 0, cpad - x.shape[1])).permute(0, 2, 3, 1)[None]

    def from_cl(self, x, c=None):
        y = x[0].permute(0, 3, 1, 2)
        return (y if c is None else y[:, :c]).float()

    def _rh(self, zr, h):
        C = h.shape[-1]
        return zr[..., C:2 * C] * h

    def _blend(self, zr, h, q):
        z = zr[..., :h.shape[-1]]
        return (1 - z) * h + z * q

    def pool2x(self, x):
        return F.avg_pool2d(x[0].permute(0, 3, 1, 2), 3, stride=2, padding=1).permute(0, 2, 3, 1)[None]

    def interp(self, x, dest):
        y = F.interpolate(x[0].permute(0, 3, 1, 2), dest.shape[2:4], mode="bilinear", align_corners=True)
        return y.permute(0, 2, 3, 1)[None]


class _CpuUpdate(_CpuMixin, UmmaRaftUpdate):
    pass


class _CpuIgevUpdate(_CpuMixin, UmmaIgevUpdate):
    pass


def test_igev_update_step_matches_torch_update_block():
    import stereo_toolbox_b200.igev_stereo as ig
    torch.manual_seed(0)
    net = ig.IGEVStereo().eval()
    ub, a = net.update_block, net.args
    N, h, w = 1, 12, 20
    sizes = [(h, w), (6, 10), (3, 5)]
    g = torch.Generator().manual_seed(1)
    net_list = [torch.tanh(torch.randn(N, 128, *sizes[i], generator=g)) for i in range(3)]
    inp_list = [[0.5 * torch.randn(N, 128, *sizes[i], generator=g) for _ in range(3)] for i in range(3)]
    corr = torch.randn(N, a.corr_levels * (2 * a.corr_radius + 1) * 9, h, w, generator=g)
    disp = 5.0 * torch.rand(N, 1, h, w, generator=g)
    with torch.no_grad():
        want_net, want_mask, want_delta = ub([t.clone() for t in net_list], inp_list, corr, disp)
    upd = _CpuIgevUpdate(ub, a)
    with torch.no_grad():
        upd._prepare()
        got_net, got_delta = _run_with_doubling(upd, [upd.to_cl(t) for t in net_list],
                                                [(upd.to_cl(torch.cat((cz, cr), 1)), upd.to_cl(cq)) for cz, cr, cq in inp_list],
                                                corr, disp)
    for gt, wt in zip(got_net, want_net):
        assert (upd.from_cl(gt) - wt).abs().max().item() < 1e-5
    assert got_delta.shape == want_delta.shape == (N, 1, h, w)
    assert (got_delta - want_delta).abs().max().item() < 1e-5
    assert (upd.mask(got_net[0]) - want_mask).abs().max().item() < 1e-4


@pytest.mark.parametrize("n_gru_layers", [3, 2, 1])
def test_update_step_matches_torch_update_block(n_gru_layers):
    import argparse
    torch.manual_seed(0)
    net = rs.RAFTStereo(argparse.Namespace(n_gru_layers=n_gru_layers)).eval()
    ub, a = net.update_block, net.args
    N, h, w = 2, 12, 20
    sizes = [(h, w), ((h - 1) // 2 + 1, (w - 1) // 2 + 1)]
    sizes.append(((sizes[1][0] - 1) // 2 + 1, (sizes[1][1] - 1) // 2 + 1))
    g = torch.Generator().manual_seed(1)
    net_list = [torch.tanh(torch.randn(N, 128, *sizes[i], generator=g)) for i in range(n_gru_layers)]
    inp_list = [[0.5 * torch.randn(N, 128, *sizes[i], generator=g) for _ in range(3)] for i in range(n_gru_layers)]
    corr = torch.randn(N, a.corr_levels * (2 * a.corr_radius + 1), h, w, generator=g)
    flow = torch.randn(N, 2, h, w, generator=g)
    flow[:, 1] = 0
    with torch.no_grad():
        want_net, want_mask, want_delta = ub([t.clone() for t in net_list], inp_list, corr, flow,
                                             iter32=n_gru_layers == 3, iter16=n_gru_layers >= 2)
    upd = _CpuUpdate(ub, a)
    with torch.no_grad():
        upd._prepare()
        got_net, got_delta = _run_with_doubling(upd, [upd.to_cl(t) for t in net_list],
                                                [(upd.to_cl(torch.cat((cz, cr), 1)), upd.to_cl(cq)) for cz, cr, cq in inp_list],
                                                corr, flow)
    for gt, wt in zip(got_net, want_net):
        assert (upd.from_cl(gt) - wt).abs().max().item() < 1e-5
    assert (got_delta - want_delta).abs().max().item() < 1e-5
    assert (upd.mask(got_net[0]) - want_mask).abs().max().item() < 1e-4


def _run_with_doubling(upd, net, ctx, corr, flow):
    """UmmaRaftUpdate.step on the stand-ins.  The product code computes C = t.shape[-1] // 2 (split storage: two halves per
    channel); the stand-in tensors carry one element per channel, so they are fed with an ignored zero second half."""
    dbl = lambda t: torch.cat((t, torch.zeros_like(t)), -1)
    half = lambda t: t[..., : t.shape[-1] // 2]

    class FxD:
        def __init__(self, inner):
            self.inner, self._plans = inner, inner._plans

        def conv(self, conv, bn, x, act="none", residual=None, cin_range=None, with_shift=True):
            y = self.inner.conv(conv, bn, half(x), act, None if residual is None else half(residual), cin_range, with_shift)
            return dbl(y)

    upd.fx = FxD(upd.fx)
    upd_to_cl, upd_from_cl = upd.to_cl, upd.from_cl
    upd.to_cl = lambda x, cpad=None: dbl(upd_to_cl(x, cpad))
    upd.from_cl = lambda x, c=None: upd_from_cl(half(x), c)
    rh, blend, pool, interp = upd._rh, upd._blend, upd.pool2x, upd.interp
    upd._rh = lambda zr, h: dbl(rh(half(zr), half(h)))
    upd._blend = lambda zr, h, q: dbl(blend(half(zr), half(h), half(q)))
    upd.pool2x = lambda x: dbl(pool(half(x)))
    upd.interp = lambda x, dest: dbl(interp(half(x), dest))
    net2, delta = upd.step([dbl(t) for t in net], [(dbl(a), dbl(b)) for a, b in ctx], corr, flow)
    upd.to_cl, upd.from_cl = upd_to_cl, upd_from_cl
    return [half(t) for t in net2], delta
