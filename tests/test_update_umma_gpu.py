"""RAFT-Stereo's update block on the tensor-core 2-D conv path (update_umma.UmmaRaftUpdate, csrc/gru2d.cu, the sigmoid /
tanh epilogues of csrc/conv3d_umma.cu) against torch's fp32 arithmetic of the same reference-named modules, and the whole
model against the REFERENCE's output (fixture).  Bar: the fp32 bar (<= 1e-3 px), the format is 'fp16x2'."""
import argparse

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from conftest import load_golden, golden_state

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def rnd(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def _upd():
    import stereo_toolbox_b200.raft_stereo as rs
    from stereo_toolbox_b200.update_umma import UmmaRaftUpdate
    torch.manual_seed(0)
    net = rs.RAFTStereo().cuda().eval()
    return net, UmmaRaftUpdate(net.update_block, net.args)


def test_gate_pool_interp_kernels():
    _, upd = _upd()
    N, H, W, C = 2, 13, 21, 128
    h, q = rnd(1, N, C, H, W), rnd(2, N, C, H, W)
    zr = torch.sigmoid(rnd(3, N, 2 * C, H, W))
    hc, qc, zc = upd.to_cl(h.cuda()), upd.to_cl(q.cuda()), upd.to_cl(zr.cuda())
    h2, q2, z2 = upd.from_cl(hc).cpu(), upd.from_cl(qc).cpu(), upd.from_cl(zc).cpu()        # the stored (22-bit) values
    got = upd.from_cl(upd._rh(zc, hc)).cpu()
    torch.testing.assert_close(got, z2[:, C:] * h2, rtol=1e-6, atol=1e-6)
    got = upd.from_cl(upd._blend(zc, hc, qc)).cpu()
    torch.testing.assert_close(got, (1 - z2[:, :C]) * h2 + z2[:, :C] * q2, rtol=1e-6, atol=1e-6)
    got = upd.from_cl(upd.pool2x(hc)).cpu()
    torch.testing.assert_close(got, F.avg_pool2d(h2, 3, stride=2, padding=1), rtol=1e-6, atol=1e-6)
    small = rnd(4, N, C, 7, 11)
    sc = upd.to_cl(small.cuda())
    got = upd.from_cl(upd.interp(sc, hc)).cpu()
    want = F.interpolate(upd.from_cl(sc).cpu(), (H, W), mode="bilinear", align_corners=True)
    torch.testing.assert_close(got, want, rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("act", ["sigmoid", "tanh"])
def test_conv_gate_epilogues(act):
    """bias + residual + sigmoid / tanh in the conv epilogue, convolution over a concatenation as a residual chain."""
    _, upd = _upd()
    conv = nn.Conv2d(256, 128, 3, padding=1).cuda()
    a, b = rnd(1, 2, 128, 17, 33), rnd(2, 2, 128, 17, 33)
    c = 0.5 * rnd(3, 2, 128, 17, 33)
    ac, bc, cc = upd.to_cl(a.cuda()), upd.to_cl(b.cuda()), upd.to_cl(c.cuda())
    got = upd.from_cl(upd._conv_cat(conv, [(ac, (0, 128)), (bc, (128, 256))], act, first=cc)).cpu()
    with torch.no_grad():
        want = conv.double()(torch.cat((a, b), 1).double().cuda()).cpu() + c.double()
    want = (torch.sigmoid(want) if act == "sigmoid" else torch.tanh(want)).float()
    assert (got - want).abs().max().item() < 5e-6


def test_update_step_matches_torch():
    net, upd = _upd()
    ub, a = net.update_block, net.args
    N, h, w = 1, 24, 40
    sizes = [(h, w), (12, 20), (6, 10)]
    net_list = [torch.tanh(rnd(10 + i, N, 128, *sizes[i])).cuda() for i in range(3)]
    inp_list = [[(0.5 * rnd(20 + 3 * i + j, N, 128, *sizes[i])).cuda() for j in range(3)] for i in range(3)]
    corr = rnd(5, N, 36, h, w).cuda()
    flow = rnd(6, N, 2, h, w).cuda()
    flow[:, 1] = 0
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            want_net, want_mask, want_delta = ub([t.clone() for t in net_list], inp_list, corr, flow)
            got_net, got_delta = upd.step([upd.to_cl(t) for t in net_list], upd.context(inp_list), corr, flow)
            got_mask = upd.mask(got_net[0])
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    for i, (g, wt) in enumerate(zip(got_net, want_net)):
        err = (upd.from_cl(g) - wt).abs().max().item()
        print(f"hidden state level {i}: max err {err:.3e}")
        assert err < 2e-5
    err = (got_delta - want_delta).abs().max().item()
    print(f"delta flow: max err {err:.3e} (scale {want_delta.abs().max().item():.3f})")
    assert err < 2e-5 * max(1.0, want_delta.abs().max().item())
    assert (got_mask - want_mask).abs().max().item() < 1e-4


@pytest.mark.parametrize("graph", [False, True])
def test_raft_stereo_golden_update_on_tensor_cores(graph):
    import stereo_toolbox_b200 as S
    from stereo_toolbox_b200.synth import synth_pair
    g = load_golden("raft_stereo.npz")
    sd, meta = golden_state("raft_stereo", calib=False)
    net = S.RAFTStereo()
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    assert getattr(net, "update_mode", "auto") == "auto"          # the default: tensor-core update block for CUDA inference
    net.cuda_graph = graph
    left, right = synth_pair(1, 64, 128, seed=2, shift=meta["shift"])
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False       # the 2-D encoders are torch; keep them fp32 for the comparison
    try:
        with torch.no_grad():
            out = net(left.cuda(), right.cuda(), iters=meta["iters"]).cpu()
            if graph:                             # second call: replay of the cached graph with refreshed static inputs
                out2 = net(left.cuda(), right.cuda(), iters=meta["iters"]).cpu()
                assert torch.equal(out, out2)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    epe = (out - g["disp"]).abs().mean().item()
    print(f"RAFT-Stereo, update block on tcgen05 (graph={graph}): EPE vs reference {epe:.3e} px")
    assert epe < 1e-3, f"EPE vs reference {epe}"
