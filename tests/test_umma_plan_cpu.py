"""Host-side tap tables of the tcgen05 convolution (aggregation_umma.UmmaPlan) checked WITHOUT a GPU: a few lines of
numpy-style torch re-state what csrc/conv3d_umma.cu does with a plan (plane / chunk / sub-tile addressing, merged taps,
parity-class blocks, K-chunks along the pseudo-depth axis) and the result is compared with torch's own convolution in
fp64.  This is test infrastructure (it never runs on the product path): it pins the PLAN, the -m gpu tests pin the kernel.
"""
import pytest
import torch
import torch.nn as nn

from stereo_toolbox_b200.aggregation_umma import UmmaPlan, split_pack, split_unpack


def emulate(plan: UmmaPlan, x: torch.Tensor) -> torch.Tensor:
    """x [B,Cin,D,H,W] fp64 -> conv output [B,Cout,Do,Ho,Wo] fp64 following the plan's tables the way the kernel does."""
    B, Cin, Di, Hi, Wi = x.shape
    split = plan.split
    cst = 2 if split else 1
    # storage view of the input: channels-last, (hi, lo) interleaved per 16 channels when split
    xcl = x.permute(0, 2, 3, 4, 1).float()
    xs = (split_pack(xcl) if split else xcl.half()).double()                     # [B,D,H,W,Cst]
    Do, Ho, Wo = plan.out_size(Di), plan.out_size(Hi), plan.out_size(Wi)
    nk, kc = plan.nk, plan.kc
    kdepth = getattr(plan, "kdepth", False)
    G = (getattr(plan, "kgroup", 0) or nk) if kdepth else 1     # chunks per pass of the pseudo-depth form (kgroup: fewer than nk)
    npass = nk // G if kdepth else nk
    wt = plan.wt.double()                  # [tile][nk][cpad][kc]  or (kdepth)  [pass][chunk in pass][tile][cpad][kc] flattened below
    if kdepth:
        wt = wt.reshape(npass * plan.nwtiles, 1, plan.cpad, kc)
    ntaps = plan.ntaps
    dz, dh, dw, sub, widx = (list(plan.c_dz), list(plan.c_dh), list(plan.c_dw), list(plan.c_sub), list(plan.c_widx))
    merge = 3 if plan.merge else 1
    lane_merged = bool(plan.merge) or bool(getattr(plan, "pair_merge", False))   # column block j is realigned by j lanes -> block 0
    nblk = list(plan.c_nblk) if plan.c_nblk is not None else [merge] * ntaps
    cls0 = list(plan.c_cls0) if plan.c_cls0 is not None else [0] * ntaps
    tb, te = list(plan.c_tb), list(plan.c_te)
    od0, oh0, ow0 = list(plan.c_od0), list(plan.c_oh0), list(plan.c_ow0)
    in_s, out_s = plan.in_stride, plan.out_stride
    nsteps, nh, nw = (Di, Hi, Wi) if plan.tr else (Do, Ho, Wo)
    sd_in = G if kdepth else in_s
    out = torch.zeros(B, Do, Ho, Wo, plan.cpad, dtype=torch.float64)
    cb = 8 if plan.deconv_merge else (2 if getattr(plan, "pair_merge", False) else merge)

    def gather(P, chunk_k, hh, ww):
        """rows of the A operand: input storage elements [B, nh, nw, kc] at plane P, positions (hh[jh], ww[jw])."""
        if kdepth:
            depth, chunk = P // G, chunk_k * G + P % G        # floor division; chunk_k = the pass
        else:
            depth, chunk = P, chunk_k
        a = torch.zeros(B, nh, nw, kc, dtype=torch.float64)
        if depth < 0 or depth >= Di:
            return a
        hv = (hh >= 0) & (hh < Hi)
        wv = (ww >= 0) & (ww < Wi)
        src = xs[:, depth][:, hh.clamp(0, Hi - 1)][:, :, ww.clamp(0, Wi - 1)][..., chunk * kc:(chunk + 1) * kc]
        return src * (hv.view(1, -1, 1, 1) & wv.view(1, 1, -1, 1))

    jh, jw = torch.arange(nh), torch.arange(nw)
    for s in range(nsteps):
        for c in range(len(tb)):
            acc = torch.zeros(B, nh, nw, cb, plan.cpad, dtype=torch.float64)
            for kp in range(npass):
                for t in range(tb[c], te[c]):
                    P = s * sd_in + dz[t]
                    for j in range(nblk[t]):
                        shift_w = j if lane_merged else 0             # merged tiles: column block j is realigned by j lanes
                        hh = (jh + plan.in_off) * in_s + (sub[t] >> 1) + in_s * dh[t]
                        ww = (jw + plan.in_off) * in_s + (sub[t] & 1) + in_s * (dw[t] + shift_w)
                        a = gather(P, kp, hh, ww)
                        w = wt[kp * plan.nwtiles + widx[t] + j, 0] if kdepth else wt[widx[t] + j, kp]       # [cpad][kc]
                        # (split: hi*hi + hi*lo + lo*hi = the product of the joined values minus lo*lo, below fp32 rounding)
                        blk = 0 if lane_merged else cls0[t] + j
                        contrib = torch.einsum("bhwk,ok->bhwo", _join(a, split), _join(w, split))
                        acc[:, :, :, blk] += contrib
            order = getattr(plan, "class_order", list(range(8)))      # column block -> output-parity class (Gray order: flags bit14)
            for blk in range(cb if plan.deconv_merge else 1):
                pc = order[blk]
                cd, ch, cw = ((pc >> 2, (pc >> 1) & 1, pc & 1) if plan.deconv_merge else (od0[c], oh0[c], ow0[c]))
                od = s * out_s + cd
                if od >= Do:
                    continue
                oh, ow = jh * out_s + ch, jw * out_s + cw
                okh, okw = oh < Ho, ow < Wo
                out[:, od][:, oh[okh][:, None], ow[okw][None, :]] = acc[:, :, :, blk][:, okh][:, :, okw]
    out = out * (2.0 ** -plan.wexp)
    if plan.shift is not None:
        out[..., :plan.cout] += plan.shift.double()
    return out[..., :plan.cout].permute(0, 4, 1, 2, 3)


def _join(t, split):
    """storage elements along the last axis -> logical values (hi + lo) when split."""
    if not split:
        return t
    c2 = t.shape[-1]
    v = t.reshape(*t.shape[:-1], c2 // 32, 2, 16)
    return (v[..., 0, :] + v[..., 1, :]).reshape(*t.shape[:-1], c2 // 2)


CASES = [
    # cin, cout, k, stride, pad, transposed, split, (D,H,W), expect_kdepth
    (32, 32, 3, 1, 1, False, True, (3, 5, 6), False),
    (64, 32, 3, 1, 1, False, True, (3, 4, 5), False),        # K-split passes
    (64, 64, 1, 1, 0, False, True, (2, 4, 5), True),         # k1: both K-chunks in TMEM
    (64, 32, 3, 2, 1, True, True, (2, 3, 4), True),          # merged transposed conv, chunks along the pseudo-depth axis
    (128, 64, 3, 2, 1, True, True, (2, 2, 3), True),         # ... 4 chunks as two passes of two (kgroup)
    (32, 64, 3, 2, 1, False, True, (4, 6, 6), False),        # strided conv: parity sub-tiles, kw 0 / 2 pair-merged
    (64, 128, 3, 2, 1, False, True, (5, 7, 9), False),
    (32, 64, 3, 2, 1, False, False, (4, 6, 8), False),
    (48, 32, 4, 2, 1, True, True, (2, 3, 3), False),         # IGEV: k4 s2 p1 transposed conv from the 48-channel level (3 K-chunks of 32)
    (48, 48, 3, 1, 1, False, True, (2, 3, 4), False),
    (16, 8, 4, 2, 1, True, True, (2, 2, 3), False),
    (64, 32, 3, 2, 1, True, False, (2, 3, 4), False),        # single fp16
    (32, 32, 3, 1, 1, False, False, (3, 4, 5), False),
]


@pytest.mark.parametrize("cin,cout,k,stride,pad,tr,split,dims,want_kdepth", CASES)
def test_plan_tables_reproduce_the_convolution(cin, cout, k, stride, pad, tr, split, dims, want_kdepth):
    g = torch.Generator().manual_seed(cin * 7 + cout + k + stride)
    D, H, W = dims
    opad = 1 if (tr and k == 3) else 0
    conv = (nn.ConvTranspose3d(cin, cout, k, stride=stride, padding=pad, output_padding=opad, bias=False) if tr
            else nn.Conv3d(cin, cout, k, stride, pad, bias=False))
    bn = nn.BatchNorm3d(cout)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * (2.0 / (cin * k ** 3)) ** 0.5)
        bn.weight.copy_(0.75 + 0.5 * torch.rand(cout, generator=g)); bn.bias.copy_(0.1 * torch.randn(cout, generator=g))
        bn.running_mean.copy_(0.1 * torch.randn(cout, generator=g)); bn.running_var.copy_(0.5 + torch.rand(cout, generator=g))
    bn.eval()
    plan = UmmaPlan(conv, bn, cin, torch.float16, split)
    assert plan.umma_ok
    assert bool(getattr(plan, "kdepth", False)) == want_kdepth
    if (cin, cout, tr, split) == (128, 64, True, True):
        assert plan.kgroup == 2 and plan.nk == 4
    if tr and split and k == 3 and cin >= 64:
        # 16-channel slices: Gray order of the class blocks, 10 runs per chunk instead of 14; w-pairs stay adjacent
        assert plan.class_order == [0, 1, 3, 2, 6, 7, 5, 4] and plan.ntaps == 10 * (plan.kgroup or plan.nk)
        assert all((plan.class_order[2 * p] ^ plan.class_order[2 * p + 1]) == 1 for p in range(4))
    x = torch.randn(1, cin, D, H, W, generator=g)
    if not split:
        x = x.half().float()
    got = emulate(plan, x.double())
    with torch.no_grad():
        want = bn.double()(conv.double()(x.double()))
    scale = max(1.0, want.abs().max().item())
    tol = 4e-6 if split else 2e-3            # single fp16: the weights are rounded to 11 bits
    assert got.shape == want.shape
    assert (got - want).abs().max().item() < tol * scale


# ----------------------------------------------------------------------------------------------- 2-D plans (features_umma.Conv2dPlan)
def emulate2d(plan, x: torch.Tensor) -> torch.Tensor:
    """x [N,Cin,H,W] fp64 -> [N,Cout,Ho,Wo] fp64 following a Conv2dPlan's tables the way csrc/conv3d_umma.cu reads them: dz = 0
    everywhere, or (K-chunks along the pseudo-depth axis) dz = the chunk a tap reads; kw-merged taps carry three column
    blocks realigned by ``dil`` lanes each; stride-2 taps read (h, w)-parity sub-tiles."""
    N, Cin, H, W = x.shape
    split = plan.flags & 64
    xcl = x.permute(0, 2, 3, 1).float()
    ct = plan.cin
    xcl = torch.nn.functional.pad(xcl, (0, ct - Cin))
    xs = (split_pack(xcl) if split else xcl.half()).double()                 # [N,H,W,Cst]
    Ho, Wo = plan.out_size(H), plan.out_size(W)
    kc, nk, kdepth = plan.kc, plan.nk, plan.kdepth
    cpad = plan.wt.shape[-2]
    wt = plan.wt.double()                                                    # kdepth: [nk][k*k][cpad][kc] (chunk-major tiles); else [k*k][nk][cpad][kc]
    if kdepth:
        wt = wt.reshape(-1, cpad, kc)
    dz, dh, dw, sub, widx = (list(a) for a in plan.c)
    s, dil = plan.in_stride, plan.dil
    jh, jw = torch.arange(Ho), torch.arange(Wo)
    out = torch.zeros(N, Ho, Wo, cpad, dtype=torch.float64)
    for kp in range(1 if kdepth else nk):
        for t in range(plan.ntaps):
            chunk = dz[t] if kdepth else kp
            for j in range(3 if plan.merge else 1):
                hh = (jh + plan.in_off) * s + (sub[t] >> 1) + s * dh[t]
                ww = (jw + plan.in_off) * s + (sub[t] & 1) + s * (dw[t] + j * dil)
                hv, wv = (hh >= 0) & (hh < H), (ww >= 0) & (ww < W)
                a = xs[:, hh.clamp(0, H - 1)][:, :, ww.clamp(0, W - 1)][..., chunk * kc:(chunk + 1) * kc]
                a = a * (hv.view(1, -1, 1, 1) & wv.view(1, 1, -1, 1))
                w = wt[widx[t] + j] if kdepth else wt[widx[t] + j, kp]
                out += torch.einsum("bhwk,ok->bhwo", _join(a, split), _join(w, split))
    wexp = (plan.flags >> 16) & 127
    out = out * (2.0 ** -wexp)
    if plan.shift is not None:
        out[..., :plan.cout] += plan.shift.double()
    return out[..., :plan.cout].permute(0, 3, 1, 2)


@pytest.mark.parametrize("cin,cout,k,stride,dil,split,bias,hw", [
    (32, 32, 3, 1, 1, True, False, (9, 11)),         # kw-merged
    (128, 128, 3, 1, 2, True, False, (10, 9)),       # dilated, four K-chunks in TMEM, merged taps 2 lanes apart
    (32, 64, 3, 2, 1, True, False, (10, 12)),        # stride 2: parity sub-tiles
    (32, 64, 1, 2, 1, True, False, (9, 11)),         # 1x1 stride-2 shortcut
    (16, 64, 7, 1, 1, True, True, (8, 9)),           # 7x7 with bias (the update block's flow conv)
    (64, 64, 3, 1, 1, False, False, (7, 8)),         # single fp16
])
def test_conv2d_plan_tables_reproduce_the_convolution(cin, cout, k, stride, dil, split, bias, hw):
    from stereo_toolbox_b200.features_umma import Conv2dPlan
    g = torch.Generator().manual_seed(cin + cout + k + stride + dil)
    pad = dil if k == 3 else (k // 2 if k > 1 else 0)
    conv = nn.Conv2d(cin, cout, k, stride, pad, dil, bias=bias)
    bn = None if bias else nn.BatchNorm2d(cout)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * (2.0 / (cin * k * k)) ** 0.5)
        if bias:
            conv.bias.copy_(0.1 * torch.randn(cout, generator=g))
        else:
            bn.weight.copy_(0.75 + 0.5 * torch.rand(cout, generator=g)); bn.bias.copy_(0.1 * torch.randn(cout, generator=g))
            bn.running_mean.copy_(0.1 * torch.randn(cout, generator=g)); bn.running_var.copy_(0.5 + torch.rand(cout, generator=g))
            bn.eval()
    plan = Conv2dPlan(conv, bn, cin, torch.float16, split)
    x = torch.randn(2, cin, *hw, generator=g)
    if not split:
        x = x.half().float()
    got = emulate2d(plan, x.double())
    with torch.no_grad():
        want = conv.double()(x.double())
        if bn is not None:
            want = bn.double()(want)
    scale = max(1.0, want.abs().max().item())
    assert got.shape == want.shape
    assert (got - want).abs().max().item() < (4e-6 if split else 2e-3) * scale
