"""Localise errors of the split (fp16x2) 2-D extractor: each layer vs torch fp32 on the SAME input (one trip)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import golden_state
import stereo_toolbox_b200 as S
from stereo_toolbox_b200.features_umma import UmmaGwcFeatures
from stereo_toolbox_b200.aggregation_umma import to_channels_last, from_channels_last, UmmaBackend
from stereo_toolbox_b200.synth import synth_pair
import torch.nn.functional as F

torch.backends.cudnn.allow_tf32 = False
sd, meta = golden_state("gwcnet_gc")
net = S.GwcNet_GC(32); net.load_state_dict(sd); net = net.cuda().eval()
fe = net.feature_extraction
U = UmmaGwcFeatures("fp16x2")

def tocl(x):   # [N,C,H,W] -> [1,N,H,W,2C]
    c = x.shape[1]; cp = 16 if c < 16 else c
    return to_channels_last(x, cp, torch.float16, split=True).view(1, x.shape[0], x.shape[2], x.shape[3], 2 * cp)
def fromcl(y, c):
    _, N, H, W, C2 = y.shape
    return from_channels_last(y.view(N, H, W, C2), c, split=True)
def chk(name, got, want):
    e = (got - want).abs()
    print(f"{name:38s} max {e.max().item():.3e} mean {e.mean().item():.3e} |want| {want.abs().mean().item():.3f}", flush=True)

with torch.no_grad():
    left, right = synth_pair(2, 64, 160, seed=7, shift=6)
    x = torch.cat((left, right), 0).cuda()
    def layer(name, conv, bn, xin, act="none", res=None):
        y = conv(xin); 
        if bn is not None: y = bn(y)
        if res is not None: y = y + res
        if act == "relu": y = F.relu(y)
        got = U.conv(conv, bn, tocl(xin), act, None if res is None else tocl(res))
        chk(name + f" {conv.in_channels}->{conv.out_channels} k{conv.kernel_size[0]} s{conv.stride[0]} d{conv.dilation[0]}", fromcl(got, conv.out_channels), y)
        return y
    fc = fe.firstconv
    y = layer("firstconv0", fc[0][0], fc[0][1], x, "relu")
    y = layer("firstconv2", fc[2][0], fc[2][1], y, "relu")
    y = layer("firstconv4", fc[4][0], fc[4][1], y, "relu")
    def block(name, blk, xin):
        a = layer(name + ".conv1", blk.conv1[0][0], blk.conv1[0][1], xin, "relu")
        sh = xin if blk.downsample is None else layer(name + ".down", blk.downsample[0], blk.downsample[1], xin)
        return layer(name + ".conv2", blk.conv2[0], blk.conv2[1], a, "none", sh)
    for i, blk in enumerate(fe.layer1[:1]): y = block(f"layer1.{i}", blk, y)
    l2 = y
    for i, blk in enumerate(fe.layer2[:2]): l2 = block(f"layer2.{i}", blk, l2)
    l3 = l2
    for i, blk in enumerate(fe.layer3[:2]): l3 = block(f"layer3.{i}", blk, l3)
    l4 = l3
    for i, blk in enumerate(fe.layer4[:2]): l4 = block(f"layer4.{i}", blk, l4)
    gwc = torch.cat((l2, l3, l4), 1)
    lc = fe.lastconv
    y = layer("lastconv0", lc[0][0], lc[0][1], gwc, "relu")
    cat = layer("lastconv2", lc[2], None, y)
    # whole extractor + the channels-last volume hand-over
    want_l, want_r = fe(left.cuda()), fe(right.cuda())
    got_l, got_r = U(fe, left.cuda(), right.cuda())
    for k in ("gwc_feature", "concat_feature"):
        chk("whole " + k, got_l[k], want_l[k])
    be = UmmaBackend("fp16x2")
    cl, _ = U(fe, left.cuda(), right.cuda(), channels_last_out=True)
    vol = be.volume_from_cl(cl["_cl"], 8, 40)
    vol2 = be.volume_gwc_concat(got_l["gwc_feature"], got_r["gwc_feature"], got_l["concat_feature"], got_r["concat_feature"], 8, 40)
    chk("volume from_cl vs nchw", from_channels_last(vol, 64, split=True), from_channels_last(vol2, 64, split=True))
