#!/usr/bin/env python
"""PSMNet's SPP extractor on the split tcgen05 kernel vs torch fp32, stage by stage (relative error of each tensor)."""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import golden_state
import stereo_toolbox_b200 as S
from stereo_toolbox_b200.synth import synth_pair
from stereo_toolbox_b200.features_umma import UmmaGwcFeatures
from stereo_toolbox_b200.aggregation_umma import from_channels_last

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
sd, _ = golden_state("psmnet")
H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (576, 960)
left, right = synth_pair(1, H, W, seed=1, shift=23)
net = S.PSMNet(192, precision="fp16x2"); net.load_state_dict(sd, strict=True); net = net.cuda().eval()
fe = net.feature_extraction
rel = lambda a, b: ((a - b).abs().max() / b.abs().max()).item()
with torch.no_grad():
    x = torch.cat((left, right), 0).cuda()
    t = fe.layer1(fe.firstconv(x)); l2 = fe.layer2(t); l3 = fe.layer3(l2); l4 = fe.layer4(l3)
    want = fe(x)
    ux = UmmaGwcFeatures("fp16x2")
    u2, u3, u4 = ux._trunk(fe, left.cuda(), right.cuda())
    cl = lambda t: from_channels_last(t.view(t.shape[1], t.shape[2], t.shape[3], t.shape[4]), split=True)
    print("l2 rel err", rel(cl(u2), l2), " l3", rel(cl(u3), l3), " l4", rel(cl(u4), l4))
    gl, gr = ux.psm(fe, left.cuda(), right.cuda())
    got = torch.cat((gl, gr), 0)
    print("final feature rel err", rel(got, want), "mean abs", (got - want).abs().mean().item(), "scale", want.abs().max().item())
    # branches from the exact l4 vs from the umma l4
    for i in (4, 3, 2, 1):
        b_t = F.interpolate(getattr(fe, f"branch{i}")(l4), l4.shape[2:], mode="bilinear", align_corners=False)
        b_u = F.interpolate(getattr(fe, f"branch{i}")(cl(u4)), l4.shape[2:], mode="bilinear", align_corners=False)
        print(f"branch{i} rel err", rel(b_u, b_t))
    # lastconv on exact inputs through the umma chain
    to = lambda t: __import__("stereo_toolbox_b200.aggregation_umma", fromlist=["x"]).to_channels_last(t, t.shape[1], torch.float16, split=True).view(1, t.shape[0], t.shape[2], t.shape[3], 2 * t.shape[1])
    br = torch.cat([F.interpolate(getattr(fe, f"branch{i}")(l4), l4.shape[2:], mode="bilinear", align_corners=False) for i in (4, 3, 2, 1)], 1)
    y = ux._convbn_multi(fe.lastconv[0], (to(l2), to(l4), to(br)), "relu")
    y_t = fe.lastconv[1](fe.lastconv[0](torch.cat((l2, l4, br), 1)))
    print("lastconv[0] (320->128 as 3 chained addends) on exact inputs: rel err", rel(cl(y), y_t))
    z = ux.conv(fe.lastconv[2], None, y)
    print("lastconv[2] rel err", rel(cl(z)[:, :32], fe.lastconv[2](y_t)))
