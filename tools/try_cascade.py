import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from conftest import load_golden, golden_state
import stereo_toolbox_b200 as S
from stereo_toolbox_b200.synth import synth_pair
g = load_golden("cfnet.npz"); sd, meta = golden_state("cfnet")
left, right = synth_pair(1, 64, 128, seed=6, shift=meta["shift"])
res = {}
for prec in ("fp32", "fp16x2", "fp16"):
    net = S.CFNet(meta["maxdisp"], precision=prec); net.load_state_dict(sd, strict=True); net = net.cuda().eval()
    torch.backends.cudnn.allow_tf32 = False
    with torch.no_grad():
        out = net(left.cuda(), right.cuda())
    out = out[-1] if isinstance(out, (list, tuple)) else out
    res[prec] = (out.cpu(), {k: v.cpu() for k, v in net._last.items()})
    d = (out.cpu().reshape(g["disp"].shape) - g["disp"]).abs()
    print(prec, "final EPE", d.mean().item(), "median", d.median().item(), "frac>0.01", (d > 0.01).float().mean().item())
for prec in ("fp16x2", "fp16"):
    for k in res["fp32"][1]:
        d = (res[prec][1][k] - res["fp32"][1][k]).abs()
        print(prec, k, "vs fp32 path: mean", d.mean().item(), "median", d.median().item(), "max", d.max().item())
