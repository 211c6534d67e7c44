import sys, torch, traceback
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from conftest import load_golden, golden_state
import stereo_toolbox_b200 as S
from stereo_toolbox_b200.synth import synth_pair
for key, ctor, seed in (("cfnet", S.CFNet, 6), ("pcwnet_gc", S.PCWNet_GC, 7)):
    g = load_golden(f"{key}.npz"); sd, meta = golden_state(key)
    left, right = synth_pair(1, 64, 128, seed=seed, shift=meta["shift"])
    for prec in ("fp32", "fp16x2"):
        try:
            net = ctor(meta["maxdisp"], precision=prec); net.load_state_dict(sd, strict=True); net = net.cuda().eval()
            torch.backends.cudnn.allow_tf32 = False
            with torch.no_grad():
                out = net(left.cuda(), right.cuda())
            out = out[-1] if isinstance(out, (list, tuple)) else out
            d = (out.cpu().reshape(g["disp"].shape) - g["disp"]).abs()
            print(key, prec, "EPE", d.mean().item(), "median", d.median().item())
        except Exception as e:
            print(key, prec, "FAILED:", repr(e)[:300]); traceback.print_exc(limit=4)
