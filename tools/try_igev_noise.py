"""How far apart are two valid fp32 evaluations of IGEV-Stereo's 32-iteration loop at 1152x1920 (untrained weights)?  torch / cuDNN
update block with NCHW vs NHWC glue (different cuDNN kernels, same arithmetic) vs the tcgen05 update block."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stereo_toolbox_b200 as S
from stereo_toolbox_b200.synth import synth_pair, synth_state_dict
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
net = S.IGEVStereo({"max_disp": 256})
net.load_state_dict(synth_state_dict(net.state_dict(), 0), strict=True)
net = net.cuda().eval()
left, right = (t.cuda() for t in synth_pair(1, 1152, 1920, seed=4, shift=9))
outs = {}
with torch.no_grad():
    for tag, mode, cl in (("torch_nchw", "torch", False), ("torch_nhwc", "torch", True), ("umma_nhwc", "auto", True), ("umma_nchw", "auto", False)):
        net.update_mode, net.channels_last = mode, cl
        for it in (4, 32):
            outs[(tag, it)] = net(left, right, iters=it).float().cpu()
for it in (4, 32):
    ref = outs[("torch_nchw", it)]
    for tag in ("torch_nhwc", "umma_nhwc", "umma_nchw"):
        d = (outs[(tag, it)] - ref).abs()
        print(f"iters={it:2d} {tag:11s} vs torch_nchw: mean {d.mean().item():.3e} px  median {d.median().item():.3e}  max {d.max().item():.3e}")
