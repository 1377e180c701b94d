"""Developer tool: two fused 1080p video frames (set_style + stylize) on one stream — the command ncu wraps for the
cWCT-on-the-state kernels (gram_tc_kernel<4,STATE>, factor_kernel, apply_state_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vstnet_b200 import RevResNet, cWCT
from vstnet_b200.video import VideoStylizer
dev = torch.device("cuda:0"); torch.manual_seed(0)
net = RevResNet(hidden_dim=16, sp_steps=2).to(dev).eval()
g = torch.Generator(device=dev)
style = torch.rand(1, 3, 1080, 1920, device=dev, generator=g.manual_seed(1))
frame = torch.rand(1, 3, 1080, 1920, device=dev, generator=g.manual_seed(2))
vs = VideoStylizer(net, cWCT(), n_streams=1)
vs.set_style(style)
for _ in range(2):
    out = vs.stylize(frame)
torch.cuda.synchronize()
print("ok", tuple(out.shape), float(out.mean()))
