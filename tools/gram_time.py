"""Developer tool: time of the state Gram / apply kernels of a fused 1080p frame (profile table)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vstnet_b200 import RevResNet, cWCT, _lib
from vstnet_b200.video import VideoStylizer
dev = torch.device("cuda:0"); torch.manual_seed(0)
net = RevResNet(hidden_dim=16, sp_steps=2).to(dev).eval()
net.range_check = "off"
g = torch.Generator(device=dev)
style = torch.rand(1, 3, 1080, 1920, device=dev, generator=g.manual_seed(1))
frame = torch.rand(1, 3, 1080, 1920, device=dev, generator=g.manual_seed(2))
vs = VideoStylizer(net, cWCT(), n_streams=1)
vs.set_style(style)
for _ in range(2): vs.stylize(frame)
torch.cuda.synchronize()
_lib.profile_enable(True)
for _ in range(5): vs.stylize(frame)
_lib.profile_enable(False); torch.cuda.synchronize()
for k, v in sorted(_lib.profile_collect().items()):
    if "cwct" in k: print("%-26s %.4f ms/launch x %d" % (k, v["ms"] / v["launches"], v["launches"]))
