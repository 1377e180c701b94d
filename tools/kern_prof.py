"""Developer tool: per-launch time of every kernel class of one 1080p encode pass (range check off, so knock-out builds
whose results are garbage keep the f16x2 path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vstnet_b200 import RevResNet, _lib
dev = torch.device("cuda:0"); torch.manual_seed(0)
net = RevResNet(hidden_dim=16, sp_steps=2).to(dev).eval()
net.range_check = "off"
x = torch.rand(1, 3, 1080, 1920, device=dev)
for _ in range(3): net(x)
torch.cuda.synchronize()
_lib.profile_enable(True)
for _ in range(5): net(x)
_lib.profile_enable(False); torch.cuda.synchronize()
pat = sys.argv[1] if len(sys.argv) > 1 else ""
for k, v in sorted(_lib.profile_collect().items()):
    if pat in k: print("%-30s %.4f ms/launch x %d" % (k, v["ms"] / v["launches"], v["launches"]))
