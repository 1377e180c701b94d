"""One encode pass at 1080p (developer tool: the command ncu wraps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vstnet_b200 import RevResNet

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x2"
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1080, 1920)
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = RevResNet(hidden_dim=16, sp_steps=2, precision=prec).to(dev).eval()
x = torch.rand(1, 3, H, W, device=dev)
z = net(x)
torch.cuda.synchronize()
print("ok", tuple(z.shape), float(z.abs().mean()))
