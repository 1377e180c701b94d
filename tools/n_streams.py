import sys, os, time
sys.path.insert(0, "/root/repo")
import torch
from vstnet_b200 import RevResNet
from vstnet_b200.video import VideoStylizer
dev = torch.device("cuda:0"); torch.manual_seed(0)
net = RevResNet(hidden_dim=16, sp_steps=2).to(dev).eval()
H, W = 1080, 1920
frames = [torch.rand(1, 3, H, W, device=dev) for _ in range(4)]
style = torch.rand(1, 3, H, W, device=dev)
for ns in (1, 3, 4, 5, 6, 8, 4):
    vs = VideoStylizer(net, n_streams=ns); vs.set_style(style)
    for _ in vs.stylize_frames(frames[i % 4] for i in range(2 * ns + 2)): pass
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in vs.stylize_frames(frames[i % 4] for i in range(48)): pass
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("%d streams: %.2f frames/s" % (ns, 48 / dt))
