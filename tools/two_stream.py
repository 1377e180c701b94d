"""Developer experiment: frames/s with the frames of a video dealt to 1 / 2 / 3 compute streams of one GPU
(kernel tails and wave-quantisation gaps of one frame are filled by the other frame's kernels)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vstnet_b200 import RevResNet, cWCT
from vstnet_b200.video import VideoStylizer
dev = torch.device("cuda:0"); torch.manual_seed(0)
net = RevResNet(hidden_dim=16, sp_steps=2).to(dev).eval()
vs = VideoStylizer(net)
H, W = 1080, 1920
vs.set_style(torch.rand(1, 3, H, W, device=dev))
frames = [torch.rand(1, 3, H, W, device=dev) for _ in range(4)]
def run(ns, n=40):
    streams = [torch.cuda.Stream(dev) for _ in range(ns)]
    outs = [None] * ns
    torch.cuda.synchronize()
    for k in range(ns * 2):                      # warm-up (workspaces per stream)
        with torch.cuda.stream(streams[k % ns]): outs[k % ns] = vs.stylize(frames[k % 4])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(n):
        with torch.cuda.stream(streams[k % ns]): outs[k % ns] = vs.stylize(frames[k % 4])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%d stream(s): %.2f frames/s (%.3f ms / frame)" % (ns, n / dt, 1e3 * dt / n))
for ns in (1, 2, 3, 1, 2):
    run(ns)
