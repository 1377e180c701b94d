"""Developer tool: one small pass through every kernel of the path, for compute-sanitizer
   (compute-sanitizer --tool memcheck python tools/sanitize_small.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vstnet_b200 import RevResNet, cWCT
from vstnet_b200.video import VideoStylizer
dev = torch.device("cuda:0"); torch.manual_seed(0)
net = RevResNet(hidden_dim=16, sp_steps=2).to(dev).eval()
for h, w in ((72, 136), (132, 260)):
    x = torch.rand(1, 3, h, w, device=dev); s = torch.rand(1, 3, h, w, device=dev)
    z = net(x); zs = net(s)
    cw = cWCT()
    zt = cw.transfer(z, zs)
    m1 = torch.from_numpy((np.arange(z.shape[2] * z.shape[3]).reshape(1, z.shape[2], z.shape[3]) % 3).astype(np.uint8)).to(dev)
    zt2 = cw.transfer(z.clone(), zs, m1, m1)
    out = net.inverse(zt)
    vs = VideoStylizer(net, cWCT(), n_streams=1); vs.set_style(s)
    o2 = vs.stylize(x)
    torch.cuda.synchronize()
    print("ok", h, w, float(out.mean()), float(o2.mean()))
