"""Developer check: the whole path is deterministic run to run (catches protocol races in the role pipelines):
N repetitions of encode -> cWCT -> decode at several sizes, every output compared bit for bit with the first."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vstnet_b200 import RevResNet
from vstnet_b200.video import VideoStylizer
dev = torch.device("cuda:0")
ok = True
for mode, kw in (("photo", dict(hidden_dim=16, sp_steps=2)), ("art", dict(hidden_dim=64, sp_steps=1))):
    torch.manual_seed(0)
    net = RevResNet(**kw).to(dev).eval()
    for (h, w) in ((1080, 1920), (264, 520), (72, 136), (600, 40)):
        g = torch.Generator(device=dev).manual_seed(h * w)
        style = torch.rand(1, 3, h, w, device=dev, generator=g)
        frames = [torch.rand(1, 3, h, w, device=dev, generator=g) for _ in range(2)]
        vs = VideoStylizer(net, n_streams=4)
        vs.set_style(style)
        ref = [vs.stylize(f).clone() for f in frames]
        bad = 0
        for rep in range(12 if h >= 1000 else 30):
            outs = [y.clone() for y in vs.stylize_frames(frames)]
            torch.cuda.synchronize()
            bad += sum(int(not torch.equal(a, b)) for a, b in zip(outs, ref))
        print("%s %dx%d: %d mismatching outputs" % (mode, h, w, bad))
        ok &= bad == 0
print("DETERMINISTIC" if ok else "NON-DETERMINISTIC")
