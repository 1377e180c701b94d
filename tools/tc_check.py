"""Developer check of the tcgen05 conv path: accuracy per precision mode vs the CPU oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import vst_oracle as O
from tests.helpers import MODES, build_net, cpu_state_dict
from vstnet_b200 import cWCT

dev = torch.device("cuda:0")
mode = sys.argv[1] if len(sys.argv) > 1 else "photo"
h, w = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (72, 136)
net = build_net(mode, 0, 7)
sd = cpu_state_dict(net)
net = net.to(dev)
g = torch.Generator().manual_seed(5)
x, s = torch.rand(1, 3, h, w, generator=g), torch.rand(1, 3, h, w, generator=g)
with torch.no_grad():
    zr = O.revnet_forward(sd, x, **MODES[mode]); zsr = O.revnet_forward(sd, s, **MODES[mode])
    xr_ref = O.revnet_inverse(sd, zr, **MODES[mode])
    yr = O.revnet_inverse(sd, O.cwct_transfer(zr, zsr), **MODES[mode])
e_ref = (xr_ref - x).abs()
print("reference round trip: max %.3e mean %.3e" % (e_ref.max(), e_ref.mean()))
for prec in ("fp32", "tf32", "tf32x2", "tf32x3", "f16x2"):
    net.precision = prec
    z = net(x.to(dev)); zs = net(s.to(dev))
    xr = net(z, forward=False)
    y = net(cWCT().transfer(z, zs), forward=False)
    torch.cuda.synchronize()
    e = (xr.cpu() - x).abs()
    print("%-7s z err %.3e | stylized err %.3e | round trip max %.3e mean %.3e" % (
        prec, (z.cpu() - zr).abs().max(), (y.cpu() - yr).abs().max(), e.max(), e.mean()))
