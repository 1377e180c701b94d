"""GPU parity of the fused row-streaming tensor-core block kernel (csrc/block_tc.cu) in isolation: small
networks whose stride-1 blocks all have C = 16 or C = 64, so the kernel's strips, segments, reflection
rows / columns and ring wrap-arounds are exercised directly against the CPU oracle.

Tolerance: latent max-abs <= 2e-4 (f16x2 arithmetic over <= 5 blocks; the full-network bar is 1e-3) and the
forward -> inverse round trip at fp32 rounding level (<= 4e-6 max).
"""
import pytest
import torch

from oracle import vst_oracle as O
from tests.helpers import cpu_state_dict, fill_biases

pytestmark = pytest.mark.gpu

NETS = {
    "c16": dict(nBlocks=[2], nStrides=[1], nChannels=[16], hidden_dim=16, sp_steps=0),
    "c64": dict(nBlocks=[1, 2], nStrides=[1, 2], nChannels=[16, 64], hidden_dim=64, sp_steps=0),
}


@pytest.mark.parametrize("name,h,w", [("c16", 8, 8), ("c16", 40, 72), ("c16", 132, 260), ("c16", 20, 508),
                                      ("c16", 300, 124), ("c64", 16, 16), ("c64", 72, 136), ("c64", 264, 520),
                                      ("c64", 600, 40)])
def test_block_tc_vs_oracle(name, h, w):
    from vstnet_b200 import RevResNet
    dev = torch.device("cuda:0")
    kw = NETS[name]
    torch.manual_seed(5)
    net = RevResNet(**kw, precision="f16x2").eval()
    fill_biases(net, 11)
    sd = cpu_state_dict(net)
    net = net.to(dev)
    g = torch.Generator().manual_seed(h * 1000 + w)
    x = torch.rand(1, 3, h, w, generator=g)
    okw = dict(nBlocks=tuple(kw["nBlocks"]), nStrides=tuple(kw["nStrides"]), nChannels=tuple(kw["nChannels"]),
               hidden_dim=kw["hidden_dim"], sp_steps=kw["sp_steps"])
    with torch.no_grad():
        zr = O.revnet_forward(sd, x, **okw)
    z = net(x.to(dev))
    torch.cuda.synchronize()
    err = float((z.cpu() - zr).abs().max())
    assert err <= 2e-4, err
    xr = net.inverse(z)
    rt = float((xr.cpu() - x).abs().max())
    assert rt <= 4e-6, rt
