"""Developer tool: per-launch time of the fused block kernels of one 1080p encode pass (needs a -DBTC_KNOCKOUT build
for VST_BTC_KO to have an effect: bit 0 no x loads, bit 1 no coupling-operand loads, bit 2 no stores)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vstnet_b200 import RevResNet, _lib
dev = torch.device("cuda:0"); torch.manual_seed(0)
net = RevResNet(hidden_dim=16, sp_steps=2).to(dev).eval()
x = torch.rand(1, 3, 1080, 1920, device=dev)
for _ in range(3): net(x)
torch.cuda.synchronize()
_lib.profile_enable(True)
for _ in range(5): net(x)
_lib.profile_enable(False); torch.cuda.synchronize()
for k, v in sorted(_lib.profile_collect().items()):
    if "rev_block" in k: print("KO=%s %-26s %.4f ms/launch x %d" % (os.environ.get("VST_BTC_KO", "0"), k, v["ms"] / v["launches"], v["launches"]))
