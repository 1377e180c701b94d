"""Developer tool: role timeline of CTA 0 of the last tcgen05 conv launch of a given shape
(VST_TC_TRACE="cin,cout" makes the library stamp clock64() per role event)."""
import sys, os, ctypes as C
shape = sys.argv[1] if len(sys.argv) > 1 else "16,64"
os.environ["VST_TC_TRACE"] = shape
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vstnet_b200 import RevResNet, _lib

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = RevResNet(hidden_dim=16, sp_steps=2, precision=(sys.argv[2] if len(sys.argv) > 2 else "f16x2")).to(dev).eval()
x = torch.rand(1, 3, 1080, 1920, device=dev)
net(x); torch.cuda.synchronize()
net(x); torch.cuda.synchronize()
lib = C.CDLL(_lib.LIB_PATH)
buf = (C.c_longlong * (8 * 4096))()
rc = lib.vst_debug_tc_trace(buf, 8 * 4096)
a = np.frombuffer(buf, dtype=np.int64).reshape(8, 4096)
names = ["prod_issue", "conv_loaded", "conv_ready", "mma_start", "mma_commit", "mma_tile_begin", "epi_acc_full", "epi_done"]
t0 = a[a > 0].min()
print("shape", shape, "rc", rc)
for r in range(8):
    v = a[r][a[r] > 0] - t0
    print("%-15s n=%4d : %s" % (names[r], len(v), " ".join("%7d" % q for q in v[:(40 if r == 7 else 36)])))
