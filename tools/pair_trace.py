"""Developer tool: leader / peer timeline of CTA pair 0 of the 64>256 pair conv (VST_TC_TRACE=64,256)."""
import sys, os, ctypes as C
os.environ["VST_TC_TRACE"] = "64,256"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vstnet_b200 import RevResNet, _lib
dev = torch.device("cuda:0"); torch.manual_seed(0)
net = RevResNet(hidden_dim=16, sp_steps=2).to(dev).eval()
x = torch.rand(1, 3, 1080, 1920, device=dev)
net(x); torch.cuda.synchronize(); net(x); torch.cuda.synchronize()
lib = C.CDLL(_lib.LIB_PATH)
buf = (C.c_longlong * (8 * 4096))()
print("rc", lib.vst_debug_tc_trace(buf, 8 * 4096))
a = np.frombuffer(buf, dtype=np.int64).reshape(8, 4096)
t0 = a[a > 0].min()
f = lambda r, lo, n: " ".join("%6d" % (q - t0 if q > 0 else -1) for q in a[r][lo:lo + n])
n = 32
print("it             :", " ".join("%6d" % i for i in range(n)))
print("L prod issue   :", f(0, 0, n))
print("P prod issue   :", f(0, 2048, n))
print("P relay loaded :", f(1, 2048, n))
print("L wait begins  :", f(1, 0, n))
print("L local loaded :", f(2, 0, n))
print("L peer loaded  :", f(3, 0, n))
print("L issued       :", f(4, 0, n))
print("tile begin     :", f(5, 0, 8))
print("epi acc full   :", f(6, 0, 8))
print("epi done       :", f(7, 0, 8))
