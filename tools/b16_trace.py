"""Developer tool: phase timeline of one interior CTA of rev_block16_kernel (VST_TC_TRACE=16,4)."""
import sys, os, ctypes as C
os.environ["VST_TC_TRACE"] = "16,4"
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
a = np.frombuffer(buf, dtype=np.int64)[:32]
v = a[a > 0]
names = ["start", "pre-weights-wait", "weights", "g0", "g1", "g2", "g3", "t1+reflect", "conv2+reflect", "conv3a", "store a", "conv3b", "store b"]
print("stamps:", " ".join("%d" % (q - v[0]) for q in v))
for i in range(1, len(v)):
    print("  %-18s %6d cycles" % (names[i] if i < len(names) else "?", v[i] - v[i - 1]))
