"""Developer tool: per-role timeline of one interior CTA of rev_block_tc_kernel (VST_TC_TRACE=1016,0 | 1064,0)."""
import sys, os, ctypes as C
Cc = int(sys.argv[1]) if len(sys.argv) > 1 else 16
os.environ["VST_TC_TRACE"] = "%d,0" % (1000 + Cc)
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
def show(role, name, per, labels, rows=range(10, 16)):
    r = a[role]
    print("== %s (cycles since CTA start; per row: %s)" % (name, ", ".join(labels)))
    for k in rows:
        v = r[per * k: per * k + per]
        if (v > 0).any():
            print("  row %3d: " % k + " ".join("%8d" % (q - t0 if q > 0 else -1) for q in v))
    # steady-state period
    s = r[per * 10::per][:40]; s = s[s > 0]
    if len(s) > 2: print("  period: %.0f cycles/row" % ((s[-1] - s[0]) / (len(s) - 1)))
show(0, "issuer", 6, ["begin", "conv3 go", "conv3 done", "conv2 go", "conv1 go", "end"])
show(1, "converter", 3, ["begin", "x_empty ok", "arrived"])
show(4, "E3", 4, ["begin", "acc ok", "barrier", "end"])
show(5, "E3 block 0 detail", 8, ["tmem_ld done", "published", "barrier", "math done", "stored"])
show(6, "E3 chunk detail (C=64)", 8, ["chunk0 start", "tmem done", "math done", "stored", "chunk1 done"])
print("== per-role accounting: total cycles in the row loop, cycles inside barrier waits, steps")
for role, name in [(0, "issuer"), (3, "converter"), (1, "E1"), (2, "E2"), (4, "E3")]:
    tot, wt, n = a[7][role * 8: role * 8 + 3]
    if n > 0: print("  %-10s total %8d  wait %8d  steps %4d  ->  busy %.0f cycles/step, wait %.0f cycles/step" % (name, tot, wt, n, (tot - wt) / n, wt / n))
