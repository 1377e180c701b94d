"""Developer tool: role timeline of one CTA of gram_tc_masked_kernel (library built with -DGTC_TRACE as tools/ab/libTrace.so)."""
import os, shutil, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
shutil.copy(os.path.join(root, "vstnet_b200/libvstb200.so"), "/tmp/lib_keep.so")
shutil.copy(os.path.join(root, "tools/ab/libTrace.so"), os.path.join(root, "vstnet_b200/libvstb200.so"))
sys.path.insert(0, root)
import numpy as np, torch
from vstnet_b200 import cWCT
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
z = torch.randn(1, 32, 1080, 1920, device=dev, generator=g)
m = np.zeros((1, 1080, 1920), np.uint8)
for i in range(2):
    for j in range(4):
        m[0, i * 540:(i + 1) * 540, j * 480:(j + 1) * 480] = i * 4 + j
mt = torch.from_numpy(m).to(dev)
cw = cWCT()
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    cw._stats(z[0], 32, 1080 * 1920, mt.reshape(-1), 8, st, (1080, 1920))
torch.cuda.synchronize()
shutil.copy("/tmp/lib_keep.so", os.path.join(root, "vstnet_b200/libvstb200.so"))
