"""Runs every cWCT kernel once per variant on benchmark-sized latents (for ncu / timing):
photo 1080p latent (C=32, n=1920x1080) unmasked and with 8 labels; artistic latent (C=128, n=1024x1024) unmasked and masked.
    python tools/cwct_prof.py [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vstnet_b200 import cWCT, _lib  # noqa: E402


def blocky(h, w, gy, gx, perm):
    m = np.zeros((h, w), np.uint8)
    ys, xs = np.linspace(0, h, gy + 1).astype(int), np.linspace(0, w, gx + 1).astype(int)
    for i in range(gy):
        for j in range(gx):
            m[ys[i]:ys[i + 1], xs[j]:xs[j + 1]] = perm[i * gx + j]
    return m[None]


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    cw = cWCT()
    for C, h, w in ((32, 1080, 1920), (128, 1024, 1024)):
        zc = torch.randn(1, C, h, w, device=dev, generator=g)
        zs = torch.randn(1, C, h, w, device=dev, generator=g) * 0.5 + 0.1
        cm = torch.from_numpy(blocky(h, w, 2, 4, [0, 1, 2, 3, 4, 5, 6, 7])).to(dev)
        sm = torch.from_numpy(blocky(h, w, 4, 2, [3, 1, 0, 2, 7, 6, 4, 5])).to(dev)
        for _ in range(reps):
            _lib.profile_enable(True)
            cw.transfer(zc, zs)
            cw.transfer(zc.clone(), zs, cm, sm)
            _lib.profile_enable(False)
            torch.cuda.synchronize()
            for k, v in sorted(_lib.profile_collect().items()):
                print("C=%d %-22s %.4f ms/launch x %d  %.0f GB/s" % (C, k, v["ms"] / v["launches"], v["launches"],
                                                                    v["bytes"] / v["ms"] / 1e6 if v["ms"] > 0 else 0))
        del zc, zs
    print("ok")


if __name__ == "__main__":
    main()
