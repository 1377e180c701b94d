"""Shared test helpers (CPU side): weight construction, hashing, synthetic inputs."""
import hashlib
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
MODES = {"photo": dict(hidden_dim=16, sp_steps=2), "art": dict(hidden_dim=64, sp_steps=1)}


def sd_sha256(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def fill_biases(net, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, p in net.named_parameters():
            if k.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)


def build_net(mode, seed=0, bias_seed=None, **kw):
    """The product module with the reference's default random init (CPU parameters)."""
    from vstnet_b200 import RevResNet
    torch.manual_seed(seed)
    net = RevResNet(**MODES[mode], **kw).eval()
    if bias_seed is not None:
        fill_biases(net, bias_seed)
    return net


def cpu_state_dict(net):
    return {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}


def rand_img(seed, h, w):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(1, 3, h, w, generator=g)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def blocky_mask(h, w, gy, gx, perm):
    m = np.zeros((h, w), np.uint8)
    ys = np.linspace(0, h, gy + 1).astype(int)
    xs = np.linspace(0, w, gx + 1).astype(int)
    k = 0
    for i in range(gy):
        for j in range(gx):
            m[ys[i]:ys[i + 1], xs[j]:xs[j + 1]] = perm[k]
            k += 1
    return m[None]
