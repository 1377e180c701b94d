"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz from the untouched reference.

Run in the build container (``python -m oracle.make_golden``); needs /root/reference.
The fixtures are small and committed; this script is committed beside them so they can
be regenerated.  Weights are NOT stored: they are the reference's default random init
under ``torch.manual_seed(seed)`` (biases optionally re-filled from a second seeded
generator), which ``vstnet_b200.RevResNet`` reproduces bit-for-bit; a sha256 of the
state_dict is stored so the tests can prove it.
"""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

MODES = {"photo": dict(hidden_dim=16, sp_steps=2), "art": dict(hidden_dim=64, sp_steps=1)}


def sd_sha256(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def fill_biases(net, seed):
    """Deterministic non-zero biases (the reference zeroes them at init, RevResNet.py:91-94;
    trained checkpoints have non-zero ones)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, p in net.named_parameters():
            if k.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)


def build_ref(mode, seed, bias_seed=None):
    RevResNet, _ = ref_shim.load()
    torch.manual_seed(seed)
    net = RevResNet(**MODES[mode]).eval()
    if bias_seed is not None:
        fill_biases(net, bias_seed)
    return net


def rand_img(seed, h, w):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(1, 3, h, w, generator=g)


def blocky_mask(h, w, gy, gx, perm, seed=None):
    """uint8 [1,h,w]: a gy x gx grid of rectangles labelled by ``perm``."""
    m = np.zeros((h, w), np.uint8)
    ys = np.linspace(0, h, gy + 1).astype(int)
    xs = np.linspace(0, w, gx + 1).astype(int)
    k = 0
    for i in range(gy):
        for j in range(gx):
            m[ys[i]:ys[i + 1], xs[j]:xs[j + 1]] = perm[k]
            k += 1
    return m[None]


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    RevResNet, cWCT = ref_shim.load()
    cw = cWCT()

    with torch.no_grad():
        # ---- RevResNet encode / decode, both modes, zero and non-zero biases
        for mode in MODES:
            for bias_seed in (None, 7):
                net = build_ref(mode, 0, bias_seed)
                x = rand_img(11, 32, 48)
                z = net(x, forward=True)
                xr = net(z, forward=False)
                g = torch.Generator().manual_seed(12)
                zr = torch.randn(z.shape, generator=g) * 0.5
                xd = net(zr, forward=False)          # decode of an arbitrary latent
                tag = "%s_b%s" % (mode, "0" if bias_seed is None else str(bias_seed))
                np.savez_compressed(
                    os.path.join(OUT, "revnet_%s.npz" % tag),
                    sha256=np.array(sd_sha256(net.state_dict())),
                    x=x.numpy(), z=z.numpy(), x_roundtrip=xr.numpy(),
                    z_rand=zr.numpy(), x_dec=xd.numpy(),
                    roundtrip_err=np.array((xr - x).abs().max().item()))
                print(tag, "z", tuple(z.shape), "rt", (xr - x).abs().max().item())

        # ---- cWCT: unmasked (via the working interpolation path), alpha, multi-style
        g = torch.Generator().manual_seed(21)
        for C in (32, 128):
            zc = torch.randn(1, C, 24, 20, generator=g) * 0.4 + 0.1
            zs = torch.randn(1, C, 16, 28, generator=g) * 0.7 - 0.2
            zs2 = torch.rand(1, C, 12, 36, generator=g)
            # correlate channels so the covariance is not ~diagonal
            mix = torch.randn(C, C, generator=g) / C ** 0.5
            zc = torch.einsum("ij,bjhw->bihw", mix, zc).contiguous()
            zs = torch.einsum("ij,bjhw->bihw", mix.T, zs).contiguous()
            o0 = cw.interpolation(zc, [zs], [1.0], 0.0)
            o5 = cw.interpolation(zc, [zs], [1.0], 0.5)
            om = cw.interpolation(zc, [zs, zs2], [0.3, 0.7], 0.25)
            # per-sample 2-D path == intended _transfer
            o2 = cw.coloring(cw.whitening(zc[0].reshape(C, -1)), zs[0].reshape(C, -1)).reshape(zc.shape)
            assert (o2 - o0).abs().max().item() < 1e-5
            np.savez_compressed(os.path.join(OUT, "cwct_plain_c%d.npz" % C), zc=zc.numpy(), zs=zs.numpy(),
                                zs2=zs2.numpy(), out_a0=o0.numpy(), out_a05=o5.numpy(), out_multi=om.numpy())
            print("cwct C", C, o0.abs().max().item())

        # ---- cWCT masked: 6 labels, one invalid (label 5: too few style px), one absent in content
        C = 32
        zc = torch.randn(1, C, 24, 32, generator=g) * 0.4
        zs = torch.randn(1, C, 20, 36, generator=g) * 0.6 + 0.3
        cm = blocky_mask(24, 32, 2, 3, [0, 1, 2, 3, 4, 5])
        sm = blocky_mask(20, 36, 3, 2, [3, 1, 0, 2, 4, 7])
        sm[0, :2, :3] = 5                      # 6 px of label 5 in style -> invalid (<=10)
        zc_in = zc.clone()
        om = cw.transfer(zc_in, zs, cm, sm)
        assert (zc_in - om).abs().max().item() == 0.0      # reference mutates its input in place
        np.savez_compressed(os.path.join(OUT, "cwct_masked_c32.npz"), zc=zc.numpy(), zs=zs.numpy(),
                            cmask=cm, smask=sm, out=om.numpy())
        # all-ones masks == unmasked
        ones_c, ones_s = np.ones((1, 24, 32), np.uint8), np.ones((1, 20, 36), np.uint8)
        oo = cw.transfer(zc.clone(), zs, ones_c, ones_s)
        oi = cw.interpolation(zc, [zs], [1.0], 0.0)
        print("masked ok; ones-vs-interp", (oo - oi).abs().max().item())

        # ---- end to end, tiny versions of cfg1 / cfg2 / cfg3
        for mode, h, w, alpha in (("photo", 64, 48, None), ("art", 64, 48, 0.5)):
            net = build_ref(mode, 0, 7)
            c, s = rand_img(31, h, w), rand_img(32, h + 8, w - 8)
            zc, zs = net(c), net(s)
            if alpha is None:
                zcs = cw.interpolation(zc, [zs], [1.0], 0.0)
            else:
                zcs = cw.interpolation(zc, [zs], [1.0], alpha)
            y = net(zcs, forward=False)
            np.savez_compressed(os.path.join(OUT, "e2e_%s.npz" % mode), content=c.numpy(), style=s.numpy(),
                                stylized=y.numpy(), alpha_c=np.array(-1.0 if alpha is None else alpha))
            print("e2e", mode, y.min().item(), y.max().item())
        net = build_ref("photo", 0, 7)
        c, s = rand_img(41, 64, 64), rand_img(42, 48, 80)
        cm = blocky_mask(64, 64, 2, 2, [0, 1, 2, 3])
        sm = blocky_mask(48, 80, 2, 2, [2, 3, 1, 0])
        zc, zs = net(c), net(s)
        y = net(cw.transfer(zc, zs, cm, sm), forward=False)
        np.savez_compressed(os.path.join(OUT, "e2e_photo_masked.npz"), content=c.numpy(), style=s.numpy(),
                            cmask=cm, smask=sm, stylized=y.numpy())
        print("e2e masked", y.min().item(), y.max().item())


if __name__ == "__main__":
    main()
