"""TEST INFRASTRUCTURE ONLY — torch-fp32 CPU restatement of the reference hot path.

This is the *oracle*: a plain, functional, CPU fp32 restatement of what
``/root/reference/models/RevResNet.py`` and ``/root/reference/models/cWCT.py`` compute,
written from the algorithm, not copied from the reference.
It exists to check the CUDA path; the product (``vstnet_b200``) never imports it.

Parity status: **pinned** — ``oracle/make_golden.py`` imports the untouched reference in
the build container (with a stub for its absent ``todos`` debug package) and records
its outputs in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this
restatement against those fixtures.

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- RevResNet
def arch_blocks(nBlocks=(10, 10, 10), nStrides=(1, 2, 2), nChannels=(16, 64, 256)):
    """(channel, stride) per block of ``stack``.  ref: RevResNet.py:192-201."""
    out = []
    for ch, depth, st in zip(nChannels, nBlocks, nStrides):
        out += [(ch, st)] + [(ch, 1)] * (depth - 1)
    return out


def space_to_depth(x):
    """out[b,(dy*2+dx)*C+c,h,w] = in[b,c,2h+dy,2w+dx].  ref: RevResNet.py:34-37 (squeeze)."""
    b, c, h, w = x.shape
    x = x.reshape(b, c, h // 2, 2, w // 2, 2)       # b c h dy w dx
    x = x.permute(0, 3, 5, 1, 2, 4)                  # b dy dx c h w
    return x.reshape(b, 4 * c, h // 2, w // 2).contiguous()


def depth_to_space(x):
    """Inverse of space_to_depth.  ref: RevResNet.py:40-43 (unsqueeze), :141-144 (spread)."""
    b, c4, h, w = x.shape
    c = c4 // 4
    x = x.reshape(b, 2, 2, c, h, w)                  # b dy dx c h w
    x = x.permute(0, 3, 4, 1, 5, 2)                  # b c h dy w dx
    return x.reshape(b, c, 2 * h, 2 * w).contiguous()


def _conv(x, w, b, stride=1):
    """ReflectionPad2d(1) + 3x3 conv with bias.  ref: RevResNet.py:79-88."""
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w, b, stride=stride)


def block_F(sd, prefix, x, stride):
    """F(x) = conv(relu(conv(relu(conv_s(x))))).  ref: RevResNet.py:79-88."""
    t = torch.relu(_conv(x, sd[prefix + "conv.1.weight"], sd[prefix + "conv.1.bias"], stride))
    t = torch.relu(_conv(t, sd[prefix + "conv.4.weight"], sd[prefix + "conv.4.bias"]))
    return _conv(t, sd[prefix + "conv.7.weight"], sd[prefix + "conv.7.bias"])


def block_forward(sd, prefix, x1, x2, stride):
    """(x1,x2) -> (x2', F(x2)+x1').  ref: RevResNet.py:96-104."""
    f = block_F(sd, prefix, x2, stride)
    if stride == 2:
        x1, x2 = space_to_depth(x1), space_to_depth(x2)
    return x2, f + x1


def block_inverse(sd, prefix, x2, y1, stride):
    """(x2,y1) -> (x1,x2).  ref: RevResNet.py:106-116."""
    if stride == 2:
        x2 = depth_to_space(x2)
    x1 = y1 - block_F(sd, prefix, x2, stride)
    if stride == 2:
        x1 = depth_to_space(x1)
    return x1, x2


def revnet_forward(sd, x, nBlocks=(10, 10, 10), nStrides=(1, 2, 2), nChannels=(16, 64, 256),
                   hidden_dim=16, sp_steps=2, n_cr_blocks=2):
    """Encode.  ref: RevResNet.py:210-223 (_forward), :131-146 (channel_reduction.forward)."""
    b, cin, h, w = x.shape
    pad = 2 * nChannels[0] - cin                                        # ref :184
    x = torch.cat([x, x.new_zeros(b, pad, h, w)], 1)                    # injective_pad :24-28
    n = x.shape[1] // 2
    x1, x2 = x[:, :n].contiguous(), x[:, n:].contiguous()               # split :8-12
    for i, (ch, st) in enumerate(arch_blocks(nBlocks, nStrides, nChannels)):
        x1, x2 = block_forward(sd, f"stack.{i}.", x1, x2, st)
    cr_pad = hidden_dim * 4 ** sp_steps - nChannels[-1]                 # ref :122
    if cr_pad:
        z = x1.new_zeros(x1.shape[0], cr_pad, *x1.shape[2:])
        x1, x2 = torch.cat([x1, z], 1), torch.cat([x2, z], 1)
    for i in range(n_cr_blocks):
        x1, x2 = block_forward(sd, f"channel_reduction.block_list.{i}.", x1, x2, 1)
    x = torch.cat([x1, x2], 1)
    for _ in range(sp_steps):
        x = depth_to_space(x)
    return x


def revnet_inverse(sd, z, nBlocks=(10, 10, 10), nStrides=(1, 2, 2), nChannels=(16, 64, 256),
                   in_channel=3, hidden_dim=16, sp_steps=2, n_cr_blocks=2):
    """Decode.  ref: RevResNet.py:225-239 (_inverse), :148-163 (channel_reduction.inverse)."""
    x = z
    for _ in range(sp_steps):
        x = space_to_depth(x)
    n = x.shape[1] // 2
    x1, x2 = x[:, :n].contiguous(), x[:, n:].contiguous()
    for i in reversed(range(n_cr_blocks)):
        x1, x2 = block_inverse(sd, f"channel_reduction.block_list.{i}.", x1, x2, 1)
    cr_pad = hidden_dim * 4 ** sp_steps - nChannels[-1]
    if cr_pad:
        x1, x2 = x1[:, :x1.shape[1] - cr_pad], x2[:, :x2.shape[1] - cr_pad]
    blocks = arch_blocks(nBlocks, nStrides, nChannels)
    for i in reversed(range(len(blocks))):
        x1, x2 = block_inverse(sd, f"stack.{i}.", x1, x2, blocks[i][1])
    x = torch.cat([x1, x2], 1)
    return x[:, :in_channel].contiguous()                               # injective_pad.inverse :30-31


def init_state_dict(seed=0, nBlocks=(10, 10, 10), nStrides=(1, 2, 2), nChannels=(16, 64, 256), mult=4,
                    hidden_dim=16, sp_steps=2, n_cr_blocks=2):
    """The reference's random initialisation as a plain ``state_dict`` (no module of the product involved):
    ``torch.manual_seed(seed)``, then one default-initialised ``nn.Conv2d`` per convolution in the reference's
    construction order (per block conv.1, conv.4, conv.7; stack first, then channel_reduction), biases zeroed.
    ref: RevResNet.py:68-94 (residual_block.__init__ / init_layers), :119-129, :166-201."""
    import torch.nn as nn
    torch.manual_seed(seed)
    sd = {}

    def block(prefix, channel, stride):
        in_ch = channel if stride == 1 else channel // 4
        mid = channel // mult
        for k, (ci, co, st) in zip((1, 4, 7), ((in_ch, mid, stride), (mid, mid, 1), (mid, channel, 1))):
            conv = nn.Conv2d(ci, co, kernel_size=3, stride=st, padding=0, bias=True)
            sd["%sconv.%d.weight" % (prefix, k)] = conv.weight.detach().clone()
            sd["%sconv.%d.bias" % (prefix, k)] = torch.zeros(co)

    for i, (ch, st) in enumerate(arch_blocks(nBlocks, nStrides, nChannels)):
        block("stack.%d." % i, ch, st)
    for i in range(n_cr_blocks):
        block("channel_reduction.block_list.%d." % i, hidden_dim * 4 ** sp_steps, 1)
    return sd


# --------------------------------------------------------------------------- cWCT
def _chol_with_jitter(cov, eps):
    """Cholesky; on failure add eps*I, then 2eps*I, ... cumulatively.  ref: cWCT.py:111-128."""
    try:
        return torch.linalg.cholesky(cov)
    except RuntimeError:
        pass
    iden = torch.eye(cov.shape[-1], dtype=cov.dtype)
    e = eps
    while True:
        try:
            cov = cov + iden * e
            return torch.linalg.cholesky(cov)
        except RuntimeError:
            e = e + eps


def _mean_cov(x):
    """x [C,n] -> (mean [C], centred x, cov [C,C] with n-1 divisor).  ref: cWCT.py:138-144."""
    mu = x.mean(-1)
    xc = x - mu[:, None]
    return mu, xc, (xc @ xc.T) / (x.shape[-1] - 1)


def wct_2d(xc_feat, xs_feat, eps=2e-5, alpha_c=0.0, style_list=None, alpha_s=None):
    """Closed form of whitening+coloring for one [C,n_c] / [C,n_s] pair.

    out = mixL @ inv(Lc) @ (x-mu_c) + mix_mu.  ref: cWCT.py:134-164 (alpha_c=0, one style),
    :206-262 (interpolation).
    """
    mu_c, xc, cov_c = _mean_cov(xc_feat)
    Lc = _chol_with_jitter(cov_c, eps)
    inv_Lc = torch.inverse(Lc)
    whiten = inv_Lc @ xc
    styles = style_list if style_list is not None else [xs_feat]
    alphas = alpha_s if alpha_s is not None else [1.0]
    mixL = torch.zeros_like(Lc)
    mixmu = torch.zeros_like(mu_c)
    for s, a in zip(styles, alphas):
        mu_s, _, cov_s = _mean_cov(s)
        mixL = mixL + _chol_with_jitter(cov_s, eps) * a
        mixmu = mixmu + mu_s * a
    if alpha_c != 0.0:
        mixL = mixL * (1 - alpha_c) + Lc * alpha_c
        mixmu = mixmu * (1 - alpha_c) + mu_c * alpha_c
    return mixL @ whiten + mixmu[:, None]


def cwct_transfer(zc, zs, eps=2e-5):
    """Unmasked transfer, intended (upstream) semantics.  ref: cWCT.py:24-47."""
    B, C = zc.shape[:2]
    out = [wct_2d(zc[i].reshape(C, -1), zs[i].reshape(C, -1), eps) for i in range(B)]
    return torch.stack(out).reshape(zc.shape)


def cwct_interpolation(zc, zs_list, alpha_s_list, alpha_c=0.0, eps=2e-5):
    """ref: cWCT.py:206-262."""
    B, C = zc.shape[:2]
    out = []
    for i in range(B):
        out.append(wct_2d(zc[i].reshape(C, -1), None, eps, alpha_c,
                          [s[i].reshape(C, -1) for s in zs_list], list(alpha_s_list)))
    return torch.stack(out).reshape(zc.shape)


def label_validity(cmask, smask):
    """ref: cWCT.py:166-189.  Returns (label_set, indicator[max_label+1])."""
    cm, sm = np.asarray(cmask).reshape(-1), np.asarray(smask).reshape(-1)
    labels = np.unique(cm)
    ind = np.zeros(int(cm.max()) + 1)
    for l in labels:
        a, b = int((cm == l).sum()), int((sm == l).sum())
        ind[l] = a > 10 and b > 10 and a / b < 100 and b / a < 100
    return labels, ind


def cwct_transfer_seg(zc, zs, cmask, smask, eps=2e-5):
    """Per-label masked transfer; pixels of invalid labels keep content features.

    ref: cWCT.py:49-109.  (Returns a new tensor; the reference additionally writes the
    result into the caller's content_feat, cWCT.py:103.)
    """
    B, C = zc.shape[:2]
    out = zc.clone().reshape(B, C, -1)
    for i in range(B):
        xc, xs = zc[i].reshape(C, -1), zs[i].reshape(C, -1)
        cm, sm = np.asarray(cmask[i]).reshape(-1), np.asarray(smask[i]).reshape(-1)
        labels, ind = label_validity(cmask[i], smask[i])
        for l in labels:
            if not ind[l]:
                continue
            ci = torch.from_numpy(np.nonzero(cm == l)[0])
            si = torch.from_numpy(np.nonzero(sm == l)[0])
            out[i][:, ci] = wct_2d(xc[:, ci], xs[:, si], eps)
    return out.reshape(zc.shape)


# --------------------------------------------------------------------------- mask preparation
def seg_self_remapping(seg, table, min_ratio):
    """ref: models/segmentation/SegReMapping.py:49-76 (numpy version, the one the entry points call)."""
    seg = np.asarray(seg)
    n = seg.size
    labels = list(np.unique(seg))
    ratio = [np.float32((seg == l).sum()) / np.float32(n) for l in labels]
    new = list(labels)
    for i, l in enumerate(labels):
        if ratio[i] < np.float32(min_ratio):
            for j in range(table.shape[0]):
                nl = table[j, l]
                if nl in labels and ratio[labels.index(nl)] >= np.float32(min_ratio):
                    new[i] = nl
                    break
    out = seg.copy()
    for i, l in enumerate(labels):
        out[seg == l] = new[i]
    return out


def seg_cross_remapping(content_seg, style_seg, table):
    """ref: models/segmentation/SegReMapping.py:19-46."""
    content_seg, style_seg = np.asarray(content_seg), np.asarray(style_seg)
    cl, sl = list(np.unique(content_seg)), list(np.unique(style_seg))
    new = list(cl)
    for i, l in enumerate(cl):
        if l in sl:
            continue
        for j in range(table.shape[0]):
            nl = table[j, l]
            if nl in sl:
                new[i] = nl
                break
    out = content_seg.copy()
    for i, l in enumerate(cl):
        out[content_seg == l] = new[i]
    return out


SEG_COLOR_TABLE = [((0, 0, 255), 3), ((0, 255, 0), 2), ((0, 0, 0), 0), ((255, 255, 255), 1), ((255, 0, 0), 4),
                   ((255, 255, 0), 5), ((128, 128, 128), 6), ((0, 255, 255), 7), ((255, 0, 255), 8)]


def seg_labels_from_colors(rgb):
    """ref: utils/utils.py:105-137 — exact table colour, else the nearest in L1; the first minimum in dict order wins."""
    a = np.asarray(rgb).astype(np.int64)
    best = np.full(a.shape[:2], 1 << 30, np.int64)
    lab = np.zeros(a.shape[:2], np.uint8)
    for col, l in SEG_COLOR_TABLE:
        d = np.abs(a - np.array(col)).sum(-1)
        m = d < best
        best[m], lab[m] = d[m], l
    return lab
