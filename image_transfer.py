#!/usr/bin/env python
"""image_transfer.py — the reference's image entry point (image_transfer.py:15-221) on vstnet_b200.

Same flags; the hot calls are the reference's three lines (encode, cWCT, decode) against the drop-in
``models`` shim.  ``--auto_seg`` needs SegFormer/mmseg, which is outside this path and
is rejected; ``--synthetic HxW`` runs random-init weights on random images (the reference's checkpoints
are not shipped).
"""
import argparse
import os
import sys

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from vstnet_b200.hostio import img_resize, load_segment, to_uint8_hwc   # noqa: E402


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument('--mode', type=str, default='photorealistic')
    p.add_argument('--ckpoint', type=str, default='checkpoints/photo_image.pt')
    p.add_argument('--content', type=str, default='data/content/01.jpg')
    p.add_argument('--style', type=str, default='data/style/01.jpg')
    p.add_argument('--out_dir', type=str, default="output")
    p.add_argument('--max_size', type=int, default=1280)
    p.add_argument('--alpha_c', type=float, default=None)
    p.add_argument('--content_seg', type=str, default=None)
    p.add_argument('--style_seg', type=str, default=None)
    p.add_argument('--auto_seg', action='store_true', default=False)
    p.add_argument('--precision', type=str, default='f16x2', help="conv arithmetic: f16x2 | tf32x2 | tf32x3 | tf32 | fp32")
    p.add_argument('--synthetic', type=str, default=None, help="HxW: random images + random-init weights")
    return p


def build_network(mode, precision="f16x2"):
    from models.RevResNet import RevResNet
    if mode.lower() == "photorealistic":
        return RevResNet(hidden_dim=16, sp_steps=2, precision=precision)       # ref :45
    if mode.lower() == "artistic":
        return RevResNet(hidden_dim=64, sp_steps=1, precision=precision)       # ref :47
    raise NotImplementedError(mode)


def stylize(RevNetwork, cwct, content, style, content_seg=None, style_seg=None, alpha_c=None):
    """ref: image_transfer.py:172-201."""
    with torch.no_grad():
        if hasattr(RevNetwork, "encode_pair"):                  # the two encodes of ref :181-182 on two streams
            z_c, z_s = RevNetwork.encode_pair(content, style)
        else:
            z_c = RevNetwork(content, forward=True)
            z_s = RevNetwork(style, forward=True)
        if alpha_c is not None and content_seg is None and style_seg is None:
            assert 0.0 <= alpha_c <= 1.0
            z_cs = cwct.interpolation(z_c, styl_feat_list=[z_s], alpha_s_list=[1.0], alpha_c=alpha_c)
        else:
            z_cs = cwct.transfer(z_c, z_s, content_seg, style_seg)
        return RevNetwork(z_cs, forward=False)


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.auto_seg:
        raise SystemExit("--auto_seg needs an external ADE20K segmenter (mmseg SegFormer), which is outside the "
                         "stylization hot path; pass --content_seg/--style_seg label images instead")
    if not torch.cuda.is_available():
        raise SystemExit("vstnet_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    device = torch.device("cuda")
    os.makedirs(args.out_dir, exist_ok=True)

    RevNetwork = build_network(args.mode, args.precision)
    from torchvision import transforms
    if args.synthetic:
        h, w = (int(v) for v in args.synthetic.lower().split("x"))
        g = torch.Generator().manual_seed(0)
        content = torch.rand(1, 3, h // 4 * 4, w // 4 * 4, generator=g)
        style = torch.rand(1, 3, h // 4 * 4, w // 4 * 4, generator=g)
        content_seg = style_seg = None
        name = "synthetic_%dx%d.png" % (h, w)
    else:
        state_dict = torch.load(args.ckpoint)
        RevNetwork.load_state_dict(state_dict['state_dict'])                   # ref :52-53
        content = Image.open(args.content).convert('RGB')
        style = Image.open(args.style).convert('RGB')
        content = img_resize(content, args.max_size, down_scale=RevNetwork.down_scale)
        style = img_resize(style, args.max_size, down_scale=RevNetwork.down_scale)
        content_seg = style_seg = None
        if args.content_seg is not None and args.style_seg is not None:      # ref :150-160
            content_seg = load_segment(args.content_seg, content.size)[None, ...]
            style_seg = load_segment(args.style_seg, style.size)[None, ...]
        content = transforms.ToTensor()(content).unsqueeze(0)
        style = transforms.ToTensor()(style).unsqueeze(0)
        name = "%s_%s.png" % (os.path.basename(args.content).split(".")[0], os.path.basename(args.style).split(".")[0])
    RevNetwork = RevNetwork.to(device).eval()
    from models.cWCT import cWCT
    cwct = cWCT()

    stylized = stylize(RevNetwork, cwct, content.to(device), style.to(device), content_seg, style_seg, args.alpha_c)
    path = os.path.join(args.out_dir, name)
    Image.fromarray(to_uint8_hwc(stylized)).save(path, quality=100)         # ref :217-221
    print("Save at %s" % path)
    return stylized


if __name__ == "__main__":
    main()
