"""Host-side pre/post-processing shared by the entry points (PIL / numpy only, no CUDA).

Mirrors the reference's ``utils/utils.py`` behaviour that the stylization path depends on:
``img_resize`` (:90-101) and ``load_segment`` (:104-153, colour -> label 0..8).  ``load_segment`` is
vectorised (the reference loops over pixels in Python) but keeps the same colour table and the
nearest-colour (L1) rule for colours that are not in it.
"""
from __future__ import annotations

import os

import numpy as np
from PIL import Image

SEG_COLORS = np.array([  # ref: utils/utils.py:106-116, index = label
    (0, 0, 0), (255, 255, 255), (0, 255, 0), (0, 0, 255), (255, 0, 0), (255, 255, 0), (128, 128, 128),
    (0, 255, 255), (255, 0, 255)], dtype=np.int32)


def resized_dims(w, h, max_size, down_scale=None):
    """The (w, h) sequence ``img_resize`` goes through: first the long edge capped at ``max_size`` (aspect kept,
    both sides truncated toward zero), then both sides floored to multiples of ``down_scale``.  Each entry is one
    PIL bicubic resize in the reference (utils/utils.py:90-101) — including the second one when it is a no-op."""
    steps = []
    long_edge = max(w, h)
    if long_edge > max_size:
        w, h = int(1.0 * w / long_edge * max_size), int(1.0 * h / long_edge * max_size)
        steps.append((w, h))
    if down_scale is not None:
        w, h = w // down_scale * down_scale, h // down_scale * down_scale
        steps.append((w, h))
    return steps


def img_resize(img, max_size, down_scale=None):
    """Cap the long edge at ``max_size`` and floor H, W to multiples of ``down_scale`` (PIL bicubic per step)."""
    for size in resized_dims(img.size[0], img.size[1], max_size, None if down_scale is None else int(down_scale)):
        img = img.resize(size, Image.BICUBIC)
    return img


def labels_from_colors(rgb):
    """uint8 [H,W,3] colour-coded segmentation -> uint8 [H,W] labels (nearest table colour in L1)."""
    a = np.asarray(rgb, dtype=np.int32)
    d = np.abs(a[:, :, None, :] - SEG_COLORS[None, None, :, :]).sum(-1)          # [H,W,9]
    # the reference visits the table in dict order (blue, green, black, white, red, ...) and keeps the first
    # strict minimum; reproduce that order for ties
    order = np.array([3, 2, 0, 1, 4, 5, 6, 7, 8])
    return order[np.argmin(d[:, :, order], axis=-1)].astype(np.uint8)


def load_segment(image_path, size=None, device=None):
    """ref: utils/utils.py:104-153.  ``size`` = (w, h) nearest-neighbour resize before labelling.  With ``device`` the
    colour -> label step runs on the GPU (``vstnet_b200.segmentation.labels_from_colors``) and a uint8 CUDA tensor is
    returned, which ``cWCT.transfer`` takes as it is; otherwise the vectorised numpy rule below."""
    if not os.path.exists(image_path):
        print("Can not find image path: %s " % image_path)
        return None
    image = Image.open(image_path).convert("RGB")
    if size is not None:
        image = image.resize((size[0], size[1]), Image.NEAREST)
    if device is not None:
        from .segmentation import labels_from_colors as device_labels
        return device_labels(np.array(image), device)
    return labels_from_colors(np.array(image))


def to_uint8_hwc(stylized):
    """fp32 CUDA/CPU [1,3,H,W] -> uint8 numpy [H,W,3]: mul(255).clamp(0,255).byte() (image_transfer.py:218)."""
    return stylized[0].mul(255).clamp(0, 255).byte().permute(1, 2, 0).cpu().numpy()
