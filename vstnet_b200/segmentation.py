"""Mask preparation on the device: the reference's ``models/segmentation/SegReMapping.py`` (label re-mapping of ADE20K
segmentations) and ``utils/utils.py:load_segment``'s colour -> label rule, as streaming CUDA kernels behind the C ABI
(``vst_seg_remap``, ``vst_seg_labels_from_colors``).  Label maps go in as uint8 CUDA tensors (numpy arrays are uploaded)
and come out as uint8 CUDA tensors that ``cWCT.transfer`` takes as they are — no host round trip, no ``np.unique`` /
per-label boolean masks / per-pixel Python loop.  There is no CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _as_device_labels(seg, device):
    if isinstance(seg, torch.Tensor):
        t = seg
    else:
        a = np.ascontiguousarray(np.asarray(seg))
        if a.dtype != np.uint8:
            if a.min() < 0 or a.max() > 255:
                raise ValueError("labels must fit uint8")
            a = a.astype(np.uint8)
        t = torch.from_numpy(a)
    if t.dtype != torch.uint8:
        t = t.to(torch.uint8)
    if device is None:
        device = t.device if t.is_cuda else torch.device("cuda")
    return t.to(device).contiguous()


class SegReMapping:
    """ref: models/segmentation/SegReMapping.py:5-76 (same constructor and methods; CUDA tensors in and out)."""

    def __init__(self, mapping_name, min_ratio=0.01, device=None):
        table = np.load(mapping_name) if isinstance(mapping_name, str) else np.asarray(mapping_name)
        if table.ndim != 2 or table.shape[1] > 256:
            raise ValueError("label_mapping must be [rows, n_classes <= 256]")
        self.label_mapping = table
        self.min_ratio = min_ratio
        self.label_ipt = []
        self._lib = _lib.load()
        self._dev_table = {}
        self.device = device

    def _table(self, device):
        k = str(device)
        if k not in self._dev_table:
            self._dev_table[k] = torch.from_numpy(np.ascontiguousarray(self.label_mapping.astype(np.int32))).to(device)
        return self._dev_table[k]

    def _remap(self, seg, style_seg):
        s = _as_device_labels(seg, self.device)
        dev = s.device
        t = _as_device_labels(style_seg, dev) if style_seg is not None else None
        out = torch.empty_like(s)
        scratch = torch.empty(int(self._lib.vst_seg_scratch_bytes()), dtype=torch.uint8, device=dev)
        table = self._table(dev)
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(self._lib.vst_seg_remap(s.data_ptr(), s.numel(), t.data_ptr() if t is not None else None,
                                               t.numel() if t is not None else 0, table.data_ptr(), int(table.shape[0]),
                                               int(table.shape[1]), float(self.min_ratio), out.data_ptr(),
                                               scratch.data_ptr(), st), "vst_seg_remap")
        return out

    @torch.no_grad()
    def self_remapping(self, seg):
        """Labels with a small share of the map are merged into their closest large label (ref :49-76)."""
        return self._remap(seg, None)

    @torch.no_grad()
    def cross_remapping(self, content_seg, style_seg):
        """Content labels the style lacks are assigned the best matching style label (ref :19-46)."""
        return self._remap(content_seg, style_seg)


@torch.no_grad()
def labels_from_colors(rgb, device=None):
    """uint8 [H,W,3] colour-coded segmentation (numpy or tensor) -> uint8 CUDA [H,W] labels 0..8 (ref: utils/utils.py:105-137)."""
    t = rgb if isinstance(rgb, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(rgb, dtype=np.uint8)))
    if t.dim() != 3 or t.shape[2] != 3 or t.dtype != torch.uint8:
        raise ValueError("expected a uint8 [H,W,3] image")
    if device is None:
        device = t.device if t.is_cuda else torch.device("cuda")
    t = t.to(device).contiguous()
    out = torch.empty(t.shape[0], t.shape[1], dtype=torch.uint8, device=device)
    lib = _lib.load()
    with torch.cuda.device(device):
        _lib.check(lib.vst_seg_labels_from_colors(t.data_ptr(), out.numel(), out.data_ptr(),
                                                  torch.cuda.current_stream(device).cuda_stream), "vst_seg_labels_from_colors")
    return out
