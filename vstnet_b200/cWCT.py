"""Drop-in replacement for the reference's ``models/cWCT.py`` on B200.

Same constructor and methods: ``cWCT(eps=2e-5, use_double=False)``,
``transfer(content_feat, style_feat, cmask=None, smask=None)``,
``interpolation(content_feat, styl_feat_list, alpha_s_list, alpha_c=0.0)``
(ref: cWCT.py:9-262).  All arithmetic runs on the device in three hand-written kernels behind
the C ABI (stats -> factor -> apply, see csrc/cwct.cu); there is no host round trip, no
``np.where`` / index upload per label, and no CPU fallback.

Deliberate differences from the fork, both documented in DESIGN.md:
  * ``transfer`` without masks works (the fork's 3-D ``whitening`` raises); it
    returns the intended upstream result, equal to ``interpolation(c, [s], [1.0], 0.0)``.
  * masks may also be uint8 CUDA tensors (no upload); numpy uint8 ``[B,H,W]`` is accepted as in
    the reference and must be at latent resolution (cWCT.py:72-73 uses them as-is).
As in the reference, the masked path writes its result into ``content_feat`` in place
(cWCT.py:103) and returns it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib


def _check_feat(x, what):
    if not isinstance(x, torch.Tensor) or x.dim() != 4:
        raise ValueError("%s must be a 4-D tensor [B,N,H,W]" % what)
    if not x.is_cuda:
        raise RuntimeError("vstnet_b200.cWCT runs on CUDA (sm_100a) only; %s is on %s — there is no CPU fallback"
                           % (what, x.device))
    if x.dtype != torch.float32:
        raise ValueError("%s must be float32 (got %s)" % (what, x.dtype))


class cWCT(nn.Module):
    """Cholesky decomposition based WCT (ref: cWCT.py:9-16)."""

    def __init__(self, eps=2e-5, use_double=False):
        super().__init__()
        self.eps = eps
        self.use_double = use_double
        self._lib = _lib.load()
        self.last_status = None     # int32 [L] device tensor of the last call: #jitter retries per label
        self._pending = []          # (event, pinned status copy) of earlier unmasked calls, checked without a sync
        self._pin_ring = {}         # numpy mask staging: nbytes -> [slot index, [(pinned buffer, event), ...]]

    # ------------------------------------------------------------------ deferred failure reporting
    def _defer_status_check(self, status, stream):
        """An unmasked call whose Cholesky never succeeded (status -1 after 64 jitter retries) returns the content
        features unchanged.  The hot path must not synchronise, so the status words are copied to pinned memory
        asynchronously and inspected at the start of a LATER call, once their event has completed: a warning then
        names the failure.  ``check_status()`` inspects them now (synchronising)."""
        if len(self._pending) >= 8:                      # never grows without bound: the oldest is simply dropped
            self._pending.pop(0)
        host = torch.empty(status.numel(), dtype=torch.int32, pin_memory=True)   # caching host allocator
        host.copy_(status, non_blocking=True)            # on the current stream, after the factor kernel
        ev = torch.cuda.Event()
        ev.record(stream)
        self._pending.append((ev, host))

    def _poll_status(self, block=False):
        keep = []
        for ev, host in self._pending:
            if block:
                ev.synchronize()
            if ev.query():
                if bool((host < 0).any()):
                    import warnings
                    warnings.warn("vstnet_b200.cWCT: a Cholesky factorisation did not succeed within 64 jitter retries; "
                                  "that sample's content features were returned unstylized", RuntimeWarning)
            else:
                keep.append((ev, host))
        self._pending = keep

    def check_status(self):
        """Synchronise on the earlier calls' status words and warn about failed factorisations; returns last_status
        as a host list."""
        self._poll_status(block=True)
        return None if self.last_status is None else self.last_status.cpu().tolist()

    # ------------------------------------------------------------------ low-level steps
    def _stats(self, feat2d, C_, n, labels, L, stream, hw=None):
        buf = torch.empty(int(self._lib.vst_cwct_stats_bytes(C_, L)), dtype=torch.uint8, device=feat2d.device)
        if labels is not None and hw is not None and hw[0] * hw[1] == n:
            # the 2-D shape lets large per-label maps run on the tensor cores (strip-wise traversal)
            _lib.check(self._lib.vst_cwct_stats2d(feat2d.data_ptr(), C_, int(hw[0]), int(hw[1]), labels.data_ptr(), L,
                                                  buf.data_ptr(), stream), "vst_cwct_stats2d")
            return buf
        _lib.check(self._lib.vst_cwct_stats(feat2d.data_ptr(), C_, n, labels.data_ptr() if labels is not None else None,
                                            L, buf.data_ptr(), stream), "vst_cwct_stats")
        return buf

    def _factor(self, cstats, sstats_list, alpha_s, alpha_c, C_, L, masked, device, stream):
        T = torch.empty(L, C_, C_, dtype=torch.float32, device=device)
        mu = torch.empty(L, C_, dtype=torch.float32, device=device)
        beta = torch.empty(L, C_, dtype=torch.float32, device=device)
        valid = torch.empty(L, dtype=torch.int32, device=device)
        status = torch.empty(L, dtype=torch.int32, device=device)
        K = len(sstats_list)
        ptrs = (C.c_void_p * K)(*[s.data_ptr() for s in sstats_list])
        alphas = (C.c_float * K)(*[float(a) for a in alpha_s])
        _lib.check(self._lib.vst_cwct_factor(cstats.data_ptr(), ptrs, alphas, K, float(alpha_c), float(self.eps), C_, L,
                                             int(masked), int(bool(self.use_double)), T.data_ptr(), mu.data_ptr(),
                                             beta.data_ptr(), valid.data_ptr(), status.data_ptr(), stream),
                   "vst_cwct_factor")
        self.last_status = status
        if not masked:
            self._poll_status()
            self._defer_status_check(status, torch.cuda.current_stream(device))
        return T, mu, beta, valid

    def _apply(self, feat2d, out2d, C_, n, labels, L, T, mu, beta, valid, stream):
        _lib.check(self._lib.vst_cwct_apply(feat2d.data_ptr(), out2d.data_ptr(), C_, n,
                                            labels.data_ptr() if labels is not None else None, L, T.data_ptr(),
                                            mu.data_ptr(), beta.data_ptr(), valid.data_ptr(), stream), "vst_cwct_apply")

    def _mask_to_device(self, mask, n, device, what, hw=None):
        """One sample's label map -> (flat uint8 device tensor of n labels, number of label slots L).

        numpy masks (the reference's type) are staged through a small ring of pinned buffers and uploaded
        asynchronously on the current stream; their maximum is taken on the host, so L = max+1 as in the
        reference (cWCT.py:173-175) with no device synchronisation.  CUDA uint8 tensors are used in place and
        sized L = 256 (every possible label gets a slot: nothing has to be read back).  A 2-D map whose size
        differs from the latent's ``hw = (H, W)`` is resized on the device, nearest neighbour with PIL's
        sampling — the reference's ``cWCT.resize`` (cWCT.py:191-197), whose call this fork comments out
        (:72-73), so that there masks must already be at latent resolution (never true in artistic mode)."""
        if isinstance(mask, torch.Tensor):
            if mask.dtype != torch.uint8:
                raise ValueError("%s must be uint8" % what)
            if mask.is_cuda:
                m2, L = mask.to(device), _lib.MAX_LABELS
            else:
                m2, L = self._upload(mask.contiguous(), device), int(mask.max()) + 1
        else:
            a = np.ascontiguousarray(np.asarray(mask))
            if a.dtype != np.uint8:
                raise ValueError("%s must be uint8 (got %s)" % (what, a.dtype))
            m2, L = self._upload(torch.from_numpy(a), device), int(a.max()) + 1
        if m2.numel() != n and hw is not None and m2.dim() == 2:
            src = m2.contiguous()
            dst = torch.empty(hw[0] * hw[1], dtype=torch.uint8, device=device)
            scratch = torch.empty(hw[0] + hw[1], dtype=torch.int32, device=device)
            lib = _lib.load()
            _lib.check(lib.vst_mask_resize_nearest(src.data_ptr(), int(src.shape[0]), int(src.shape[1]), dst.data_ptr(),
                                                   int(hw[0]), int(hw[1]), scratch.data_ptr(),
                                                   torch.cuda.current_stream(device).cuda_stream),
                       "vst_mask_resize_nearest")
            m2 = dst
        m = m2.reshape(-1).contiguous()
        if m.numel() != n:
            raise ValueError("%s has %d labels but the feature map has %d positions; masks must be 2-D label maps or "
                             "flat maps at latent resolution (ref: cWCT.py:72-73)" % (what, m.numel(), n))
        return m, L

    def _upload(self, host, device):
        """Host uint8 tensor -> device, asynchronously on the current stream through a ring of 4 pinned buffers per
        size (a pageable source would make the copy synchronous)."""
        if host.is_pinned():
            return host.to(device, non_blocking=True)
        key = (host.numel(), str(device))
        ring = self._pin_ring.setdefault(key, [0, []])
        if len(ring[1]) < 4:
            ring[1].append((torch.empty(host.numel(), dtype=torch.uint8, pin_memory=True), torch.cuda.Event()))
            slot = len(ring[1]) - 1
        else:
            slot = ring[0] = (ring[0] + 1) % 4
            ring[1][slot][1].synchronize()              # the upload that last used this buffer has completed
        buf, ev = ring[1][slot]
        buf.copy_(host.reshape(-1))
        dev_t = torch.empty(host.shape, dtype=torch.uint8, device=device)
        dev_t.view(-1).copy_(buf, non_blocking=True)
        ev.record(torch.cuda.current_stream(device))
        return dev_t

    # ------------------------------------------------------------------ the reference's helper methods
    def _factor_part(self, which, stats, C_, device, stream):
        T = torch.empty(1, C_, C_, dtype=torch.float32, device=device)
        mu = torch.empty(1, C_, dtype=torch.float32, device=device)
        beta = torch.empty(1, C_, dtype=torch.float32, device=device)
        valid = torch.empty(1, dtype=torch.int32, device=device)
        status = torch.empty(1, dtype=torch.int32, device=device)
        fn = self._lib.vst_cwct_whiten_factor if which == "whiten" else self._lib.vst_cwct_color_factor
        _lib.check(fn(stats.data_ptr(), float(self.eps), C_, int(bool(self.use_double)), T.data_ptr(), mu.data_ptr(),
                      beta.data_ptr(), valid.data_ptr(), status.data_ptr(), stream), "vst_cwct_%s_factor" % which)
        self.last_status = status
        return T, mu, beta, valid

    @staticmethod
    def _as_batches(x, what):
        if not isinstance(x, torch.Tensor) or x.dim() not in (2, 3):
            raise ValueError("%s must be [C,n] or [B,C,n]" % what)
        if not x.is_cuda:
            raise RuntimeError("vstnet_b200.cWCT runs on CUDA (sm_100a) only; %s is on %s — there is no CPU fallback"
                               % (what, x.device))
        if x.dtype != torch.float32:
            raise ValueError("%s must be float32 (got %s)" % (what, x.dtype))
        xb = x.contiguous()
        return xb[None] if x.dim() == 2 else xb

    @torch.no_grad()
    def whitening(self, x):
        """``inv(chol(cov(x))) @ (x - mean(x))`` for x ``[C,n]`` (ref: cWCT.py:134-149; the fork's version is 2-D
        only, ``[B,C,n]`` is accepted here with the upstream per-sample meaning)."""
        xb = self._as_batches(x, "x")
        out = torch.empty_like(xb)
        B, C_, n = xb.shape
        with torch.cuda.device(xb.device):
            st = torch.cuda.current_stream(xb.device).cuda_stream
            for i in range(B):
                stats = self._stats(xb[i], C_, n, None, 1, st)
                T, mu, beta, valid = self._factor_part("whiten", stats, C_, xb.device, st)
                self._apply(xb[i], out[i], C_, n, None, 1, T, mu, beta, valid, st)
        return out[0] if x.dim() == 2 else out

    @torch.no_grad()
    def coloring(self, content_whiten_feat, style_feat):
        """``chol(cov(style)) @ whiten + mean(style)`` (ref: cWCT.py:152-164)."""
        wb, sb = self._as_batches(content_whiten_feat, "content_whiten_feat"), self._as_batches(style_feat, "style_feat")
        if wb.shape[:2] != sb.shape[:2]:
            raise ValueError("content and style features must agree in batch and channels")
        out = torch.empty_like(wb)
        B, C_, n = wb.shape
        with torch.cuda.device(wb.device):
            st = torch.cuda.current_stream(wb.device).cuda_stream
            for i in range(B):
                stats = self._stats(sb[i], C_, sb.shape[2], None, 1, st)
                T, mu, beta, valid = self._factor_part("color", stats, C_, wb.device, st)
                self._apply(wb[i], out[i], C_, n, None, 1, T, mu, beta, valid, st)
        return out[0] if content_whiten_feat.dim() == 2 else out

    @torch.no_grad()
    def cholesky_dec(self, conv, invert=False):
        """Cholesky factor of ``conv`` ``[C,C]`` with the cumulative ``eps*I`` retry, optionally inverted — on the
        device, without the reference's exception / host round trip per retry (ref: cWCT.py:111-132)."""
        if not isinstance(conv, torch.Tensor) or conv.dim() != 2 or conv.shape[0] != conv.shape[1]:
            raise ValueError("conv must be a square matrix [C,C]")
        if not conv.is_cuda:
            raise RuntimeError("vstnet_b200.cWCT runs on CUDA (sm_100a) only; conv is on %s" % conv.device)
        if conv.dtype not in (torch.float32, torch.float64):
            raise ValueError("conv must be float32 or float64")
        C_ = conv.shape[0]
        if C_ > 128:
            raise ValueError("vstnet_b200.cWCT supports at most 128 feature channels (got %d)" % C_)
        a = conv.contiguous()
        out = torch.empty_like(a)
        status = torch.empty(1, dtype=torch.int32, device=a.device)
        with torch.cuda.device(a.device):
            st = torch.cuda.current_stream(a.device).cuda_stream
            _lib.check(self._lib.vst_cwct_cholesky(a.data_ptr(), C_, int(a.dtype == torch.float64), float(self.eps),
                                                   int(bool(invert)), out.data_ptr(), status.data_ptr(), st),
                       "vst_cwct_cholesky")
        self.last_status = status
        return out

    # ------------------------------------------------------------------ reference API
    def transfer(self, content_feat, style_feat, cmask=None, smask=None):
        """ref: cWCT.py:18-22."""
        if cmask is None or smask is None:
            return self._transfer(content_feat, style_feat)
        return self._transfer_seg(content_feat, style_feat, cmask, smask)

    @torch.no_grad()
    def _transfer(self, content_feat, style_feat):
        """Unmasked whitening + colouring (ref: cWCT.py:24-47, intended semantics)."""
        return self.interpolation(content_feat, [style_feat], [1.0], 0.0)

    @torch.no_grad()
    def interpolation(self, content_feat, styl_feat_list, alpha_s_list, alpha_c=0.0):
        """ref: cWCT.py:206-262."""
        assert len(styl_feat_list) == len(alpha_s_list)
        _check_feat(content_feat, "content_feat")
        B, N, cH, cW = content_feat.shape
        if len(styl_feat_list) < 1 or len(styl_feat_list) > _lib.MAX_STYLES:
            raise ValueError("need 1..%d style features" % _lib.MAX_STYLES)
        for s in styl_feat_list:
            _check_feat(s, "style_feat")
            assert s.shape[0] == B and s.shape[1] == N
        if N > 128:
            raise ValueError("vstnet_b200.cWCT supports at most 128 feature channels (got %d)" % N)
        dev = content_feat.device
        content = content_feat.contiguous()
        styles = [s.contiguous() for s in styl_feat_list]
        out = torch.empty_like(content)
        n = cH * cW
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            for i in range(B):
                cst = self._stats(content[i], N, n, None, 1, st)
                sst = [self._stats(s[i], N, s.shape[2] * s.shape[3], None, 1, st) for s in styles]
                T, mu, beta, valid = self._factor(cst, sst, alpha_s_list, alpha_c, N, 1, False, dev, st)
                self._apply(content[i], out[i], N, n, None, 1, T, mu, beta, valid, st)
        return out

    # ------------------------------------------------------------------ style hoisting (video)
    @torch.no_grad()
    def precompute_style(self, style_feat, smask=None, n_labels=None):
        """Style statistics computed once per video instead of once per frame (the reference
        re-encodes and re-factorises the style every frame, video_transfer.py:195).

        Returns an opaque dict whose ``stats`` tensors (one uint8 device buffer per sample) are what
        ``vstnet_b200.video.broadcast_style_stats`` sends to the other ranks."""
        _check_feat(style_feat, "style_feat")
        B, N, sH, sW = style_feat.shape
        dev = style_feat.device
        style = style_feat.contiguous()
        ns = sH * sW
        stats, L = [], 1
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            if smask is not None:
                masks = [self._mask_to_device(smask[i], ns, dev, "smask", (sH, sW)) for i in range(B)]
                L = min(max(Li for _, Li in masks), _lib.MAX_LABELS)
                if n_labels is not None:                # a fixed number of label slots (multi-GPU: every rank
                    L = max(L, int(n_labels))           # knows the buffer size without communication)
            for i in range(B):
                stats.append(self._stats(style[i], N, ns, masks[i][0] if smask is not None else None, L, st, (sH, sW)))
        return {"stats": stats, "L": L, "masked": smask is not None, "C": N}

    @torch.no_grad()
    def transfer_precomputed(self, content_feat, style_pre, cmask=None, alpha_c=0.0, out=None):
        """``transfer`` / ``interpolation`` (one style) against hoisted style statistics."""
        _check_feat(content_feat, "content_feat")
        B, N, cH, cW = content_feat.shape
        if N != style_pre["C"] or len(style_pre["stats"]) != B:
            raise ValueError("style statistics were computed for a different shape")
        if style_pre["masked"] != (cmask is not None):
            raise ValueError("content and style masks must both be given or both be None")
        dev = content_feat.device
        content = content_feat.contiguous()
        masked, L = style_pre["masked"], style_pre["L"]
        if out is None:
            out = content if masked else torch.empty_like(content)
        n = cH * cW
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            for i in range(B):
                cm = self._mask_to_device(cmask[i], n, dev, "cmask", (cH, cW))[0] if masked else None
                cst = self._stats(content[i], N, n, cm, L, st, (cH, cW))
                T, mu, beta, valid = self._factor(cst, [style_pre["stats"][i]], [1.0], 0.0 if masked else alpha_c, N, L,
                                                  masked, dev, st)
                self._apply(content[i], out[i], N, n, cm, L, T, mu, beta, valid, st)
        return out

    @torch.no_grad()
    def _transfer_seg(self, content_feat, style_feat, cmask, smask):
        """Per-label masked transfer (ref: cWCT.py:49-109).  Mutates and returns content_feat."""
        _check_feat(content_feat, "content_feat")
        _check_feat(style_feat, "style_feat")
        B, N, cH, cW = content_feat.shape
        _, _, sH, sW = style_feat.shape
        if N > 128:
            raise ValueError("vstnet_b200.cWCT supports at most 128 feature channels (got %d)" % N)
        if not content_feat.is_contiguous():
            raise ValueError("masked transfer writes into content_feat in place (ref: cWCT.py:103); it must be "
                             "contiguous")
        if len(cmask) != B or len(smask) != B:
            raise ValueError("masks must have a leading batch dimension of %d" % B)
        dev = content_feat.device
        style = style_feat.contiguous()
        nc, ns = cH * cW, sH * sW
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            for i in range(B):
                cm, L = self._mask_to_device(cmask[i], nc, dev, "cmask", (cH, cW))
                sm, _ = self._mask_to_device(smask[i], ns, dev, "smask", (sH, sW))
                # the reference sizes its validity table max(content label)+1 (cWCT.py:173-175);
                # label 255 overflows its uint8 arithmetic there and raises IndexError (host masks only:
                # a device mask is never read back, label 255 is then an ordinary label).
                if L > 255 and not (isinstance(cmask[i], torch.Tensor) and cmask[i].is_cuda):
                    raise IndexError("content label 255 is not supported (ref: cWCT.py:173 overflows uint8)")
                cst = self._stats(content_feat[i], N, nc, cm, L, st, (cH, cW))
                sst = self._stats(style[i], N, ns, sm, L, st, (sH, sW))
                T, mu, beta, valid = self._factor(cst, [sst], [1.0], 0.0, N, L, True, dev, st)
                self._apply(content_feat[i], content_feat[i], N, nc, cm, L, T, mu, beta, valid, st)
        return content_feat
