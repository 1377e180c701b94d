"""Drop-in replacement for the reference's ``models/RevResNet.py`` on B200.

Same constructor arguments, parameter names (``stack.{i}.conv.{1,4,7}.{weight,bias}``,
``channel_reduction.block_list.{j}.conv.{1,4,7}.{weight,bias}``) and call surface
(``net(x, forward=True|False)``, ``_forward``, ``_inverse``, ``.down_scale`` ...), so a
reference checkpoint loads with ``load_state_dict`` unchanged (ref: RevResNet.py:166-239,
image_transfer.py:44-55).  The modules below only *hold parameters*; all arithmetic runs in
hand-written sm_100a CUDA behind the C ABI of ``libvstb200.so`` (include/vstb200.h).
There is no CPU or eager-PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib


class residual_block(nn.Module):
    """Parameter container for one additive-coupling block (ref: RevResNet.py:68-94).

    ``conv`` is an ``nn.Sequential`` whose entries 1, 4 and 7 are the three 3x3 convolutions, so
    the state_dict keys (and the RNG draws of the default init) equal the reference's.
    """

    def __init__(self, channel, stride=1, mult=4, kernel=3):
        super().__init__()
        if kernel != 3:
            raise ValueError("vstnet_b200 implements the reference's 3x3 kernels only")
        self.stride = stride
        in_ch = channel if stride == 1 else channel // 4
        mid = channel // mult
        slots = [nn.Identity() for _ in range(8)]
        slots[1] = nn.Conv2d(in_ch, mid, kernel_size=3, stride=stride, padding=0, bias=True)
        slots[4] = nn.Conv2d(mid, mid, kernel_size=3, padding=0, bias=True)
        slots[7] = nn.Conv2d(mid, channel, kernel_size=3, padding=0, bias=True)
        self.conv = nn.Sequential(*slots)
        for k in (1, 4, 7):
            self.conv[k].bias.data.zero_()          # ref: RevResNet.py:91-94

    def forward(self, x):
        raise RuntimeError("residual_block is a parameter container; call RevResNet(x, forward=...)")


class channel_reduction(nn.Module):
    """Parameter container (ref: RevResNet.py:119-129)."""

    def __init__(self, in_ch, out_ch, sp_steps=2, n_blocks=2, kernel=3):
        super().__init__()
        self.pad = out_ch * 4 ** sp_steps - in_ch
        self.sp_steps = sp_steps
        self.n_blocks = n_blocks
        self.block_list = nn.ModuleList(
            [residual_block(out_ch * 4 ** sp_steps, stride=1, mult=4, kernel=kernel) for _ in range(n_blocks)])

    def forward(self, x):
        raise RuntimeError("channel_reduction is a parameter container; call RevResNet(x, forward=...)")


class RevResNet(nn.Module):
    def __init__(self, nBlocks=[10, 10, 10], nStrides=[1, 2, 2], nChannels=[16, 64, 256], in_channel=3, mult=4,
                 hidden_dim=16, sp_steps=2, kernel=3, precision="f16x2"):
        super().__init__()
        if not nChannels:
            nChannels = [in_channel * 2, in_channel * 2 * 4, in_channel * 2 * 4 ** 2]
        self.nBlocks = list(nBlocks)
        self.nStrides = list(nStrides)
        self.nChannels = list(nChannels)
        self.in_channel = in_channel
        self.mult = mult
        self.hidden_dim = hidden_dim
        self.sp_steps = sp_steps
        self.pad = 2 * nChannels[0] - in_channel
        self.in_ch = nChannels[0]
        self.down_scale = np.prod(np.array(nStrides))

        blocks = []
        for channel, depth, stride in zip(nChannels, nBlocks, nStrides):
            for i in range(depth):
                blocks.append(residual_block(channel, stride if i == 0 else 1, mult=mult, kernel=kernel))
        self.stack = nn.ModuleList(blocks)
        self.channel_reduction = channel_reduction(nChannels[-1], hidden_dim, sp_steps=sp_steps, kernel=kernel)

        # ---- native plan (host-only object; owns no device memory)
        lib = _lib.load()
        cfg = _lib.RevnetConfig()
        if len(nBlocks) > _lib.MAX_STAGES:
            raise ValueError("too many stages")
        cfg.n_stages = len(nBlocks)
        for i in range(len(nBlocks)):
            cfg.n_blocks[i], cfg.n_strides[i], cfg.n_channels[i] = int(nBlocks[i]), int(nStrides[i]), int(nChannels[i])
        cfg.in_channel, cfg.mult, cfg.hidden_dim, cfg.sp_steps = in_channel, mult, hidden_dim, sp_steps
        cfg.n_cr_blocks = self.channel_reduction.n_blocks
        h = C.c_void_p()
        rc = lib.vst_revnet_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise ValueError("unsupported RevResNet configuration: %s" % lib.vst_last_error().decode())
        self._h = h
        self._lib = lib
        self.precision = precision
        self._packed = None
        self._packed_key = None
        self._packed_event = None
        self._packed_streams = set()
        self._ws = {}
        # fp16 operand range guard of the f16x2 arithmetic (|activation| must stay below 1023.75; the reference's
        # trained checkpoint produces |z| ~ 1.2, image_transfer.py:183-205).  The kernels raise a bit in the status word
        # at the head of the workspace; "deferred" reads it back asynchronously and, when a LATER call finds it set,
        # warns and switches this module to tf32x2 (fp32 exponent range) for good; "strict" synchronises after every
        # call and transparently re-runs the flagged call in tf32x2; "off" skips the read-back.
        self.range_check = "deferred"
        self._range_pending = []
        n_params = sum(p.numel() for p in self.parameters())
        assert n_params == lib.vst_revnet_param_floats(h), "parameter layout mismatch with the native plan"

    # ------------------------------------------------------------------ properties
    @property
    def precision(self):
        return self._precision

    @precision.setter
    def precision(self, name):
        if name not in _lib.PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(_lib.PRECISIONS))
        _lib.check(self._lib.vst_revnet_set_precision(self._h, _lib.PRECISIONS[name]), "vst_revnet_set_precision")
        self._precision = name
        self._packed_key = None

    @property
    def latent_channels(self):
        return 2 * self.hidden_dim

    def __del__(self):
        try:
            h = self.__dict__.pop("_h", None)
            if h:
                self.__dict__["_lib"].vst_revnet_destroy(h)
        except Exception:
            pass

    # ------------------------------------------------------------------ device-side state
    def _packed_weights(self, device):
        """The weights repacked into the kernels' layouts (device buffer, cached until a parameter changes).

        Packing is asynchronous on the stream that first needs it; an event recorded behind the pack kernels is
        waited on by every OTHER stream that uses the buffer (encode_pair's side stream, VideoStylizer's compute
        streams), and the buffer is ``record_stream``'d for them so the allocator never recycles it under a reader."""
        params = list(self.state_dict(keep_vars=True).values())
        key = (str(device), self._precision) + tuple((p.data_ptr(), p._version) for p in params)
        st = torch.cuda.current_stream(device)
        if key != self._packed_key:
            for p in params:
                if p.device != device or p.dtype != torch.float32:
                    raise RuntimeError("RevResNet parameters must be float32 on %s (got %s on %s); call .to(device)"
                                       % (device, p.dtype, p.device))
            flat = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
            packed = torch.empty(int(self._lib.vst_revnet_packed_bytes(self._h)) + 16, dtype=torch.uint8, device=device)
            _lib.check(self._lib.vst_revnet_pack_weights(self._h, flat.data_ptr(), packed.data_ptr(), st.cuda_stream),
                       "vst_revnet_pack_weights")
            flat.record_stream(st)
            ev = torch.cuda.Event()
            ev.record(st)
            self._packed, self._packed_key = packed, key
            self._packed_event, self._packed_streams = ev, {st.cuda_stream}
        elif st.cuda_stream not in self._packed_streams:
            st.wait_event(self._packed_event)          # ordered behind the pack kernels of the packing stream
            self._packed.record_stream(st)
            self._packed_streams.add(st.cuda_stream)
        return self._packed

    def _workspace(self, device, B, H, W):
        need = int(self._lib.vst_revnet_workspace_bytes(self._h, B, H, W))
        st = torch.cuda.current_stream(device)
        k = (str(device), st.cuda_stream)
        ws = self._ws.get(k)
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=device)
            self._ws[k] = ws
        return ws

    @staticmethod
    def _check_input(x, what):
        if not isinstance(x, torch.Tensor) or x.dim() != 4:
            raise ValueError("%s must be a 4-D tensor [B,C,H,W]" % what)
        if not x.is_cuda:
            raise RuntimeError("vstnet_b200.RevResNet runs on CUDA (sm_100a) only; %s is on %s — there is no CPU "
                               "fallback" % (what, x.device))
        if x.dtype != torch.float32:
            raise ValueError("%s must be float32 (got %s)" % (what, x.dtype))

    # ------------------------------------------------------------------ API
    def forward(self, x, forward=True):
        return self._forward(x) if forward else self._inverse(x)

    def inverse(self, z):
        return self._inverse(z)

    @torch.no_grad()
    def encode_pair(self, a, b):
        """``(self(a), self(b))`` with the two independent encodes on two CUDA streams (the content and the style image
        of image_transfer.py:181-182).  Small images leave most SMs idle in the deep stages (a 512x512 image has 64 conv
        tiles for 148 SMs); the second encode fills them.  Results are bit-identical to the sequential calls and are
        ready on the caller's current stream."""
        self._check_input(a, "a")
        self._check_input(b, "b")
        dev = a.device
        cur = torch.cuda.current_stream(dev)
        with torch.cuda.device(dev):
            self._packed_weights(dev)                  # packed on `cur`, before the side stream forks from it
        side = self.__dict__.get("_side_stream")
        if side is None or side.device != dev:
            side = torch.cuda.Stream(dev)
            self.__dict__["_side_stream"] = side
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            zb = self._forward(b)
        za = self._forward(a)
        cur.wait_stream(side)
        zb.record_stream(cur)
        return za, zb

    # ------------------------------------------------------------------ launch + range guard
    def _range_fallback(self, when):
        import warnings
        warnings.warn("vstnet_b200.RevResNet: an activation exceeded the fp16 operand range of precision 'f16x2' "
                      "(|x| >= 1023.75) %s; switching this module to precision 'tf32x2'" % when, RuntimeWarning)
        self.precision = "tf32x2"

    def _poll_range(self, block=False):
        keep, hit = [], False
        for ev, host, fused in self._range_pending:
            if block:
                ev.synchronize()
            if ev.query():
                hit = hit or bool(int(host[0]) & 1)
                if fused and int(host[1]) < 0:   # fused frame path: the cWCT factorisation of that frame failed
                    import warnings
                    warnings.warn("vstnet_b200: a Cholesky factorisation did not succeed within 64 jitter retries; that "
                                  "frame was returned unstylized", RuntimeWarning)
            else:
                keep.append((ev, host, fused))
        self._range_pending = keep
        return hit

    def check_status(self):
        """Synchronise on the status words of the earlier calls; True if one of them left the f16x2 range (the module
        has then been switched to tf32x2)."""
        hit = self._poll_range(block=True)
        if hit and self._precision == "f16x2":
            self._range_fallback("in an earlier call, whose result is invalid")
        return hit

    def _launch(self, fn, what, src, dst, B, H, W, call=None):
        """Enqueue one pass (``fn`` = forward / inverse, or ``call(packed, ws, stream)`` for the fused frame path) on the
        current stream and run the f16x2 range guard around it."""
        dev = src.device
        with torch.cuda.device(dev):
            while True:
                packed, ws = self._packed_weights(dev), self._workspace(dev, B, H, W)
                st = torch.cuda.current_stream(dev)
                if call is not None:
                    _lib.check(call(packed, ws, st.cuda_stream), what)
                else:
                    _lib.check(fn(self._h, packed.data_ptr(), src.data_ptr(), dst.data_ptr(), B, H, W, ws.data_ptr(),
                                  ws.numel(), st.cuda_stream), what)
                if (self._precision != "f16x2" and call is None) or self.range_check == "off":
                    return
                if self._poll_range() and self._precision == "f16x2":
                    self._range_fallback("in an earlier call, whose result is invalid")
                    return
                host = torch.empty(4, dtype=torch.int32, pin_memory=True)
                host.copy_(ws[:16].view(torch.int32), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(st)
                if self.range_check == "strict" and self._precision == "f16x2":
                    ev.synchronize()
                    if int(host[0]) & 1:
                        self._range_fallback("in this call; re-running it")
                        continue
                    return
                if len(self._range_pending) >= 16:
                    self._range_pending.pop(0)
                self._range_pending.append((ev, host, call is not None))
                return

    # ------------------------------------------------------------------ fused frame path (video)
    def stylize_supported(self, H, W):
        """True if ``stylize_frame`` can take an H x W frame (both reference modes; latent width >= 32)."""
        return bool(self._lib.vst_revnet_stylize_supported(self._h, int(H), int(W)))

    @torch.no_grad()
    def stylize_frame(self, frame, style_stats, alpha_c=0.0, eps=2e-5, use_double=False, out=None, bgr=False):
        """encode -> cWCT against hoisted style statistics -> decode as ONE native call that never materialises the
        latent (``vst_revnet_stylize``): the statistics are taken from, and the transform applied to, the network's own
        state.  ``frame``: fp32 CUDA ``[1,3,H,W]`` (returns fp32 ``[1,3,H,W]``) or uint8 CUDA ``[H,W,3]`` (returns uint8
        ``[H,W,3]``, ``mul(255).clamp(0,255).byte()`` semantics).  ``style_stats``: a buffer of
        ``cWCT.precompute_style(...)["stats"]``.  Equals ``self(cwct.transfer(self(frame), style), forward=False)`` up
        to the summation order of the statistics (ref: video_transfer.py:192-206)."""
        u8 = frame.dtype == torch.uint8
        if not frame.is_cuda:
            raise RuntimeError("vstnet_b200.RevResNet runs on CUDA (sm_100a) only; frame is on %s" % frame.device)
        if u8:
            if frame.dim() != 3 or frame.shape[2] != 3:
                raise ValueError("uint8 frames must be [H,W,3]")
            H, W = int(frame.shape[0]), int(frame.shape[1])
        else:
            self._check_input(frame, "frame")
            if frame.shape[0] != 1 or frame.shape[1] != 3:
                raise ValueError("fp32 frames must be [1,3,H,W]")
            H, W = int(frame.shape[2]), int(frame.shape[3])
        if not self.stylize_supported(H, W):
            raise ValueError("the fused frame path does not support this configuration / size (%dx%d)" % (H, W))
        frame = frame.contiguous()
        if out is None:
            out = torch.empty_like(frame)
        elif out.shape != frame.shape or out.dtype != frame.dtype or not out.is_contiguous():
            raise ValueError("out must be a contiguous tensor like frame")
        lib, h = self._lib, self._h

        def call(packed, ws, stream):
            return lib.vst_revnet_stylize(h, packed.data_ptr(), frame.data_ptr(), out.data_ptr(), int(u8), int(bool(bgr)), H, W,
                                          style_stats.data_ptr(), float(alpha_c), float(eps), int(bool(use_double)),
                                          ws.data_ptr(), ws.numel(), stream)
        self._launch(None, "vst_revnet_stylize", frame, out, 1, H, W, call=call)
        return out

    @torch.no_grad()
    def _forward(self, x):
        """Encode (ref: RevResNet.py:210-223)."""
        self._check_input(x, "x")
        B, Cin, H, W = x.shape
        ds = int(self.down_scale)
        if Cin != self.in_channel:
            raise ValueError("expected %d input channels, got %d" % (self.in_channel, Cin))
        if H % ds or W % ds or H // ds < 2 or W // ds < 2:
            raise ValueError("H and W must be multiples of %d and at least %d (got %dx%d)" % (ds, 2 * ds, H, W))
        x = x.contiguous()
        f = 2 ** self.sp_steps
        z = torch.empty(B, self.latent_channels, H // ds * f, W // ds * f, dtype=torch.float32, device=x.device)
        self._launch(self._lib.vst_revnet_forward, "vst_revnet_forward", x, z, B, H, W)
        return z

    @torch.no_grad()
    def _inverse(self, z):
        """Decode (ref: RevResNet.py:225-239)."""
        self._check_input(z, "z")
        B, Cz, h, w = z.shape
        ds, f = int(self.down_scale), 2 ** self.sp_steps
        if Cz != self.latent_channels or h % f or w % f:
            raise ValueError("latent must be [B,%d,k*%d,k*%d] (got %s)" % (self.latent_channels, f, f, tuple(z.shape)))
        H, W = h // f * ds, w // f * ds
        z = z.contiguous()
        x = torch.empty(B, self.in_channel, H, W, dtype=torch.float32, device=z.device)
        self._launch(self._lib.vst_revnet_inverse, "vst_revnet_inverse", z, x, B, H, W)
        return x
