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
        self._ws = {}
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
        params = list(self.state_dict(keep_vars=True).values())
        key = (str(device), self._precision) + tuple((p.data_ptr(), p._version) for p in params)
        if key != self._packed_key:
            for p in params:
                if p.device != device or p.dtype != torch.float32:
                    raise RuntimeError("RevResNet parameters must be float32 on %s (got %s on %s); call .to(device)"
                                       % (device, p.dtype, p.device))
            flat = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
            packed = torch.empty(int(self._lib.vst_revnet_packed_bytes(self._h)) + 16, dtype=torch.uint8, device=device)
            st = torch.cuda.current_stream(device)
            _lib.check(self._lib.vst_revnet_pack_weights(self._h, flat.data_ptr(), packed.data_ptr(), st.cuda_stream),
                       "vst_revnet_pack_weights")
            flat.record_stream(st)
            self._packed, self._packed_key = packed, key
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
        with torch.cuda.device(x.device):
            packed, ws = self._packed_weights(x.device), self._workspace(x.device, B, H, W)
            st = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(self._lib.vst_revnet_forward(self._h, packed.data_ptr(), x.data_ptr(), z.data_ptr(), B, H, W,
                                                    ws.data_ptr(), ws.numel(), st), "vst_revnet_forward")
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
        with torch.cuda.device(z.device):
            packed, ws = self._packed_weights(z.device), self._workspace(z.device, B, H, W)
            st = torch.cuda.current_stream(z.device).cuda_stream
            _lib.check(self._lib.vst_revnet_inverse(self._h, packed.data_ptr(), z.data_ptr(), x.data_ptr(), B, H, W,
                                                    ws.data_ptr(), ws.numel(), st), "vst_revnet_inverse")
        return x
