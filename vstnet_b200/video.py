"""Frame-sharded video stylization (the multi-GPU form of the hot path).

Replaces the per-frame loop of the reference's ``video_transfer.py:160-214``.  One process per
GPU; the style image is encoded and its cWCT statistics computed ONCE (rank 0) and broadcast
with a single ``torch.distributed.broadcast`` (NCCL over NVLink); every rank then stylizes
frames ``rank, rank+W, rank+2W, ...`` with no further communication.
"""
from __future__ import annotations

import torch

from . import _lib
from .cWCT import cWCT
from .RevResNet import RevResNet


def shard_frames(n_frames, rank, world_size):
    """Frame indices owned by ``rank`` (rank-strided; every frame has exactly one owner)."""
    return list(range(rank, n_frames, world_size))


def broadcast_style_stats(pre, nbytes_of, device, group=None, src=0):
    """The path's only collective: ``src`` sends the hoisted style statistics (an opaque byte buffer per
    sample, 8.6 KB for C=32 / one label) to every rank.  ``pre`` is the dict of
    ``cWCT.precompute_style`` on ``src`` and ``None`` elsewhere; ``nbytes_of(C, L)`` sizes one buffer.
    Works on any backend (NCCL on the GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    meta = torch.zeros(4, dtype=torch.int64, device=device)
    if rank == src:
        meta[:] = torch.tensor([pre["L"], int(pre["masked"]), pre["C"], len(pre["stats"])])
    dist.broadcast(meta, src=src, group=group)
    L, masked, C_, B = (int(v) for v in meta.tolist())
    nbytes = nbytes_of(C_, L)
    stats = pre["stats"] if rank == src else [torch.empty(nbytes, dtype=torch.uint8, device=device) for _ in range(B)]
    for s in stats:
        dist.broadcast(s, src=src, group=group)
    return {"stats": stats, "L": L, "masked": bool(masked), "C": C_}


class VideoStylizer:
    def __init__(self, net: RevResNet, cwct: cWCT | None = None, alpha_c=None, n_streams=4):
        """``n_streams``: frames in flight per GPU.  Frames are independent, so consecutive frames are dealt to
        ``n_streams`` compute streams: the tail of one frame's kernels (last wave of tiles, drain) is filled by another
        frame's kernels (+5 % frames/s at 3 streams, +6 % at 4 on B200; each stream owns a ~0.6 GB workspace at 1080p)."""
        self.n_streams = max(1, int(n_streams))
        self.net = net
        self.cwct = cwct if cwct is not None else cWCT()
        self.alpha_c = alpha_c
        self.style_pre = None
        self._lib = _lib.load()
        self._pin = {}

    # ------------------------------------------------------------------ style (once per video)
    @torch.no_grad()
    def set_style(self, style=None, style_seg=None, group=None, src=0):
        """Encode the style image and hoist its statistics; broadcast from ``src`` if a process
        group is initialised.  Non-source ranks may pass ``style=None`` but must know its shape
        through the broadcast metadata."""
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        rank = dist.get_rank(group) if distributed else 0
        dev = next(self.net.parameters()).device
        if not distributed or rank == src:
            zs = self.net(style.to(dev), forward=True)
            pre = self.cwct.precompute_style(zs, style_seg)
        if distributed:
            nbytes_of = lambda C_, L: int(self._lib.vst_cwct_stats_bytes(C_, L))
            pre = broadcast_style_stats(pre if rank == src else None, nbytes_of, dev, group, src)
        self.style_pre = pre
        return pre

    # ------------------------------------------------------------------ per frame
    @torch.no_grad()
    def stylize(self, content, content_seg=None):
        """content: fp32 CUDA [1,3,H,W] in [0,1] -> stylized fp32 CUDA [1,3,H,W]
        (encode -> cWCT vs hoisted style -> decode; ref: video_transfer.py:192-206)."""
        z = self.net(content, forward=True)
        a = 0.0 if self.alpha_c is None else float(self.alpha_c)
        zcs = self.cwct.transfer_precomputed(z, self.style_pre, content_seg, a, out=z)
        return self.net(zcs, forward=False)

    @torch.no_grad()
    def stylize_frames(self, frames, content_segs=None):
        """Throughput path for a sequence of device frames (fp32 CUDA [1,3,H,W]): yields the stylized frames in
        order, ``n_streams - 1`` frames behind the input.  Frame i runs on compute stream ``i % n_streams``; every
        yielded tensor is ready on (and safe to use from) the caller's current stream."""
        dev = next(self.net.parameters()).device
        cur = torch.cuda.current_stream(dev)
        cs = self._compute_streams(dev)
        n = len(cs)
        start = torch.cuda.Event()
        start.record(cur)
        inflight = []                                   # (output, done event) in frame order
        for i, f in enumerate(frames):
            s = cs[i % n]
            if i < n:
                s.wait_event(start)                     # inputs produced on the caller's stream are complete
            with torch.cuda.stream(s):
                y = self.stylize(f, None if content_segs is None else content_segs[i])
                ev = torch.cuda.Event()
                ev.record(s)
            y.record_stream(cur)
            inflight.append((y, ev))
            if len(inflight) >= n:
                y0, e0 = inflight.pop(0)
                cur.wait_event(e0)
                yield y0
        for y0, e0 in inflight:
            cur.wait_event(e0)
            yield y0

    def _compute_streams(self, dev):
        k = ("compute", str(dev))
        if k not in self._pin or len(self._pin[k]) != self.n_streams:
            self._pin[k] = [torch.cuda.Stream(dev) for _ in range(self.n_streams)]
        return self._pin[k]

    def _pinned(self, key, shape, dtype):
        t = self._pin.get(key)
        if t is None or t.shape != torch.Size(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype).pin_memory()
            self._pin[key] = t
        return t

    @torch.no_grad()
    def stylize_host(self, frame, content_seg=None, bgr=False):
        """End-to-end call with HOST buffers.  ``frame``: pinned-or-not host tensor, either fp32
        [1,3,H,W] (what ``ToTensor`` gives, video_transfer.py:188) or uint8 [H,W,3].  Returns a
        pinned uint8 [H,W,3] host tensor (``mul(255).clamp(0,255).byte()`` semantics,
        video_transfer.py:211-214).  H2D and D2H copies are part of the call."""
        dev = next(self.net.parameters()).device
        st = torch.cuda.current_stream(dev)
        if frame.dtype == torch.uint8:
            H, W = frame.shape[0], frame.shape[1]
            u8 = frame.contiguous().to(dev, non_blocking=True)
            x = torch.empty(1, 3, H, W, dtype=torch.float32, device=dev)
            _lib.check(self._lib.vst_frame_u8_to_f32(u8.data_ptr(), x.data_ptr(), H, W, int(bgr), st.cuda_stream),
                       "vst_frame_u8_to_f32")
        else:
            H, W = frame.shape[2], frame.shape[3]
            x = frame.contiguous().to(dev, non_blocking=True)
        y = self.stylize(x, content_seg)
        o = torch.empty(H, W, 3, dtype=torch.uint8, device=dev)
        _lib.check(self._lib.vst_frame_f32_to_u8(y.data_ptr(), o.data_ptr(), H, W, int(bgr), st.cuda_stream),
                   "vst_frame_f32_to_u8")
        host = self._pinned("out", (H, W, 3), torch.uint8)
        host.copy_(o, non_blocking=True)
        st.synchronize()
        return host

    @torch.no_grad()
    def stylize_stream(self, frames, bgr=False):
        """Pipelined end-to-end path for a sequence of HOST uint8 ``[H,W,3]`` frames (what cv2 / PIL deliver;
        pinned memory makes the copies asynchronous).  Yields pinned uint8 ``[H,W,3]`` host tensors in order,
        ``n_streams`` frames behind the input: frame i is stylized on compute stream ``i % n_streams`` while later
        frames are uploaded and earlier ones downloaded on two copy streams (a ring of ``n_streams + 1`` device /
        host staging slots).  Each yielded tensor stays valid until ``n_streams`` more frames have been yielded."""
        dev = next(self.net.parameters()).device
        cs = self._compute_streams(dev)
        n = len(cs)
        R = n + 1                                                      # staging ring depth
        s_in, s_out = self._streams(dev)
        st = self._pin.get(("stream", str(dev)))
        if st is None or len(st["din"]) != R:
            st = {"din": [None] * R, "dout": [None] * R, "hout": [None] * R,
                  "ev": [[torch.cuda.Event() for _ in range(R)] for _ in range(3)]}
            self._pin[("stream", str(dev))] = st                       # staging survives across calls (pinning is slow)
        din, dout, hout = st["din"], st["dout"], st["hout"]
        ev_in, ev_done, ev_out = st["ev"]
        start = torch.cuda.Event()
        start.record(torch.cuda.current_stream(dev))
        pending = []
        for i, f in enumerate(frames):
            b = i % R
            c = cs[i % n]
            H, W = int(f.shape[0]), int(f.shape[1])
            if din[b] is None or din[b].shape != f.shape:
                din[b] = torch.empty(H, W, 3, dtype=torch.uint8, device=dev)
                dout[b] = torch.empty(H, W, 3, dtype=torch.uint8, device=dev)
                hout[b] = torch.empty(H, W, 3, dtype=torch.uint8).pin_memory()
            with torch.cuda.stream(s_in):
                if i >= R:
                    s_in.wait_event(ev_done[b])            # frame i-R no longer reads this input buffer
                din[b].copy_(f.contiguous(), non_blocking=True)
                ev_in[b].record(s_in)
            if i < n:
                c.wait_event(start)
            c.wait_event(ev_in[b])
            if i >= R:
                c.wait_event(ev_out[b])                    # frame i-R's download has left dout[b]
            with torch.cuda.stream(c):
                x = torch.empty(1, 3, H, W, dtype=torch.float32, device=dev)
                _lib.check(self._lib.vst_frame_u8_to_f32(din[b].data_ptr(), x.data_ptr(), H, W, int(bgr), c.cuda_stream),
                           "vst_frame_u8_to_f32")
                y = self.stylize(x)
                _lib.check(self._lib.vst_frame_f32_to_u8(y.data_ptr(), dout[b].data_ptr(), H, W, int(bgr), c.cuda_stream),
                           "vst_frame_f32_to_u8")
                ev_done[b].record(c)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[b])
                hout[b].copy_(dout[b], non_blocking=True)
                ev_out[b].record(s_out)
            pending.append(b)
            if len(pending) > n:
                p = pending.pop(0)
                ev_out[p].synchronize()
                yield hout[p]
        for p in pending:
            ev_out[p].synchronize()
            yield hout[p]

    def _streams(self, dev):
        k = str(dev)
        if k not in self._pin:
            self._pin[k] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        return self._pin[k]

