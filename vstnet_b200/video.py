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


STYLE_HEADER_BYTES = 32          # 4 x int64: L, masked, C, B


def pack_style_stats(pre):
    """``precompute_style`` dict -> ONE uint8 device buffer: [L, masked, C, B as int64 | stats of sample 0 | ...]."""
    head = torch.tensor([pre["L"], int(pre["masked"]), pre["C"], len(pre["stats"])], dtype=torch.int64)
    dev = pre["stats"][0].device
    return torch.cat([head.view(torch.uint8).to(dev)] + [s.reshape(-1) for s in pre["stats"]])


def unpack_style_stats(buf):
    head = buf[:STYLE_HEADER_BYTES].cpu().view(torch.int64).tolist()
    L, masked, C_, B = (int(v) for v in head)
    per = (buf.numel() - STYLE_HEADER_BYTES) // max(B, 1)
    stats = [buf[STYLE_HEADER_BYTES + i * per:STYLE_HEADER_BYTES + (i + 1) * per] for i in range(B)]
    return {"stats": stats, "L": L, "masked": bool(masked), "C": C_}


def broadcast_style_stats(pre, nbytes, device, group=None, src=0):
    """The path's only collective, exactly ONE ``broadcast``: ``src`` sends the hoisted style statistics
    (header + one opaque stats block per sample; 8.6 KB for C = 32 / one label) to every rank.  ``pre`` is the
    dict of ``cWCT.precompute_style`` on ``src`` and ``None`` elsewhere; ``nbytes`` is the buffer size, which every
    rank can compute without communication when the style layout is known up front (``style_buffer_bytes``), as it
    is for the unmasked video path.  Works on any backend (NCCL on the GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    buf = pack_style_stats(pre) if rank == src else torch.empty(nbytes, dtype=torch.uint8, device=device)
    if buf.numel() != nbytes:
        raise ValueError("style statistics occupy %d bytes, the ranks agreed on %d" % (buf.numel(), nbytes))
    dist.broadcast(buf, src=src, group=group)
    return unpack_style_stats(buf)


def style_buffer_bytes(nbytes_of, C_, L=1, B=1):
    """Size of the broadcast buffer for B samples of a C-channel latent with L label slots."""
    return STYLE_HEADER_BYTES + B * int(nbytes_of(C_, L))


class SharedFrameRing:
    """Ordered delivery of stylized uint8 frames from the ranks of ONE node to a single consumer (the video writer on
    rank 0) through a bounded ring in shared host memory — a file-backed mapping under ``/dev/shm`` — instead of
    pickling every frame through ``dist.gather_object`` (at 700 frames/s x 6.2 MB that host path, not the GPUs,
    would bound an 8-GPU run).  No device collective and no serialisation: a producer copies its frame into slot
    ``i % slots`` and publishes ``ready[slot] = i + 1``; the consumer takes frames 0, 1, 2, ... and publishes
    ``consumed``; a producer waits while ``i - slots >= consumed``.  x86 keeps the frame bytes ahead of the flag."""

    HEADER = 4096          # int64 words: [0] consumed, [8 + slot] ready

    def __init__(self, path, frame_shape, slots, create):
        import numpy as np
        self.path, self.slots = path, int(slots)
        self.frame_shape = tuple(int(v) for v in frame_shape)
        self.frame_bytes = int(np.prod(self.frame_shape))
        if 8 + self.slots > self.HEADER // 8:
            raise ValueError("at most %d slots" % (self.HEADER // 8 - 8))
        total = self.HEADER + self.slots * self.frame_bytes
        if create:
            with open(path, "wb") as f:
                f.truncate(total)
        self._mm = np.memmap(path, dtype=np.uint8, mode="r+", shape=(total,))
        self._flags = self._mm[:self.HEADER].view(np.int64)
        self._frames = self._mm[self.HEADER:].reshape((self.slots,) + self.frame_shape)

    @staticmethod
    def default_path(tag):
        import os
        base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else "/tmp"
        return os.path.join(base, "vstb200_frames_%s" % tag)

    def put(self, i, frame, timeout=600.0):
        import time
        t0 = time.monotonic()
        while i - self.slots >= int(self._flags[0]):
            if time.monotonic() - t0 > timeout:
                raise TimeoutError("frame %d: the consumer has not freed slot %d" % (i, i % self.slots))
            time.sleep(0.0002)
        k = i % self.slots
        self._frames[k][...] = frame.numpy() if hasattr(frame, "numpy") else frame
        self._flags[8 + k] = i + 1

    def get(self, i, timeout=600.0):
        """Frame i (a view into the ring, valid until ``release(i)``)."""
        import time
        t0 = time.monotonic()
        k = i % self.slots
        while int(self._flags[8 + k]) != i + 1:
            if time.monotonic() - t0 > timeout:
                raise TimeoutError("frame %d was never delivered" % i)
            time.sleep(0.0002)
        return self._frames[k]

    def release(self, i):
        self._flags[0] = i + 1

    def close(self, unlink=False):
        import os
        self._frames = self._flags = None
        mm, self._mm = self._mm, None
        del mm
        if unlink:
            try:
                os.unlink(self.path)
            except OSError:
                pass


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
        self.fused = True               # unmasked frames take the fused native call (vst_revnet_stylize)
        self._lib = _lib.load()
        self._pin = {}

    # ------------------------------------------------------------------ style (once per video)
    @torch.no_grad()
    def set_style(self, style=None, style_seg=None, group=None, src=0, masked=False):
        """Encode the style image and hoist its statistics; ONE broadcast from ``src`` if a process group is
        initialised.  Non-source ranks pass ``style=None`` (and ``masked=True`` if ``src`` has a ``style_seg``)."""
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        rank = dist.get_rank(group) if distributed else 0
        dev = next(self.net.parameters()).device
        with torch.cuda.device(dev):
            self.net._packed_weights(dev)          # every rank packs here, on the caller's stream, before any compute
                                                   # stream is forked (non-source ranks never run the net in set_style)
        # multi-GPU: the ranks agree on the buffer size without communication — C from the network, L = 1 unmasked,
        # L = 256 (every uint8 label has a slot) masked
        masked = bool(masked or style_seg is not None)
        L = None if not distributed else (_lib.MAX_LABELS if masked else 1)
        if not distributed or rank == src:
            zs = self.net(style.to(dev), forward=True)
            pre = self.cwct.precompute_style(zs, style_seg, n_labels=L if masked else None)
        if distributed:
            nbytes = style_buffer_bytes(self._lib.vst_cwct_stats_bytes, self.net.latent_channels, L, 1)
            pre = broadcast_style_stats(pre if rank == src else None, nbytes, dev, group, src)
        self.style_pre = pre
        return pre

    # ------------------------------------------------------------------ per frame
    @torch.no_grad()
    def stylize(self, content, content_seg=None):
        """content: fp32 CUDA [1,3,H,W] in [0,1] -> stylized fp32 CUDA [1,3,H,W]
        (encode -> cWCT vs hoisted style -> decode; ref: video_transfer.py:192-206)."""
        a = 0.0 if self.alpha_c is None else float(self.alpha_c)
        if self._fused_ok(content, content_seg):
            # one native call, the latent never leaves the network's state (vst_revnet_stylize)
            return self.net.stylize_frame(content, self.style_pre["stats"][0], a, self.cwct.eps, self.cwct.use_double)
        z = self.net(content, forward=True)
        zcs = self.cwct.transfer_precomputed(z, self.style_pre, content_seg, a, out=z)
        return self.net(zcs, forward=False)

    def _fused_ok(self, frame, content_seg=None):
        if not self.fused or content_seg is not None or self.style_pre["masked"] or len(self.style_pre["stats"]) != 1:
            return False
        if frame.dtype == torch.uint8:
            return frame.dim() == 3 and self.net.stylize_supported(frame.shape[0], frame.shape[1])
        return frame.dim() == 4 and frame.shape[0] == 1 and self.net.stylize_supported(frame.shape[2], frame.shape[3])

    @torch.no_grad()
    def stylize_frames(self, frames, content_segs=None):
        """Throughput path for a sequence of device frames (fp32 CUDA [1,3,H,W], or uint8 CUDA [H,W,3] on the fused
        path): yields the stylized frames in order, ``n_streams - 1`` frames behind the input.  Frame i runs on compute
        stream ``i % n_streams``; every yielded tensor is ready on (and safe to use from) the caller's current stream.
        On the fused path the outputs live in a ring of ``2 n_streams + 2`` reused buffers (no allocator traffic in the
        steady state): a yielded tensor stays valid until ``n_streams + 2`` more frames have been yielded — clone it
        if it must live longer."""
        dev = next(self.net.parameters()).device
        cur = torch.cuda.current_stream(dev)
        cs = self._compute_streams(dev)
        n = len(cs)
        inflight = []                                   # (output, done event) in frame order
        reuse, yielded = [], 0                          # ring slot per frame (fused path), frames handed out so far
        for i, f in enumerate(frames):
            s = cs[i % n]
            ev_in = torch.cuda.Event()                  # the frame may be produced lazily on the caller's stream (a
            ev_in.record(cur)                           # generator running device work): order it, and everything
            s.wait_event(ev_in)                         # enqueued before, ahead of its compute stream
            f.record_stream(s)
            seg = None if content_segs is None else content_segs[i]
            with torch.cuda.stream(s):
                if self._fused_ok(f, seg):
                    ring = self._out_ring(f, 2 * n + 2)
                    slot = ring[i % len(ring)]
                    s.wait_event(slot[1])                # the consumer's use of this buffer (recorded at yield time) is over
                    a = 0.0 if self.alpha_c is None else float(self.alpha_c)
                    y = self.net.stylize_frame(f, self.style_pre["stats"][0], a, self.cwct.eps, self.cwct.use_double, out=slot[0])
                    reuse.append(slot)
                else:
                    y = self.stylize(f, seg)
                    y.record_stream(cur)
                    reuse.append(None)
                ev = torch.cuda.Event()
                ev.record(s)
            inflight.append((y, ev))
            if len(inflight) >= n:
                y0, e0 = inflight.pop(0)
                cur.wait_event(e0)
                self._mark_consumed(reuse, yielded, cur)
                yielded += 1
                yield y0
        for y0, e0 in inflight:
            cur.wait_event(e0)
            self._mark_consumed(reuse, yielded, cur)
            yielded += 1
            yield y0

    @staticmethod
    def _mark_consumed(reuse, k, cur):
        """Frame k is being handed to the consumer: everything the consumer enqueued on `cur` for the frames yielded
        before it is recorded into their ring slots, so that the slot's next producer waits for it."""
        for j in (k - 1,):
            if j >= 0 and reuse[j] is not None:
                reuse[j][1].record(cur)

    def _out_ring(self, like, depth):
        k = ("outring", str(like.device), tuple(like.shape), like.dtype)
        ring = self._pin.get(k)
        if ring is None or len(ring) != depth:
            ring = [(torch.empty_like(like), torch.cuda.Event()) for _ in range(depth)]
            self._pin[k] = ring
        return ring

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
        if frame.dtype == torch.uint8 and content_seg is None:
            u8 = frame.contiguous().to(dev, non_blocking=True)
            if self._fused_ok(u8):
                a = 0.0 if self.alpha_c is None else float(self.alpha_c)
                o = self.net.stylize_frame(u8, self.style_pre["stats"][0], a, self.cwct.eps, self.cwct.use_double, bgr=bgr)
                host = self._pinned("out", tuple(o.shape), torch.uint8)
                host.copy_(o, non_blocking=True)
                st.synchronize()
                return host
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
        frames are uploaded and earlier ones downloaded on two copy streams (a ring of ``n_streams + 1`` device
        staging slots and ``2 n_streams + 2`` host slots).  A yielded tensor is a view of a ring slot: it stays valid
        until ``n_streams`` more frames have been yielded (the host ring is deep enough that no download in flight
        targets a slot the consumer may still hold) — copy it if it must live longer."""
        dev = next(self.net.parameters()).device
        cs = self._compute_streams(dev)
        n = len(cs)
        R = n + 1                                                      # device staging ring depth
        RH = 2 * R                                                     # host ring: frame j's slot is re-targeted by the
                                                                       # download of frame j + RH, enqueued only after
                                                                       # frame j + RH - n - 1 >= j + n + 1 was yielded
        s_in, s_out = self._streams(dev)
        st = self._pin.get(("stream", str(dev)))
        if st is None or len(st["din"]) != R:
            st = {"din": [None] * R, "dout": [None] * R, "hout": [None] * RH,
                  "ev": [[torch.cuda.Event() for _ in range(R)] for _ in range(2)],
                  "ev_out": [torch.cuda.Event() for _ in range(RH)]}
            self._pin[("stream", str(dev))] = st                       # staging survives across calls (pinning is slow)
        din, dout, hout = st["din"], st["dout"], st["hout"]
        ev_in, ev_done = st["ev"]
        ev_out = st["ev_out"]
        start = torch.cuda.Event()
        start.record(torch.cuda.current_stream(dev))
        pending = []
        for i, f in enumerate(frames):
            b = i % R
            hb = i % RH
            c = cs[i % n]
            H, W = int(f.shape[0]), int(f.shape[1])
            if din[b] is None or din[b].shape != f.shape:
                din[b] = torch.empty(H, W, 3, dtype=torch.uint8, device=dev)
                dout[b] = torch.empty(H, W, 3, dtype=torch.uint8, device=dev)
            if hout[hb] is None or hout[hb].shape != f.shape:
                hout[hb] = torch.empty(H, W, 3, dtype=torch.uint8).pin_memory()
            with torch.cuda.stream(s_in):
                if i >= R:
                    s_in.wait_event(ev_done[b])            # frame i-R no longer reads this input buffer
                din[b].copy_(f.contiguous(), non_blocking=True)
                ev_in[b].record(s_in)
            if i < n:
                c.wait_event(start)
            c.wait_event(ev_in[b])
            if i >= R:
                c.wait_event(ev_out[(i - R) % RH])         # frame i-R's download has left dout[b]
            with torch.cuda.stream(c):
                if self._fused_ok(din[b]):
                    # uint8 frame in, uint8 frame out: the format conversions live in the first / last kernel of the pass
                    a = 0.0 if self.alpha_c is None else float(self.alpha_c)
                    self.net.stylize_frame(din[b], self.style_pre["stats"][0], a, self.cwct.eps, self.cwct.use_double,
                                           out=dout[b], bgr=bgr)
                else:
                    x = torch.empty(1, 3, H, W, dtype=torch.float32, device=dev)
                    _lib.check(self._lib.vst_frame_u8_to_f32(din[b].data_ptr(), x.data_ptr(), H, W, int(bgr), c.cuda_stream),
                               "vst_frame_u8_to_f32")
                    y = self.stylize(x)
                    _lib.check(self._lib.vst_frame_f32_to_u8(y.data_ptr(), dout[b].data_ptr(), H, W, int(bgr), c.cuda_stream),
                               "vst_frame_f32_to_u8")
                ev_done[b].record(c)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[b])
                hout[hb].copy_(dout[b], non_blocking=True)
                ev_out[hb].record(s_out)
            pending.append(hb)
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

