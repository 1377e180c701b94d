#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native CAP-VSTNet stylization hot path.

Workload (BASELINE.json configs[3], the one `metric` is quoted on): photorealistic 1080p video,
random-init RevResNet (seed 0), synthetic frames, the style image encoded and its cWCT statistics
hoisted once (rank 0) and NCCL-broadcast; every rank then stylizes its own frames.  A *step* is
one 1920x1080 frame per rank: encode -> cWCT stats/factor/apply -> decode.  `value` is whole-job
frames/s with frames resident in HBM; `e2e` is the same through `VideoStylizer.stylize_stream`
with HOST buffers (pinned uint8 HWC frame H2D + uint8 result D2H inside the timed region, pipelined).
Extras: `images` (N = 1) — ms per image of the other BASELINE configs (cfg1/2/3/5) with the CPU oracle
timed beside them (cfg1/2/3) and the parity figures of the same run; `sustained` — 240 frames back to
back; `strong` — the whole 240-frame job (style encode + broadcast + frames + ordered delivery).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nproc-per-node N ... bench.py --gpus N ...

`--impl reference` times the reference's CPU implementation of the same path (oracle/vst_oracle.py, the
torch CPU ops the reference issues) and imports nothing of the product.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W = 1080, 1920
FRAMES_PER_VIDEO = 240
CONV_FLOP_PER_PX = 609984.0      # 96 convs per pass, 2*MAC
METRIC = "1080p video frames/s"
UNIT = "frames/s"
WORKLOAD = "cfg4 photorealistic video 1920x1080, style hoisted + broadcast, random-init RevResNet"
MODES = {"photo": dict(hidden_dim=16, sp_steps=2), "art": dict(hidden_dim=64, sp_steps=1)}


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(f):
        try:
            d = json.load(open(f))
            p.update({k: float(d[k]) for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained") if k in d})
            p["source"] = "measured"
        except Exception:
            pass
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# synthetic inputs of the BASELINE configs (shared with tests/test_baseline_configs.py)
# ----------------------------------------------------------------------------------------------
def blocky_mask(h, w, gy, gx, perm):
    import numpy as np
    m = np.zeros((h, w), np.uint8)
    ys, xs = np.linspace(0, h, gy + 1).astype(int), np.linspace(0, w, gx + 1).astype(int)
    for i in range(gy):
        for j in range(gx):
            m[ys[i]:ys[i + 1], xs[j]:xs[j + 1]] = perm[i * gx + j]
    return m[None]


IMAGE_CONFIGS = {
    # name: (mode, H = W, alpha_c, masked)
    "cfg1_photo_512": ("photo", 512, None, False),
    "cfg2_art_1024_alpha0.5": ("art", 1024, 0.5, False),
    "cfg3_photo_1024_masked8": ("photo", 1024, None, True),
    "cfg5_art_4096": ("art", 4096, None, False),
}


def image_config_inputs(name, size=None):
    """CPU tensors of one BASELINE image config: (mode, content, style, alpha_c, cmask, smask).  SURVEY.md 8(d):
    torch.rand images (seed 7), 8 blocky labels (2x4 content grid, permuted 4x2 style grid)."""
    import torch
    mode, s0, alpha, masked = IMAGE_CONFIGS[name]
    s = int(size or s0)
    g = torch.Generator().manual_seed(7)
    c = torch.rand(1, 3, s, s, generator=g)
    st = torch.rand(1, 3, s, s, generator=g)
    cm = blocky_mask(s, s, 2, 4, [0, 1, 2, 3, 4, 5, 6, 7]) if masked else None
    sm = blocky_mask(s, s, 4, 2, [3, 1, 0, 2, 7, 6, 4, 5]) if masked else None
    return mode, c, st, alpha, cm, sm


def oracle_stylize(sd, mode, c, s, alpha, cm, sm, with_roundtrip=False):
    """The CPU oracle on one image config -> (stylized, seconds, reference round-trip error stats or None)."""
    import torch
    from oracle import vst_oracle as O
    kw = MODES[mode]
    with torch.no_grad():
        t0 = time.perf_counter()
        zc, zs = O.revnet_forward(sd, c, **kw), O.revnet_forward(sd, s, **kw)
        if cm is not None:
            zcs = O.cwct_transfer_seg(zc, zs, cm, sm)
        elif alpha is not None:
            zcs = O.cwct_interpolation(zc, [zs], [1.0], alpha)
        else:
            zcs = O.cwct_transfer(zc, zs)
        y = O.revnet_inverse(sd, zcs, **kw)
        float(y[0, 0, 0, 0])
        dt = time.perf_counter() - t0
        rt = None
        if with_roundtrip:
            e = (O.revnet_inverse(sd, zc, **kw) - c).abs()
            rt = (float(e.max()), float(e.mean()))
    return y, dt, rt


# ----------------------------------------------------------------------------------------------
# CPU baseline: the oracle (a torch-fp32 port of the reference path) on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_frames_per_s(h, w, steps, warmup, threads):
    """Time `steps` hoisted-style frames of h x w on the CPU oracle; returns (frames/s at h x w, seconds/frame)."""
    import torch
    from oracle import vst_oracle as O
    torch.set_num_threads(threads)
    sd = O.init_state_dict(0, **MODES["photo"])           # the reference's default init; nothing of the product
    g = torch.Generator().manual_seed(99)
    style = torch.rand(1, 3, h, w, generator=g)
    frames = [torch.rand(1, 3, h, w, generator=g) for _ in range(2)]
    with torch.no_grad():
        zs = O.revnet_forward(sd, style)                      # hoisted, untimed (as in our arm)
        ts = []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            zc = O.revnet_forward(sd, frames[i % 2])
            y = O.revnet_inverse(sd, O.cwct_transfer(zc, zs))
            float(y[0, 0, 0, 0])
            if i >= warmup:
                ts.append(time.perf_counter() - t0)
    spf = sum(ts) / len(ts)
    return 1.0 / spf, spf


def bounded_cpu_sample(budget_s, steps, warmup, threads):
    """Pick a sub-frame (same aspect, multiple of 4) so (steps+warmup) CPU frames fit `budget_s`;
    cost is linear in pixels, so frames/s at 1080p = measured frames/s * (h*w)/(1080*1920)."""
    _, spf = cpu_frames_per_s(272, 480, 1, 1, threads)          # calibration on ~1/16 of the pixels
    full = spf * (H * W) / (272.0 * 480.0)
    frac = min(1.0, budget_s / max(1e-9, full * (steps + warmup)))
    scale = frac ** 0.5
    h = max(64, int(H * scale) // 4 * 4)
    w = max(64, int(W * scale) // 4 * 4)
    if frac >= 1.0:
        h, w = H, W
    fps, spf = cpu_frames_per_s(h, w, steps, warmup, threads)
    pix = (h * w) / float(H * W)
    return fps * pix, "%d frame(s) of %dx%d (%.1f%% of a 1080p frame's pixels, scaled linearly), %.2f s each" % (
        steps, w, h, 100 * pix, spf), h, w


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    value, sample, h, w = bounded_cpu_sample(150.0, args.steps, args.warmup, threads)
    assert "vstnet_b200" not in sys.modules, "the reference arm must not load the product"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_video": FRAMES_PER_VIDEO},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference is Python and cannot travel to the GPU box; this is oracle/vst_oracle.py, the same "
                "torch CPU ops (F.conv2d, linalg.cholesky) the reference issues, on all host cores, weights from "
                "oracle.init_state_dict (the reference's default init); no module of the product is imported",
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# the other BASELINE configs (single images): extras — "ms per image" = 2 encodes + cWCT + decode on the GPU,
# the CPU oracle timed on the same inputs (cfg1/2/3), and the parity of the two results
# ----------------------------------------------------------------------------------------------
def image_configs_ms(dev, precision, with_cpu):
    import torch
    from vstnet_b200 import RevResNet, cWCT

    out = {}
    nets = {}
    threads = os.cpu_count() or 1
    for name, (mode, size, alpha, masked) in IMAGE_CONFIGS.items():
        if mode not in nets:
            torch.manual_seed(0)
            nets[mode] = RevResNet(**MODES[mode], precision=precision).to(dev).eval()
        net, cw = nets[mode], cWCT()
        _, c_h, s_h, _, cm_h, sm_h = image_config_inputs(name)
        c, s = c_h.to(dev), s_h.to(dev)
        cm = torch.from_numpy(cm_h).to(dev) if masked else None
        sm = torch.from_numpy(sm_h).to(dev) if masked else None

        def run():
            zc, zs = net.encode_pair(c, s)               # what image_transfer.py's stylize() does
            if masked:
                zcs = cw.transfer(zc, zs, cm, sm)
            elif alpha is not None:
                zcs = cw.interpolation(zc, [zs], [1.0], alpha)
            else:
                zcs = cw.transfer(zc, zs)
            return net(zcs, forward=False)

        for _ in range(4):                    # allocator pools of both streams and the weight pack are warm
            run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        a.record()
        for _ in range(reps):
            y = run()
        b.record()
        torch.cuda.synchronize()
        out[name] = {"gpu_ms": round(a.elapsed_time(b) / reps, 3)}
        if with_cpu and size <= 1024:
            # CPU oracle on the same inputs and weights: the metric's "vs host-CPU torch" half, and the parity
            # of this very run (stylized pixels; round trip against the reference's own on the same input)
            torch.set_num_threads(threads)
            sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
            y_ref, dt, rt = oracle_stylize(sd, mode, c_h, s_h, alpha, cm_h, sm_h, with_roundtrip=True)
            e = (net.inverse(net(c)) - c).abs()
            out[name].update({
                "cpu_ms": round(dt * 1e3, 1), "cpu_cores": threads,
                "max_abs_vs_oracle": float((y.cpu() - y_ref).abs().max()),
                "roundtrip_max": float(e.max()), "roundtrip_mean": float(e.mean()),
                "roundtrip_ratio_vs_reference": {"max": float(e.max()) / max(rt[0], 1e-30),
                                                 "mean": float(e.mean()) / max(rt[1], 1e-30)}})
        del c, s, y
        torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------
# strong scaling: the WHOLE job — one 240-frame video over all ranks, host buffers, ordered delivery
# ----------------------------------------------------------------------------------------------
def strong_scaling_job(vs, style, host_pool, rank, world, dev, barrier):
    """Times set_style (style encode + statistics on rank 0 + the single broadcast) + every rank's share of
    FRAMES_PER_VIDEO frames through stylize_stream (H2D / D2H inside) + delivery of all frames IN ORDER to rank 0's
    consumer through the shared-memory frame ring (what video_transfer.py's writer reads).  Wall clock between two
    barriers, on rank 0."""
    import torch
    from vstnet_b200.video import SharedFrameRing, shard_frames
    mine = shard_frames(FRAMES_PER_VIDEO, rank, world)
    path = SharedFrameRing.default_path("bench_%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.getppid()))
    slots = 4 * world * vs.n_streams
    barrier()
    t0 = time.perf_counter()
    vs.set_style(style)
    if rank == 0:
        ring = SharedFrameRing(path, (H, W, 3), slots, create=True)
    barrier()
    if rank != 0:
        ring = SharedFrameRing(path, (H, W, 3), slots, create=False)
    consumer, seen = None, []
    if rank == 0:
        def consume():
            for i in range(FRAMES_PER_VIDEO):
                seen.append(int(ring.get(i)[0, 0, 0]))
                ring.release(i)
        consumer = threading.Thread(target=consume)
        consumer.start()
    for i, o in zip(mine, vs.stylize_stream(host_pool[i % len(host_pool)] for i in mine)):
        ring.put(i, o)
    if consumer is not None:
        consumer.join()
    barrier()
    dt = time.perf_counter() - t0
    ring.close(unlink=(rank == 0))
    return {"frames": FRAMES_PER_VIDEO, "seconds": dt, "value": FRAMES_PER_VIDEO / dt, "unit": UNIT,
            "delivered_in_order": len(seen) if rank == 0 else None,
            "includes": "style encode + stats (rank 0), one NCCL broadcast, H2D/D2H of every frame, ordered delivery "
                        "to rank 0 through the shared-memory frame ring; wall clock between barriers"}


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
F16_KERNELS = ("rev_block_tc", "conv3x3_tch", "conv3x3_tcH")      # kernel classes that issue kind::f16 UMMAs in f16x2 mode


def run_ours(args):
    import torch
    import torch.distributed as dist
    from vstnet_b200 import RevResNet, cWCT, _lib
    from vstnet_b200.video import VideoStylizer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    K = args.steps if args.steps is not None else max(1, FRAMES_PER_VIDEO // world // 4)
    Wm = max(3, args.warmup if args.warmup is not None else 4)

    torch.manual_seed(0)
    net = RevResNet(hidden_dim=16, sp_steps=2, precision=args.precision).to(dev).eval()
    vs = VideoStylizer(net, cWCT())
    gen = torch.Generator(device=dev)
    style = None
    if rank == 0:
        style = torch.rand(1, 3, H, W, device=dev, generator=gen.manual_seed(4321))
    vs.set_style(style)                               # rank 0 encodes + factorises, ONE broadcast
    pool = 4                                          # distinct resident frames cycled through
    frames = [torch.rand(1, 3, H, W, device=dev, generator=gen.manual_seed(1234 + rank * pool + i)) for i in range(pool)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_frames(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in vs.stylize_frames(frames[i % pool] for i in range(n)):      # frames dealt to vs.n_streams compute streams
            pass
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    # ---- device-resident throughput (`value`).  Setup (not warm-up): a few frames allocate every compute stream's
    # workspace and the output ring; then exactly W warm-up steps, then exactly K timed steps.
    n_setup = 2 * vs.n_streams + 2
    timed_frames(n_setup)
    timed_frames(Wm)
    barrier()
    sampler = ClockSampler(local).start() if rank == 0 else None
    n0 = _lib.launch_count()
    ms_local = timed_frames(K)
    launches = _lib.launch_count() - n0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    # ---- per-kernel durations: the SAME K steps once more with a CUDA-event pair around every launch
    # (the library's profiler), single stream, kept out of the `value` pass because ~330 event pairs per frame perturb it
    _lib.profile_enable(True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(K):
        y = vs.stylize(frames[i % pool])
    p1.record()
    torch.cuda.synchronize()
    _lib.profile_enable(False)
    prof = _lib.profile_collect()
    ms_prof = p0.elapsed_time(p1)
    barrier()
    ms = torch.tensor([ms_local], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    value = world * K / (ms / 1e3)

    # ---- sustained: a whole video's worth of frames back to back (power-cap behaviour on the record)
    sustained = None
    if not args.no_sustained:
        n_sus = max(K, FRAMES_PER_VIDEO // world)
        barrier()
        s2 = ClockSampler(local).start() if rank == 0 else None
        ms_s = torch.tensor([timed_frames(n_sus)], device=dev, dtype=torch.float64)
        barrier()
        ck = s2.stop() if rank == 0 else None
        if world > 1:
            dist.all_reduce(ms_s, op=dist.ReduceOp.MAX)
        sustained = {"frames": world * n_sus, "value": world * n_sus / (float(ms_s.item()) / 1e3), "unit": UNIT,
                     "ms_per_step": float(ms_s.item()) / n_sus, "clocks": ck}

    # ---- end to end through the host-buffer API (`e2e`): pinned uint8 HWC frames in (what a video decoder
    # delivers), pinned uint8 HWC frames out, both copies inside the timed region, pipelined by
    # VideoStylizer.stylize_stream (upload of frame i+1 and download of frame i-1 overlap frame i)
    host_frames = [(f[0].permute(1, 2, 0) * 255).round().clamp(0, 255).byte().contiguous().cpu().pin_memory() for f in frames]
    for _ in vs.stylize_stream(host_frames[i % pool] for i in range(max(Wm, 2 * vs.n_streams + 2))):
        pass
    barrier()
    Ke = K
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    t0.record()
    n_out = 0
    for out_host in vs.stylize_stream(host_frames[i % pool] for i in range(Ke)):
        n_out += 1
    t1.record()
    torch.cuda.synchronize()
    wall_e = (time.perf_counter() - w0) * 1e3
    assert n_out == Ke
    barrier()
    ms_e = torch.tensor([max(t0.elapsed_time(t1), wall_e)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
    e2e_value = world * Ke / (float(ms_e.item()) / 1e3)
    h2d = host_frames[0].numel()
    d2h = out_host.numel()

    # ---- strong scaling: the whole 240-frame job
    strong = None
    if not args.no_strong:
        strong = strong_scaling_job(vs, style, host_frames, rank, world, dev, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel table and the roofline of the dominant kernel, from the live per-launch events
    pk = peaks()
    issue = {"tf32": 1, "tf32x2": 2, "tf32x3": 3, "f16x2": 2}.get(args.precision, 1)
    traffic_tab = {}
    tf = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tf):
        try:
            traffic_tab = json.load(open(tf))
        except Exception:
            traffic_tab = {}

    def tensor_peak(name):
        """Dense peak of the MMA kind the kernel class issues: kind::f16 = the measured bf16 figure, kind::tf32 = half."""
        f16 = args.precision == "f16x2" and name.startswith(F16_KERNELS)
        return (pk["bf16_tflops_sustained"] if f16 else pk["bf16_tflops_sustained"] / 2.0), ("f16" if f16 else "tf32")

    def roof_of(name, d):
        per_ms = d["ms"] / max(1, d["launches"])
        gbs = d["bytes"] / d["ms"] / 1e6 if d["ms"] > 0 else 0.0
        tfl = d["flops"] / d["ms"] / 1e9 if d["ms"] > 0 else 0.0
        tpk, kind = tensor_peak(name)
        hbm_frac, ten_frac = gbs / pk["hbm_gbs"], tfl / tpk
        # the roofline that bounds the kernel is the one whose floor (work / peak) is the larger
        tensor_bound = d["flops"] > 0 and (d["flops"] * issue / (tpk * 1e12)) > (d["bytes"] / (pk["hbm_gbs"] * 1e9))
        r = {"kernel": name, "launches": d["launches"], "ms_per_launch": per_ms, "traffic": traffic_tab.get(name)}
        if tensor_bound:
            r.update({"bound": "tensor", "achieved": tfl, "peak": tpk, "unit": "TFLOP/s", "frac": ten_frac,
                      "frac_issued": ten_frac * issue,
                      "peak_source": "%s bf16 dense sustained%s (kind::%s); achieved = useful 1-term flops, %s issues %dx"
                                     % (pk["source"], "" if kind == "f16" else " / 2", kind, args.precision, issue)})
        else:
            r.update({"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": hbm_frac,
                      "peak_source": pk["source"] + " copy bandwidth"})
        return r, hbm_frac, ten_frac

    kernels = []
    for name, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        r, hf, tfr = roof_of(name, d)
        kernels.append({"kernel": name, "ms_total": round(d["ms"], 3), "launches": d["launches"],
                        "share": round(d["ms"] / ms_prof, 4), "ms_per_launch": round(r["ms_per_launch"], 4),
                        "tflops_useful": round(d["flops"] / d["ms"] / 1e9, 2) if d["ms"] > 0 else None,
                        "gbs": round(d["bytes"] / d["ms"] / 1e6, 1) if d["ms"] > 0 else None,
                        "bound": r["bound"], "frac": round(r["frac"], 4), "hbm_frac": round(hf, 4),
                        "tensor_frac_useful": round(tfr, 4)})
    top_name, top = max(prof.items(), key=lambda kv: kv[1]["ms"])
    roof, _, _ = roof_of(top_name, top)

    # ---- CPU baseline (rank 0, N == 1 only): the oracle on the host cores, bounded sample
    cpu = None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        v, sample, _, _ = bounded_cpu_sample(25.0, 1, 1, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}

    images = None
    if world == 1 and not args.no_images:
        del frames, host_frames
        torch.cuda.empty_cache()
        images = image_configs_ms(dev, args.precision, not args.no_cpu)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "f16x2": "f16x2+tf32x2 (fp32 accumulate)"}.get(args.precision, args.precision), "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "frames_per_video": FRAMES_PER_VIDEO, "frames_per_step_per_gpu": 1, "conv_precision": args.precision,
                   "l2": "per-frame working set (~1.5 GB of states) >> 126 MB L2; %d distinct frames cycled" % pool,
                   "setup": "%d untimed allocation frames (workspace per compute stream, output ring) before the %d warm-up steps" % (n_setup, Wm),
                   "parallelism": "frames sharded dp%d, no data-path collective; %d frames in flight per GPU (compute streams)" % (world, vs.n_streams)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "roofline": roof,
        "cpu_baseline": cpu,
        "useful_conv_tflops": 2 * CONV_FLOP_PER_PX * H * W * world * K / (ms / 1e3) / 1e12,
        "sustained": sustained,
        "strong": strong,
        "kernels_from": "a separate single-stream pass over the same K steps with a CUDA-event pair around every launch "
                        "(%.3f ms per step; `value` comes from the un-instrumented %d-stream pass)" % (ms_prof / K, vs.n_streams),
        "profile_pass_ms_per_step": ms_prof / K,
        "images": images,
        "kernels": kernels[:14],
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="f16x2", help="conv arithmetic: f16x2 (default) | tf32x2 | tf32x3 | tf32 | fp32")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg and the CPU side of the image configs")
    ap.add_argument("--no-images", action="store_true", help="skip the single-image configs (cfg1/2/3/5 extras)")
    ap.add_argument("--no-sustained", action="store_true", help="skip the 240-frame sustained extra")
    ap.add_argument("--no-strong", action="store_true", help="skip the whole-job strong-scaling extra")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps is None:
            args.steps = 3
        if args.warmup is None:
            args.warmup = 1
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
