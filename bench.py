#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native CAP-VSTNet stylization hot path.

Workload (BASELINE.json configs[3], the one `metric` is quoted on): photorealistic 1080p video,
random-init RevResNet (seed 0), synthetic frames, the style image encoded and its cWCT statistics
hoisted once (rank 0) and NCCL-broadcast; every rank then stylizes its own frames.  A *step* is
one 1920x1080 frame per rank: encode -> cWCT stats/factor/apply -> decode.  `value` is whole-job
frames/s with frames resident in HBM; `e2e` is the same through `VideoStylizer.stylize_stream`
with HOST buffers (pinned uint8 HWC frame H2D + uint8 result D2H inside the timed region, pipelined).
`images` (N = 1 only) adds ms per image for the other BASELINE configs (cfg1/2/3/5).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nproc-per-node N ... bench.py --gpus N ...
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
# CPU baseline: the oracle (a torch-fp32 port of the reference path) on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_frames_per_s(h, w, steps, warmup, threads):
    """Time `steps` hoisted-style frames of h x w on the CPU oracle; returns (frames/s at h x w, seconds/frame)."""
    import torch
    from oracle import vst_oracle as O
    from vstnet_b200 import RevResNet
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = RevResNet(hidden_dim=16, sp_steps=2)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
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
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg4 photorealistic video 1920x1080, style hoisted, random-init RevResNet",
                   "frames_per_video": FRAMES_PER_VIDEO},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference is Python and cannot travel to the GPU box; this is oracle/vst_oracle.py, the same "
                "torch CPU ops (F.conv2d, linalg.cholesky) the reference issues, on all host cores",
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# the other BASELINE configs (single images): reported as extras, "ms per image" = 2 encodes + cWCT + decode
# ----------------------------------------------------------------------------------------------
def image_configs_ms(dev, precision, with_cpu):
    import numpy as np
    import torch
    from vstnet_b200 import RevResNet, cWCT

    def blocky(h, w, gy, gx, perm):
        m = np.zeros((h, w), np.uint8)
        ys, xs = np.linspace(0, h, gy + 1).astype(int), np.linspace(0, w, gx + 1).astype(int)
        for i in range(gy):
            for j in range(gx):
                m[ys[i]:ys[i + 1], xs[j]:xs[j + 1]] = perm[i * gx + j]
        return m[None]

    out = {}
    nets = {}
    for name, mode, size, alpha, masked in (("cfg1_photo_512", "photo", 512, None, False),
                                            ("cfg2_art_1024_alpha0.5", "art", 1024, 0.5, False),
                                            ("cfg3_photo_1024_masked8", "photo", 1024, None, True),
                                            ("cfg5_art_4096", "art", 4096, None, False)):
        if mode not in nets:
            torch.manual_seed(0)
            kw = dict(hidden_dim=16, sp_steps=2) if mode == "photo" else dict(hidden_dim=64, sp_steps=1)
            nets[mode] = RevResNet(**kw, precision=precision).to(dev).eval()
        net, cw = nets[mode], cWCT()
        g = torch.Generator(device=dev).manual_seed(7)
        c = torch.rand(1, 3, size, size, device=dev, generator=g)
        s = torch.rand(1, 3, size, size, device=dev, generator=g)
        cm = torch.from_numpy(blocky(size, size, 2, 4, [0, 1, 2, 3, 4, 5, 6, 7])).to(dev) if masked else None
        sm = torch.from_numpy(blocky(size, size, 4, 2, [3, 1, 0, 2, 7, 6, 4, 5])).to(dev) if masked else None

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
            run()
        b.record()
        torch.cuda.synchronize()
        out[name] = {"gpu_ms": round(a.elapsed_time(b) / reps, 3)}
        del c, s
        torch.cuda.empty_cache()
    if with_cpu:                      # the reference's own CPU-runnable case (configs[0]) on the host cores
        from oracle import vst_oracle as O
        torch.manual_seed(0)
        net = RevResNet(hidden_dim=16, sp_steps=2)
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        gc = torch.Generator().manual_seed(7)
        c, s = torch.rand(1, 3, 512, 512, generator=gc), torch.rand(1, 3, 512, 512, generator=gc)
        with torch.no_grad():
            t0 = time.perf_counter()
            y = O.revnet_inverse(sd, O.cwct_transfer(O.revnet_forward(sd, c), O.revnet_forward(sd, s)))
            float(y[0, 0, 0, 0])
            out["cfg1_photo_512"]["cpu_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
            out["cfg1_photo_512"]["cpu_cores"] = os.cpu_count()
    return out


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
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
    Wm = args.warmup if args.warmup is not None else 3

    torch.manual_seed(0)
    net = RevResNet(hidden_dim=16, sp_steps=2, precision=args.precision).to(dev).eval()
    vs = VideoStylizer(net, cWCT())
    gen = torch.Generator(device=dev)
    style = None
    if rank == 0:
        style = torch.rand(1, 3, H, W, device=dev, generator=gen.manual_seed(4321))
    vs.set_style(style)                               # rank 0 encodes + factorises, one broadcast
    pool = 4                                          # distinct resident frames cycled through
    frames = [torch.rand(1, 3, H, W, device=dev, generator=gen.manual_seed(1234 + rank * pool + i)) for i in range(pool)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`)
    Wm = max(Wm, 4 * vs.n_streams)                    # every compute stream's workspace / allocator pool is warm
    for _ in vs.stylize_frames(frames[i % pool] for i in range(Wm)):
        pass
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for y in vs.stylize_frames(frames[i % pool] for i in range(K)):      # frames dealt to vs.n_streams compute streams
        pass
    e1.record()
    torch.cuda.synchronize()
    launches = _lib.launch_count() - n0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    # ---- per-kernel durations: the SAME K steps once more with a CUDA-event pair around every launch
    # (the library's profiler), kept out of the `value` pass because ~330 event pairs per frame perturb it
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
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    value = world * K / (ms / 1e3)

    # ---- end to end through the host-buffer API (`e2e`): pinned uint8 HWC frames in (what a video decoder
    # delivers), pinned uint8 HWC frames out, both copies inside the timed region, pipelined by
    # VideoStylizer.stylize_stream (upload of frame i+1 and download of frame i-1 overlap frame i)
    host_frames = [(f[0].permute(1, 2, 0) * 255).round().clamp(0, 255).byte().contiguous().cpu().pin_memory() for f in frames]
    for _ in vs.stylize_stream(host_frames[i % pool] for i in range(max(Wm, 2 * vs.n_streams))):
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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel, from the live per-launch events of the timed region
    pk = peaks()
    kernels = []
    for name, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        per = d["ms"] / max(1, d["launches"])
        kernels.append({"kernel": name, "ms_total": round(d["ms"], 3), "launches": d["launches"],
                        "share": round(d["ms"] / ms_prof, 4),
                        "tflops": round(d["flops"] / d["ms"] / 1e9, 2) if d["ms"] > 0 else None,
                        "gbs": round(d["bytes"] / d["ms"] / 1e6, 1) if d["ms"] > 0 else None,
                        "ms_per_launch": round(per, 4)})
    top_name, top = max(prof.items(), key=lambda kv: kv[1]["ms"])
    ai = top["flops"] / max(1.0, top["bytes"])
    traffic = None
    tf = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tf):
        try:
            traffic = json.load(open(tf)).get(top_name)
        except Exception:
            traffic = None
    # kind::tf32 runs at half the bf16 rate, so the tensor roof of these kernels is the measured bf16 figure / 2;
    # `achieved` counts USEFUL conv flops (one term), not the 2x / 3x issued by the split-precision modes.
    tf32_peak = pk["bf16_tflops_sustained"] / 2.0
    if ai > tf32_peak * 1e3 / pk["hbm_gbs"]:                            # above the TF32 ridge -> tensor bound
        ach = top["flops"] / top["ms"] / 1e9
        roof = {"kernel": top_name, "bound": "tensor", "achieved": ach, "peak": tf32_peak,
                "unit": "TFLOP/s", "frac": ach / tf32_peak, "traffic": traffic,
                "peak_source": pk["source"] + " bf16 dense sustained / 2 (tf32 rate); achieved = useful 1-term flops, "
                               "%s issues %dx" % (args.precision, {"tf32": 1, "tf32x2": 2, "tf32x3": 3, "f16x2": 2}.get(args.precision, 1))}
    else:
        ach = top["bytes"] / top["ms"] / 1e6
        roof = {"kernel": top_name, "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": ach / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk["source"] + " copy bandwidth"}
    roof["launches"] = top["launches"]
    roof["ms_per_launch"] = top["ms"] / max(1, top["launches"])

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
        "config": {"workload": "cfg4 photorealistic video 1920x1080, style hoisted + broadcast, random-init RevResNet",
                   "frames_per_video": FRAMES_PER_VIDEO, "frames_per_step_per_gpu": 1, "conv_precision": args.precision,
                   "l2": "per-frame working set (~1.5 GB of states) >> 126 MB L2; %d distinct frames cycled" % pool,
                   "parallelism": "frames sharded dp%d, no data-path collective; %d frames in flight per GPU (compute streams)" % (world, vs.n_streams)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "roofline": roof,
        "cpu_baseline": cpu,
        "useful_conv_tflops": 2 * CONV_FLOP_PER_PX * H * W * world * K / (ms / 1e3) / 1e12,
        "profile_pass_ms_per_step": ms_prof / K,
        "images": images,
        "kernels": kernels[:12],
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-images", action="store_true", help="skip the single-image configs (cfg1/2/3/5 extras)")
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
