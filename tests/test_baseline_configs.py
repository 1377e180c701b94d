"""GPU parity at the BASELINE configs' OWN sizes: the CUDA path (default arithmetic, as benchmarked) against the CPU
oracle on the same seeded inputs — cfg1 512x512 photo, cfg2 1024x1024 artistic alpha_c = 0.5, cfg3 1024x1024 photo
with 8-label masks (and the invalid-label variant), one cfg4 1080p frame through VideoStylizer, cfg5 at 2048x2048
end to end plus the 4096x4096 statistics against an fp64 evaluation.  These pin the persistent-tile schedules,
split-K and strip x segment geometry at the sizes bench.py times (SURVEY.md 8(d) inputs = bench.image_config_inputs).

Tolerances: stylized pixels max-abs <= 1e-3 (BASELINE.json); round trip "at or below the reference's own": max within
2 fp32 ulp of the reference's max, mean within the factor stated in tests/test_gpu_parity.py (both sides are fp32
rounding noise; the measured ratio is printed by bench.py per config).
"""
import numpy as np
import pytest
import torch

import bench
from oracle import vst_oracle as O
from tests.helpers import MODES, build_net, cpu_state_dict
from tests.test_gpu_parity import PIXEL_TOL, assert_roundtrip_at_reference_level, maxdiff

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _gpu_stylize(net, dev, c, s, alpha, cm, sm):
    from vstnet_b200 import cWCT
    cw = cWCT()
    zc, zs = net.encode_pair(c.to(dev), s.to(dev))
    z_keep = zc.clone()
    if cm is not None:
        zcs = cw.transfer(zc, zs, cm, sm)
    elif alpha is not None:
        zcs = cw.interpolation(zc, [zs], [1.0], alpha)
    else:
        zcs = cw.transfer(zc, zs)
    return net(zcs, forward=False), z_keep, zs, cw


@pytest.mark.parametrize("name", ["cfg1_photo_512", "cfg2_art_1024_alpha0.5", "cfg3_photo_1024_masked8"])
def test_image_config_vs_oracle_at_full_size(dev, name):
    mode, c, s, alpha, cm, sm = bench.image_config_inputs(name)
    net = build_net(mode, 0)
    sd = cpu_state_dict(net)
    net = net.to(dev)
    y, zc, _, _ = _gpu_stylize(net, dev, c, s, alpha, cm, sm)
    y_ref, _, rt = bench.oracle_stylize(sd, mode, c, s, alpha, cm, sm, with_roundtrip=False)
    err = maxdiff(y, y_ref)
    print("%s: max|cuda - oracle| = %.3e" % (name, err))
    assert err <= PIXEL_TOL
    with torch.no_grad():
        zr = O.revnet_forward(sd, c, **MODES[mode])
        xr_ref = O.revnet_inverse(sd, zr, **MODES[mode])
    assert maxdiff(zc, zr) <= 1e-3
    assert_roundtrip_at_reference_level(net.inverse(zc).cpu(), xr_ref, c, "f16x2")


def test_cfg3_invalid_label_variant_vs_oracle(dev):
    """cfg3 with a label that fails the validity rule (absent from the style): its pixels keep the content features,
    every other label matches the oracle, at 1024x1024."""
    mode, c, s, alpha, cm, _ = bench.image_config_inputs("cfg3_photo_1024_masked8")
    sm = bench.blocky_mask(1024, 1024, 4, 2, [3, 1, 0, 2, 7, 6, 4, 4])          # label 5 absent from the style
    net = build_net(mode, 0)
    sd = cpu_state_dict(net)
    net = net.to(dev)
    y, zc, zs, cw = _gpu_stylize(net, dev, c, s, alpha, cm, sm)
    y_ref, _, _ = bench.oracle_stylize(sd, mode, c, s, alpha, cm, sm)
    assert maxdiff(y, y_ref) <= PIXEL_TOL
    # identity on the invalid label at the latent level (the in-place result is zc itself)
    z2 = zc.clone()
    out = cw.transfer(z2, zs, cm, sm)
    keep = torch.from_numpy(cm[0] == 5).to(dev)
    assert torch.equal(out[0][:, keep], zc[0][:, keep])


def test_cfg4_one_1080p_frame_through_video_stylizer_vs_oracle(dev):
    """One 1920x1080 frame of the benchmarked video path (hoisted style statistics, VideoStylizer.stylize) against
    the oracle's encode(content), encode(style), cWCT, decode."""
    from vstnet_b200.video import VideoStylizer
    net = build_net("photo", 0)
    sd = cpu_state_dict(net)
    net = net.to(dev)
    g = torch.Generator().manual_seed(4321)
    style = torch.rand(1, 3, 1080, 1920, generator=g)
    frame = torch.rand(1, 3, 1080, 1920, generator=g)
    vs = VideoStylizer(net)
    vs.set_style(style.to(dev))
    y = vs.stylize(frame.to(dev))
    with torch.no_grad():
        zc, zs = O.revnet_forward(sd, frame), O.revnet_forward(sd, style)
        y_ref = O.revnet_inverse(sd, O.cwct_transfer(zc, zs))
        xr_ref = O.revnet_inverse(sd, zc)
    err = maxdiff(y, y_ref)
    print("cfg4 frame: max|cuda - oracle| = %.3e" % err)
    assert err <= PIXEL_TOL
    z = net(frame.to(dev))
    assert maxdiff(z, zc) <= 1e-3
    assert_roundtrip_at_reference_level(net.inverse(z).cpu(), xr_ref, frame, "f16x2")
    # and through the pipelined host path: uint8 in, uint8 out, +-1 LSB of the oracle on the same quantised input
    u8 = (frame[0].permute(1, 2, 0) * 255).round().clamp(0, 255).byte()
    outs = [o.clone() for o in vs.stylize_stream([u8.pin_memory()])]
    fq = u8.permute(2, 0, 1)[None].float() / 255
    with torch.no_grad():
        yq = O.revnet_inverse(sd, O.cwct_transfer(O.revnet_forward(sd, fq), zs))
    refq = yq[0].mul(255).clamp(0, 255).permute(1, 2, 0)
    # truncation: a float difference of 1e-3 * 255 can move a value across an integer boundary
    assert float((outs[0].float() - refq.floor()).abs().max()) <= 1.0


def test_cfg5_2048_end_to_end_vs_oracle(dev):
    """cfg5's path (artistic, large-HW Gram, HBM-bound apply) at 2048x2048 against the oracle (4096x4096 costs the
    CPU minutes and >10 GB; its statistics are checked separately below)."""
    mode, c, s, alpha, cm, sm = bench.image_config_inputs("cfg5_art_4096", size=2048)
    net = build_net(mode, 0)
    sd = cpu_state_dict(net)
    net = net.to(dev)
    y, zc, _, _ = _gpu_stylize(net, dev, c, s, alpha, cm, sm)
    y_ref, _, _ = bench.oracle_stylize(sd, mode, c, s, alpha, cm, sm)
    err = maxdiff(y, y_ref)
    print("cfg5 @2048: max|cuda - oracle| = %.3e" % err)
    assert err <= PIXEL_TOL


def _stats_vs_fp64(dev, feat, C, n):
    """vst_cwct_stats of feat [C, n] (device) vs an fp64 evaluation; returns (mean err, cov err relative to max|cov|)."""
    from vstnet_b200 import cWCT
    cw = cWCT()
    st = torch.cuda.current_stream(dev).cuda_stream
    buf = cw._stats(feat, C, n, None, 1, st)
    torch.cuda.synchronize()
    d = buf[: 8 * (1 + C + C * C)].view(torch.float64)
    piv = buf[8 * (1 + C + C * C): 8 * (1 + C + C * C) + 4 * C].view(torch.float32).double()
    cnt, ssum, gram = float(d[0]), d[1:1 + C], d[1 + C:].reshape(C, C)
    assert cnt == n
    mean = piv + ssum / n
    g = 0.5 * (gram + gram.T)
    cov = (g - torch.outer(ssum, ssum) / n) / (n - 1)
    # fp64 reference in slabs (memory)
    m64 = torch.zeros(C, dtype=torch.float64, device=dev)
    for p0 in range(0, n, 1 << 20):
        m64 += feat[:, p0:p0 + (1 << 20)].double().sum(1)
    m64 /= n
    c64 = torch.zeros(C, C, dtype=torch.float64, device=dev)
    for p0 in range(0, n, 1 << 20):
        xc = feat[:, p0:p0 + (1 << 20)].double() - m64[:, None]
        c64 += xc @ xc.T
    c64 /= (n - 1)
    return float((mean - m64).abs().max()), float((cov - c64).abs().max() / c64.abs().max())


@pytest.mark.parametrize("C,n", [(32, 16384), (32, 16384 + 4 * 37), (32, 1920 * 1080), (64, 40000), (64, 4 * 12345 + 16384),
                                 (128, 16384), (128, 512 * 512 + 4 * 9), (16, 65536 + 12)])
def test_gram_tc_kernel_vs_fp64(dev, C, n):
    """The tensor-core Gram (taken for n >= 16384, n % 4 == 0) directly against fp64: every segment count
    (C = 32 / 64 / 128, C = 16 padded), n not a multiple of the stage run, a full 1080p latent."""
    from vstnet_b200 import _lib
    g = torch.Generator(device=dev).manual_seed(C * 7 + n % 1000)
    feat = (torch.randn(C, n, device=dev, generator=g) * torch.linspace(0.2, 1.5, C, device=dev)[:, None]
            + torch.linspace(-0.7, 0.9, C, device=dev)[:, None]).contiguous()
    # correlate the channels so the covariance has structure
    feat[1::2] += 0.5 * feat[0::2]
    n0 = _lib.launch_count()
    em, ec = _stats_vs_fp64(dev, feat, C, n)
    assert _lib.launch_count() - n0 >= 2
    print("gram_tc C=%d n=%d: mean err %.2e, cov rel err %.2e" % (C, n, em, ec))
    # the 3-term tf32 split carries ~22 bits per product and the accumulator is fp32 per <= 1024-pixel run before the
    # fp64 fold: a few 1e-6 of max|cov| (measured 3e-7 .. 2.4e-6), the level of an fp32 matmul of the same length
    # (C = 128 folds into fp64 every 1024 pixels instead of 256: its 128 x 128 accumulator does not fit the drain warps'
    # registers, so every fold is a round of global atomics: measured 9e-6)
    assert em <= 2e-6 and ec <= (2e-5 if C > 64 else 5e-6)


def test_cfg5_4096_stats_vs_fp64(dev):
    """The 2.1 GB artistic latent of a 4096x4096 image: large-HW Gram (split-K over 4.2 M pixels per channel) vs fp64."""
    net = build_net("art", 0).to(dev)
    gen = torch.Generator(device=dev)
    x = torch.rand(1, 3, 4096, 4096, device=dev, generator=gen.manual_seed(21))
    z = net(x)
    del x
    em, ec = _stats_vs_fp64(dev, z[0].reshape(128, -1), 128, 2048 * 2048)
    print("cfg5 4096 stats: mean err %.2e, cov rel err %.2e" % (em, ec))
    assert em <= 2e-6 and ec <= 2e-5


# ----------------------------------------------------------------------------------------------------------------
# the reference's helper methods on the device (cWCT.py:111-164)
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C,n", [(32, 5000), (128, 3000)])
def test_whitening_coloring_cholesky_dec_vs_reference_formulas(dev, C, n):
    from vstnet_b200 import cWCT
    g = torch.Generator().manual_seed(C + n)
    A = torch.randn(C, C, generator=g) / C ** 0.5 + torch.eye(C)
    x = (A @ torch.randn(C, n, generator=g)) * 0.5 + 0.3
    s = (A.T @ torch.randn(C, n + 77, generator=g)) * 0.4 - 0.1
    cw = cWCT()
    # oracle: the reference's 2-D formulas in fp64
    xd, sd_ = x.double(), s.double()
    mu, xc, cov = O._mean_cov(xd)
    Lc = torch.linalg.cholesky(cov)
    white_ref = torch.inverse(Lc) @ xc
    w = cw.whitening(x.to(dev))
    assert w.shape == x.shape
    assert maxdiff(w, white_ref) <= 2e-4 * float(white_ref.abs().max())
    wcov = torch.cov(w.double().cpu())
    assert float((wcov - torch.eye(C, dtype=torch.float64)).abs().max()) <= 2e-3        # whitened: identity covariance
    mus, _, covs = O._mean_cov(sd_)
    col_ref = torch.linalg.cholesky(covs) @ white_ref + mus[:, None]
    col = cw.coloring(w, s.to(dev))
    assert maxdiff(col, col_ref) <= 5e-5 * max(1.0, float(col_ref.abs().max()))
    # composition == transfer
    t = cw.transfer(x.reshape(1, C, 1, n).to(dev), s.reshape(1, C, 1, n + 77).to(dev))
    assert maxdiff(t.reshape(C, n), col.cpu()) <= 2e-5
    # batched input (upstream semantics)
    wb = cw.whitening(torch.stack([x, x * 0.5 + 1]).to(dev))
    assert maxdiff(wb[0], w.cpu()) <= 1e-6 and maxdiff(wb[1], w.cpu()) <= 2e-4 * float(white_ref.abs().max())
    # cholesky_dec: plain, inverted, fp64, and a singular matrix that needs the jitter
    L = cw.cholesky_dec(cov.float().to(dev))
    assert maxdiff(L, torch.linalg.cholesky(cov.float())) <= 1e-5
    assert float(L.cpu().triu(1).abs().max()) == 0.0
    Li = cw.cholesky_dec(cov.float().to(dev), invert=True)
    assert maxdiff(Li, torch.inverse(torch.linalg.cholesky(cov))) <= 2e-4 * float(torch.inverse(Lc).abs().max())
    L64 = cw.cholesky_dec(cov.to(dev))
    assert L64.dtype == torch.float64 and maxdiff(L64, Lc) <= 1e-12
    assert int(cw.last_status.cpu()[0]) == 0
    v = torch.randn(C, 3, generator=g).double()
    sing = (v @ v.T).float()                                       # rank 3
    Lj = cw.cholesky_dec(sing.to(dev))
    k = int(cw.last_status.cpu()[0])
    assert k >= 1, "a rank-3 matrix cannot factor without jitter"
    jit = cw.eps * k * (k + 1) / 2
    rec = (Lj.double() @ Lj.double().T).cpu()
    assert float((rec - (sing.double() + jit * torch.eye(C, dtype=torch.float64))).abs().max()) <= 1e-5 * float(sing.abs().max())
    with pytest.raises(RuntimeError):
        cw.whitening(x)                                            # CPU tensor: no fallback


def test_masked_transfer_issues_no_host_sync(dev):
    """With device masks the masked path must enqueue everything without waiting for the GPU: a long kernel queued
    first is still running when transfer() returns."""
    from vstnet_b200 import cWCT
    g = torch.Generator().manual_seed(5)
    zc = torch.randn(1, 32, 64, 64, generator=g).to(dev)
    zs = torch.randn(1, 32, 64, 64, generator=g).to(dev)
    cm = torch.randint(0, 3, (1, 64, 64), generator=g, dtype=torch.uint8).to(dev)
    sm = torch.randint(0, 3, (1, 64, 64), generator=g, dtype=torch.uint8).to(dev)
    cw = cWCT()
    cw.transfer(zc.clone(), zs, cm, sm)                 # warm (allocations)
    torch.cuda.synchronize()
    big = torch.empty(1 << 28, device=dev)
    ev = torch.cuda.Event()
    for _ in range(30):
        big.normal_()                                   # ~ tens of ms of queued work
    out = cw.transfer(zc.clone(), zs, cm, sm)
    ev.record()
    assert not ev.query(), "transfer() waited for the device"
    torch.cuda.synchronize()
    ref = O.cwct_transfer_seg(zc.cpu(), zs.cpu(), cm.cpu().numpy(), sm.cpu().numpy())
    assert maxdiff(out, ref) <= 1e-4


# ----------------------------------------------------------------------------------------------------------------
# f16x2 dynamic range
# ----------------------------------------------------------------------------------------------------------------
def test_f16x2_range_guard_and_small_activations(dev):
    """(a) weights / inputs scaled to a trained-checkpoint-like dynamic range (the reference's checkpoint gives
    |z| ~ 1.2, image_transfer.py:183-205; here activations reach a few hundred) still match the oracle relative to
    their magnitude; (b) tiny activations (|x| < 4e-3: the fp16 lo term goes subnormal) keep fp32-level ABSOLUTE
    accuracy; (c) an activation beyond the fp16 operand range raises the status bit: 'strict' re-runs the call in
    tf32x2 and returns the right answer, 'deferred' warns on a later call and switches the module."""
    import warnings
    net = build_net("photo", 0, 7)
    sd = cpu_state_dict(net)
    net = net.to(dev)
    x = torch.rand(1, 3, 72, 88, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        # (a) inputs x 100: |state| up to ~ 300
        z_ref = O.revnet_forward(sd, x * 100)
        z = net(x.to(dev) * 100)
        scale = float(z_ref.abs().max())
        assert 50 < scale < 1000
        assert maxdiff(z, z_ref) <= 5e-4 * scale        # relative level of the unit-scale latent bar (Z_TOL 1e-3 at |z| ~ 2)
        assert net.check_status() is False and net.precision == "f16x2"
        # (b) inputs x 1e-3 on a bias-free net (every activation scales with the input: |x| down to ~1e-6)
        net0 = build_net("photo", 0)
        sd0 = cpu_state_dict(net0)
        net0 = net0.to(dev)
        z_ref = O.revnet_forward(sd0, x * 1e-3)
        z = net0(x.to(dev) * 1e-3)
        small = float(z_ref.abs().max())
        assert small < 1e-2
        print("tiny activations: |z| <= %.2e, err %.2e, round trip %.2e" % (small, maxdiff(z, z_ref), maxdiff(net0.inverse(z), x * 1e-3)))
        assert maxdiff(z, z_ref) <= 5e-4 * small
        assert maxdiff(net0.inverse(z), x * 1e-3) <= 1e-8          # fp32 noise at this scale is 1e-3 * 4e-7
        # (c) beyond the range
        net.range_check = "strict"
        z_ref = O.revnet_forward(sd, x * 5000)
        with warnings.catch_warnings(record=True) as wlist:
            warnings.simplefilter("always")
            z = net(x.to(dev) * 5000)
        assert any("fp16 operand range" in str(w.message) for w in wlist)
        assert net.precision == "tf32x2"
        assert bool(torch.isfinite(z).all()) and maxdiff(z, z_ref) <= 1e-3 * float(z_ref.abs().max())
        net.precision = "f16x2"
        net.range_check = "deferred"
        with warnings.catch_warnings(record=True) as wlist:
            warnings.simplefilter("always")
            net(x.to(dev) * 5000)
            torch.cuda.synchronize()
            assert net.check_status() is True
        assert any("fp16 operand range" in str(w.message) for w in wlist) and net.precision == "tf32x2"


# ----------------------------------------------------------------------------------------------------------------
# entry points
# ----------------------------------------------------------------------------------------------------------------
def test_video_transfer_entry_point_frames_vs_oracle(dev, tmp_path):
    """video_transfer.main on synthetic frames: every returned uint8 frame within 1 LSB of the oracle's stylization of
    the same uint8 input (truncating quantisation, video_transfer.py:211-214), and the video file is written."""
    import video_transfer
    from image_transfer import build_network
    torch.manual_seed(123)                                         # main() builds the network first: default init under this seed
    frames = video_transfer.main(["--synthetic", "64x96x5", "--out_dir", str(tmp_path)])
    assert sorted(frames) == [0, 1, 2, 3, 4] and (tmp_path / "synthetic_64x96x5.mp4").exists()
    torch.manual_seed(123)
    sd = cpu_state_dict(build_network("photorealistic"))
    rng = np.random.default_rng(0)
    src = [rng.integers(0, 256, (64, 96, 3), dtype=np.uint8) for _ in range(5)]
    style = torch.rand(1, 3, 64, 96, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        zs = O.revnet_forward(sd, style)
        for i in range(5):
            f = torch.from_numpy(src[i]).permute(2, 0, 1)[None].float() / 255
            y = O.revnet_inverse(sd, O.cwct_transfer(O.revnet_forward(sd, f), zs))
            ref = y[0].mul(255).clamp(0, 255).permute(1, 2, 0)
            assert frames[i].shape == (64, 96, 3)
            assert float(np.abs(frames[i].astype(np.float32) - ref.floor().numpy()).max()) <= 1.0
            assert float(np.abs(frames[i].astype(np.float32) - ref.numpy()).max()) <= 1.0 + 255 * PIXEL_TOL


def test_image_transfer_entry_point_vs_oracle(dev, tmp_path):
    import image_transfer
    from image_transfer import build_network
    torch.manual_seed(321)
    y = image_transfer.main(["--synthetic", "64x96", "--out_dir", str(tmp_path)])
    torch.manual_seed(321)
    sd = cpu_state_dict(build_network("photorealistic"))
    g = torch.Generator().manual_seed(0)
    c, s = torch.rand(1, 3, 64, 96, generator=g), torch.rand(1, 3, 64, 96, generator=g)
    with torch.no_grad():
        ref = O.revnet_inverse(sd, O.cwct_transfer(O.revnet_forward(sd, c), O.revnet_forward(sd, s)))
    assert maxdiff(y, ref) <= PIXEL_TOL


def test_encode_pair_on_a_fresh_net_with_a_backlog(dev):
    """the first encode_pair of a fresh module packs the weights while a long kernel is queued on the
    caller's stream; both encodes must still see the finished pack."""
    net = build_net("photo", 0, 7).to(dev)
    g = torch.Generator().manual_seed(44)
    a, b = torch.rand(1, 3, 72, 88, generator=g).to(dev), torch.rand(1, 3, 40, 136, generator=g).to(dev)
    big = torch.empty(1 << 28, device=dev)
    for _ in range(20):
        big.normal_()
    za, zb = net.encode_pair(a, b)
    torch.cuda.synchronize()
    ref = build_net("photo", 0, 7).to(dev)
    assert torch.equal(za, ref(a)) and torch.equal(zb, ref(b))


# ----------------------------------------------------------------------------------------------------------------
# fused frame path (vst_revnet_stylize): statistics from, and transform applied to, the network's own state
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode,h,w,alpha", [("photo", 136, 264, None), ("photo", 128, 132, 0.3), ("art", 132, 260, None),
                                             ("art", 256, 128, 0.5)])
def test_fused_frame_path_equals_unfused(dev, mode, h, w, alpha):
    """Same result as encode -> cWCT (materialised latent) -> decode, fp32 and uint8 I/O; ragged widths exercise the
    partial pixel blocks of the state Gram and the border handling of the in-place apply."""
    from vstnet_b200 import _lib
    from vstnet_b200.video import VideoStylizer
    net = build_net(mode, 0, 7).to(dev)
    assert net.stylize_supported(h, w)
    g = torch.Generator().manual_seed(h * 7 + w)
    style = torch.rand(1, 3, 96, 160, generator=g).to(dev)
    frame = torch.rand(1, 3, h, w, generator=g).to(dev)
    vs = VideoStylizer(net, alpha_c=alpha)
    vs.set_style(style)
    n0 = _lib.launch_count()
    y = vs.stylize(frame)
    fused_launches = _lib.launch_count() - n0
    vs.fused = False
    n0 = _lib.launch_count()
    y_ref = vs.stylize(frame)
    assert fused_launches < _lib.launch_count() - n0, "the fused path must save launches"
    err = maxdiff(y, y_ref.cpu())
    print("fused vs unfused %s %dx%d: %.3e" % (mode, h, w, err))
    assert err <= 2e-5
    # uint8 in / out: bit-identical to the fp32 fused path between the two conversion kernels
    vs.fused = True
    u8 = (frame[0].permute(1, 2, 0) * 255).round().clamp(0, 255).byte().contiguous()
    o = net.stylize_frame(u8, vs.style_pre["stats"][0], 0.0 if alpha is None else alpha, bgr=True)
    # byte / 255 on the CPU: a true division, as ToTensor does (torch's CUDA division by a scalar multiplies by 1/255)
    f = (u8.cpu().flip(-1).permute(2, 0, 1)[None].float() / 255).contiguous().to(dev)
    yf = net.stylize_frame(f, vs.style_pre["stats"][0], 0.0 if alpha is None else alpha)
    want = yf[0].mul(255).clamp(0, 255).byte().permute(1, 2, 0).flip(-1)
    assert torch.equal(o, want)
    assert net.check_status() is False


def test_fused_frame_path_vs_oracle_midsize(dev):
    net = build_net("photo", 0, 7)
    sd = cpu_state_dict(net)
    net = net.to(dev)
    g = torch.Generator().manual_seed(9)
    style, frame = torch.rand(1, 3, 160, 224, generator=g), torch.rand(1, 3, 200, 328, generator=g)
    from vstnet_b200.video import VideoStylizer
    vs = VideoStylizer(net)
    vs.set_style(style.to(dev))
    y = vs.stylize(frame.to(dev))
    with torch.no_grad():
        ref = O.revnet_inverse(sd, O.cwct_transfer(O.revnet_forward(sd, frame), O.revnet_forward(sd, style)))
    assert maxdiff(y, ref) <= 1e-4


# ----------------------------------------------------------------------------------------------------------------
# mask preparation on the device
# ----------------------------------------------------------------------------------------------------------------
def test_seg_remapping_on_device_vs_reference_vectors(dev):
    from models.segmentation.SegReMapping import SegReMapping
    from tests.helpers import load_golden
    g = load_golden("seg_remap.npz")
    for k in range(3):
        rm = SegReMapping(g["table"], min_ratio=float(g["ratio%d" % k]))
        cs = rm.self_remapping(g["c%d" % k])
        ss = rm.self_remapping(torch.from_numpy(g["s%d" % k]).to(dev))
        assert cs.is_cuda and cs.dtype == torch.uint8
        assert np.array_equal(cs.cpu().numpy(), g["c_self%d" % k]) and np.array_equal(ss.cpu().numpy(), g["s_self%d" % k])
        cc = rm.cross_remapping(cs, ss)
        assert np.array_equal(cc.cpu().numpy(), g["c_cross%d" % k])
    # a large random map against the oracle (odd size: unaligned tail of the vectorised histogram)
    rng = np.random.default_rng(5)
    seg = rng.choice(np.arange(150, dtype=np.uint8), size=(1021, 1531), p=np.r_[np.full(10, 0.09), np.full(140, 0.1 / 140)])
    sty = rng.choice(np.arange(0, 150, 3, dtype=np.uint8), size=(777, 1023))
    rm = SegReMapping(g["table"], min_ratio=0.01)
    a = rm.self_remapping(seg)
    assert np.array_equal(a.cpu().numpy(), O.seg_self_remapping(seg, g["table"], 0.01))
    b = rm.cross_remapping(a, sty)
    assert np.array_equal(b.cpu().numpy(), O.seg_cross_remapping(a.cpu().numpy(), sty, g["table"]))


def test_labels_from_colors_on_device(dev, tmp_path):
    from PIL import Image
    from tests.helpers import load_golden
    from vstnet_b200.hostio import labels_from_colors as host_rule, load_segment
    from vstnet_b200.segmentation import labels_from_colors
    h = load_golden("hostio.npz")
    out = labels_from_colors(h["seg"], dev)
    assert np.array_equal(out.cpu().numpy(), host_rule(h["seg"]))
    rng = np.random.default_rng(2)
    big = rng.integers(0, 256, (517, 389, 3), dtype=np.uint8)
    assert np.array_equal(labels_from_colors(big, dev).cpu().numpy(), O.seg_labels_from_colors(big))
    path = str(tmp_path / "seg.png")
    Image.fromarray(h["seg"]).save(path)
    lab = load_segment(path, device=dev)
    assert lab.is_cuda and np.array_equal(lab.cpu().numpy(), host_rule(h["seg"]))
    # device labels go straight into the masked transfer
    from vstnet_b200 import cWCT
    g = torch.Generator().manual_seed(3)
    zc, zs = torch.randn(1, 32, 24, 20, generator=g), torch.randn(1, 32, 24, 20, generator=g)
    ref = O.cwct_transfer_seg(zc, zs, lab.cpu().numpy()[None], lab.cpu().numpy()[None])
    got = cWCT().transfer(zc.to(dev), zs.to(dev), lab[None], lab[None])
    assert maxdiff(got, ref) <= 1e-4


# ----------------------------------------------------------------------------------------------------------------
# per-label statistics on the tensor cores (gram_tc_masked_kernel: n >= 16384, C = 32)
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("h,w,kind", [(128, 192, "blocky"), (136, 260, "blocky"), (96, 256, "speckled"), (200, 132, "stripes")])
def test_masked_transfer_tensor_core_stats_vs_oracle(dev, h, w, kind):
    """Label maps of the size class that takes the tensor-core masked Gram: region-like maps (one label per stage, two
    where regions meet), vertical stripes narrower than a stage (every stage mixed), per-pixel random labels (every
    stage holds every label), widths that are not a multiple of the 128-pixel stage, an invalid and an absent label."""
    from vstnet_b200 import _lib, cWCT
    g = torch.Generator().manual_seed(h + w)
    zc = torch.randn(1, 32, h, w, generator=g) * 0.5 + 0.1
    zs = torch.randn(1, 32, h, w, generator=g) * 0.3 - 0.2
    if kind == "blocky":
        cm = bench.blocky_mask(h, w, 2, 4, [0, 1, 2, 3, 4, 5, 6, 7])
        sm = bench.blocky_mask(h, w, 4, 2, [3, 1, 0, 2, 7, 6, 4, 4])            # label 5 absent from the style
        cm[0, :3, :2] = 9                                                       # 6 pixels: invalid (too few)
    elif kind == "stripes":
        cm = (np.arange(w)[None, None, :] // 20 % 5).astype(np.uint8).repeat(h, 1)
        sm = (np.arange(h)[None, :, None] // 16 % 5).astype(np.uint8).repeat(w, 2)
    else:
        cm = torch.randint(0, 4, (1, h, w), generator=g, dtype=torch.uint8).numpy()
        sm = torch.randint(0, 4, (1, h, w), generator=g, dtype=torch.uint8).numpy()
    ref = O.cwct_transfer_seg(zc, zs, cm, sm)
    n0 = _lib.launch_count()
    out = cWCT().transfer(zc.to(dev).clone(), zs.to(dev), torch.from_numpy(cm).to(dev), torch.from_numpy(sm).to(dev))
    assert _lib.launch_count() - n0 >= 4
    err = maxdiff(out, ref)
    print("masked TC stats %s %dx%d: %.2e" % (kind, h, w, err))
    assert err <= 1e-4


@pytest.mark.gpu
def test_pair_conv_cta_group2_vs_oracle():
    """conv_pair.cu (tcgen05 cta_group::2, off by default): the 64 -> 256 coupling conv on CTA pairs equals the oracle.
    The knob is read once per process, hence the subprocess; the profile table proves the pair kernel ran."""
    import subprocess, sys, os
    code = r'''
import torch
from oracle import vst_oracle as O
from tests.helpers import cpu_state_dict, fill_biases
from vstnet_b200 import RevResNet, _lib
torch.manual_seed(3)
net = RevResNet(hidden_dim=16, sp_steps=2).eval(); fill_biases(net, 7)
sd = cpu_state_dict(net); net = net.to("cuda:0")
for h, w in ((72, 136), (200, 1044), (36, 516)):
    x = torch.rand(1, 3, h, w, generator=torch.Generator().manual_seed(h))
    with torch.no_grad(): zr = O.revnet_forward(sd, x)
    _lib.profile_enable(True)
    z = net(x.to("cuda:0")); xr = net.inverse(z)
    _lib.profile_enable(False); torch.cuda.synchronize()
    prof = _lib.profile_collect()
    assert any("conv3x3_pair" in k for k in prof), sorted(prof)
    err = float((z.cpu() - zr).abs().max()); rt = float((xr.cpu() - x).abs().max())
    assert err <= 1e-3 and rt <= 4e-6, (h, w, err, rt)
    print("ok", h, w, err, rt)
'''
    env = dict(os.environ, VST_TC_PAIR="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
