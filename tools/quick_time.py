"""Quick per-stage timing of the hot path on one GPU (developer tool, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vstnet_b200 import RevResNet, cWCT, _lib

def ev_time(fn, warm=2, it=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it

def main():
    dev = torch.device("cuda:0")
    mode = sys.argv[1] if len(sys.argv) > 1 else "photo"
    H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1080, 1920)
    prec = sys.argv[4] if len(sys.argv) > 4 else "fp32"
    torch.manual_seed(0)
    kw = dict(hidden_dim=16, sp_steps=2) if mode == "photo" else dict(hidden_dim=64, sp_steps=1)
    net = RevResNet(**kw, precision=prec).to(dev).eval()
    x = torch.rand(1, 3, H, W, device=dev); s = torch.rand(1, 3, H, W, device=dev)
    cw = cWCT()
    z = net(x); zs = net(s)
    t_enc = ev_time(lambda: net(x))
    t_dec = ev_time(lambda: net(z, forward=False))
    t_wct = ev_time(lambda: cw.transfer(z, zs))
    flop = 609984.0 * H * W
    print("%s %dx%d %s: encode %.2f ms (%.1f TFLOP/s)  decode %.2f ms (%.1f TFLOP/s)  cwct %.3f ms" %
          (mode, H, W, prec, t_enc, flop / t_enc / 1e9, t_dec, flop / t_dec / 1e9, t_wct))
    n = z.shape[2] * z.shape[3]; C = z.shape[1]
    st = torch.cuda.current_stream().cuda_stream
    t_stats = ev_time(lambda: cw._stats(z[0], C, n, None, 1, st))
    cst = cw._stats(z[0], C, n, None, 1, st); sst = cw._stats(zs[0], C, n, None, 1, st)
    t_fac = ev_time(lambda: cw._factor(cst, [sst], [1.0], 0.0, C, 1, False, dev, st))
    T, mu, beta, valid = cw._factor(cst, [sst], [1.0], 0.0, C, 1, False, dev, st)
    out = torch.empty_like(z)
    t_app = ev_time(lambda: cw._apply(z[0], out[0], C, n, None, 1, T, mu, beta, valid, st))
    gb = 4.0 * C * n / 1e9
    print("  stats %.3f ms (%.0f GB/s)  factor %.3f ms  apply %.3f ms (%.0f GB/s)" %
          (t_stats, gb / t_stats * 1e3, t_fac, t_app, 2 * gb / t_app * 1e3))

if __name__ == "__main__":
    main()
