"""A/B of developer knobs on the resident-frame throughput path: python tools/fold_ab.py  (spawns itself per setting)."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child():
    import torch
    from vstnet_b200 import RevResNet, cWCT
    from vstnet_b200.video import VideoStylizer
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = RevResNet(hidden_dim=16, sp_steps=2).to(dev).eval()
    g = torch.Generator(device=dev)
    style = torch.rand(1, 3, 1080, 1920, device=dev, generator=g.manual_seed(1))
    frames = [torch.rand(1, 3, 1080, 1920, device=dev, generator=g.manual_seed(2 + i)) for i in range(4)]
    u8 = [(f[0].permute(1, 2, 0) * 255).byte().contiguous() for f in frames]
    for ns in (1, 4):
        vs = VideoStylizer(net, cWCT(), n_streams=ns)
        vs.set_style(style)
        for kind, src in (("f32", frames), ("u8dev", u8)):
            res = []
            for rep in range(4):
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in vs.stylize_frames(src[i % 4] for i in range(20)):
                    pass
                b.record()
                torch.cuda.synchronize()
                res.append(20 / a.elapsed_time(b) * 1e3)
            print("%s streams=%d %s fps: %s" % (os.environ.get("TAG"), ns, kind, " ".join("%.1f" % r for r in res)), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for tag, env in (("fold1", {"VST_FOLD0": "1"}), ("fold0", {"VST_FOLD0": "0"}), ("fold1_nopdl", {"VST_FOLD0": "1", "VST_PDL": "0"})):
            e = dict(os.environ, TAG=tag, **env)
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=e, check=False)
