#!/usr/bin/env python
"""video_transfer.py — the reference's video entry point (video_transfer.py:17-214) on vstnet_b200.

Same flags.  Differences: the style image is encoded and factorised once, not once per
frame (video_transfer.py:195 is loop-invariant); under ``torchrun --nproc-per-node N`` the frames are
rank-strided over N GPUs with one NCCL broadcast of the style statistics, and rank 0 writes the video in
frame order from a shared-memory frame ring (vstnet_b200.video.SharedFrameRing).  ``--synthetic HxWxF``
stylizes F random frames without any input files.  ``main`` returns the stylized uint8 RGB frames of this
rank as ``{frame index: array}`` (single process: all of them).
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from vstnet_b200.hostio import img_resize   # noqa: E402


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument('--mode', type=str, default='photorealistic')
    p.add_argument('--ckpoint', type=str, default='checkpoints/photo_video.pt')
    p.add_argument('--video', type=str, default='data/content/03.avi')
    p.add_argument('--style', type=str, default='data/style/03.jpeg')
    p.add_argument('--out_dir', type=str, default="output")
    p.add_argument('--max_size', type=int, default=1280)
    p.add_argument('--alpha_c', type=float, default=None)
    p.add_argument('--fps', type=int, default=10)
    p.add_argument('--auto_seg', action='store_true', default=False)
    p.add_argument('--precision', type=str, default='f16x2')
    p.add_argument('--synthetic', type=str, default=None, help="HxWxF: F random frames, random-init weights")
    return p


def read_frames(path):
    """ref: video_transfer.py:68-78 — a directory of images or a video file, as uint8 RGB arrays."""
    from PIL import Image
    if os.path.isdir(path):
        exts = ('.jpg', '.jpeg', '.png', '.ppm', '.bmp')
        files = sorted(os.path.join(d, f) for d, _, fs in os.walk(path) for f in fs if f.lower().endswith(exts))
        return [np.array(Image.open(f).convert('RGB')) for f in files]
    import cv2
    frames, cap = [], cv2.VideoCapture(path)
    while True:
        ret, frame = cap.read()
        if ret is False:
            break
        frames.append(np.ascontiguousarray(frame[..., ::-1]))
    return frames


def main(argv=None):
    import torch.distributed as dist
    from PIL import Image
    from image_transfer import build_network
    from vstnet_b200.video import VideoStylizer, shard_frames

    args = build_parser().parse_args(argv)
    if args.auto_seg:
        raise SystemExit("--auto_seg needs an external ADE20K segmenter (mmseg SegFormer), outside this path")
    if not torch.cuda.is_available():
        raise SystemExit("vstnet_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    os.makedirs(args.out_dir, exist_ok=True)

    RevNetwork = build_network(args.mode, args.precision)
    if args.synthetic:
        h, w, n = (int(v) for v in args.synthetic.lower().split("x"))
        h, w = h // 4 * 4, w // 4 * 4
        rng = np.random.default_rng(0)
        frames = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(n)]
        style = torch.rand(1, 3, h, w, generator=torch.Generator().manual_seed(1))
        name = "synthetic_%dx%dx%d.mp4" % (h, w, n)
    else:
        RevNetwork.load_state_dict(torch.load(args.ckpoint)['state_dict'])
        frames = read_frames(args.video)
        from torchvision import transforms
        style = img_resize(Image.open(args.style).convert('RGB'), args.max_size, down_scale=RevNetwork.down_scale)
        style = transforms.ToTensor()(style).unsqueeze(0)
        name = "%s_%s.mp4" % (os.path.basename(args.video).split(".")[0], os.path.basename(args.style).split(".")[0])
    RevNetwork = RevNetwork.to(device).eval()
    vs = VideoStylizer(RevNetwork, alpha_c=args.alpha_c)
    vs.set_style(style.to(device) if rank == 0 else None)

    video_h, video_w = frames[0].shape[:2]
    mine = shard_frames(len(frames), rank, world)
    down = int(RevNetwork.down_scale)

    def host_frames():
        for i in mine:
            frame = Image.fromarray(frames[i])
            frame = img_resize(frame, args.max_size, down_scale=down)                   # ref :161
            yield torch.from_numpy(np.array(frame))

    probe = img_resize(Image.fromarray(frames[0]), args.max_size, down_scale=down)
    same_size = (probe.size[1], probe.size[0]) == (video_h, video_w)

    def stylized_u8():
        """Stylized uint8 RGB HWC frames of this rank, in order, at the VIDEO's size.  When the network ran at the
        video's own size this is the pipelined uint8 path (vstnet_b200/video.py); otherwise the reference's order is
        kept — resize the float result to the video size (torchvision bicubic), then quantise (ref :208-212)."""
        if same_size:
            yield from vs.stylize_stream(host_frames(), bgr=False)
            return
        from torchvision import transforms
        to_video = transforms.Resize((video_h, video_w), interpolation=Image.BICUBIC)
        for f in host_frames():
            x = (f.to(device).permute(2, 0, 1)[None].float() / 255)
            y = to_video(vs.stylize(x))
            yield y[0].mul(255).clamp(0, 255).byte().permute(1, 2, 0).cpu()

    def write_video(get_frame):
        import cv2
        path = os.path.join(args.out_dir, name)
        wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc('m', 'p', '4', 'v'), args.fps, (video_w, video_h))
        for i in range(len(frames)):
            wr.write(np.ascontiguousarray(get_frame(i)[..., ::-1]))
        wr.release()
        print("Save at %s" % path)

    if world == 1:
        out = {}
        for i, o in zip(mine, stylized_u8()):
            out[i] = o.clone().numpy()
        write_video(lambda i: out[i])
        return out
    else:
        # ordered delivery through a bounded ring in shared host memory (no pickling, no device collective on the
        # data path): every rank copies its frames in, rank 0's writer thread takes them out in frame order
        import threading
        from vstnet_b200.video import SharedFrameRing
        tag = "%s_%s" % (os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RUN_ID", "run"))
        path = SharedFrameRing.default_path(tag)
        slots = 4 * world * vs.n_streams
        if rank == 0:
            ring = SharedFrameRing(path, (video_h, video_w, 3), slots, create=True)
        dist.barrier()
        if rank != 0:
            ring = SharedFrameRing(path, (video_h, video_w, 3), slots, create=False)
        writer, out = None, {}
        if rank == 0:
            def take(i):
                if i > 0:
                    ring.release(i - 1)
                return ring.get(i)
            writer = threading.Thread(target=write_video, args=(take,))
            writer.start()
        for i, o in zip(mine, stylized_u8()):
            ring.put(i, o)
            out[i] = o.clone().numpy()
        if writer is not None:
            writer.join()
            ring.release(len(frames) - 1)
        dist.barrier()
        ring.close(unlink=(rank == 0))
        dist.destroy_process_group()
        return out


if __name__ == "__main__":
    main()
