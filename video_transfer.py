#!/usr/bin/env python
"""video_transfer.py — the reference's video entry point (video_transfer.py:17-214) on vstnet_b200.

Same flags.  Differences: the style image is encoded and factorised once, not once per
frame (video_transfer.py:195 is loop-invariant); under ``torchrun --nproc-per-node N`` the frames are
rank-strided over N GPUs with one NCCL broadcast of the style statistics, and rank 0 writes the video in
frame order.  ``--synthetic HxWxF`` stylizes F random frames without any input files.
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
    out = {}

    def host_frames():
        for i in mine:
            frame = Image.fromarray(frames[i])
            frame = img_resize(frame, args.max_size, down_scale=RevNetwork.down_scale)  # ref :161
            yield torch.from_numpy(np.array(frame))

    # pipelined host path: upload, n_streams frames in flight on the GPU, download (vstnet_b200/video.py)
    for i, o in zip(mine, vs.stylize_stream(host_frames(), bgr=False)):
        out[i] = o.clone().numpy()                                                     # RGB uint8 HWC
    # gather frames on rank 0 in frame order (host side; no device collective on the data path)
    if world > 1:
        gathered = [None] * world
        dist.gather_object(out, gathered if rank == 0 else None, dst=0)
        if rank == 0:
            out = {k: v for d in gathered for k, v in d.items()}
    if rank == 0:
        import cv2
        path = os.path.join(args.out_dir, name)
        wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc('m', 'p', '4', 'v'), args.fps, (video_w, video_h))
        for i in range(len(frames)):
            f = out[i]
            if f.shape[0] != video_h or f.shape[1] != video_w:                          # ref :210 resize to video size
                f = np.array(Image.fromarray(f).resize((video_w, video_h), Image.BICUBIC))
            wr.write(np.ascontiguousarray(f[..., ::-1]))
        wr.release()
        print("Save at %s" % path)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
