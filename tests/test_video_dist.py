"""CPU, world_size 2 over gloo: the multi-GPU host logic of the video path — frame sharding, the single
broadcast of the style statistics, and the frame-order gather of video_transfer.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import ROOT  # noqa: F401  (puts the repo on sys.path)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vstnet_b200 import _lib
        from vstnet_b200.video import broadcast_style_stats, shard_frames
        lib = _lib.load()
        nbytes_of = lambda C_, L: int(lib.vst_cwct_stats_bytes(C_, L))
        pre = None
        if rank == 0:
            g = torch.Generator().manual_seed(11)
            n = nbytes_of(32, 3)
            pre = {"stats": [torch.randint(0, 256, (n,), generator=g, dtype=torch.uint8)], "L": 3, "masked": True, "C": 32}
        got = broadcast_style_stats(pre, nbytes_of, torch.device("cpu"))
        mine = shard_frames(n_frames, rank, world)
        # every rank "stylizes" its frames with a function of (frame index, style bytes), then rank 0 gathers
        key = int(got["stats"][0].to(torch.int64).sum())
        out = {i: np.full((2, 2), (i * 7 + key) % 251, np.int64) for i in mine}
        gathered = [None] * world
        dist.gather_object(out, gathered if rank == 0 else None, dst=0)
        torch.save({"meta": (got["L"], got["masked"], got["C"]), "bytes": got["stats"][0], "mine": mine,
                    "gathered": gathered if rank == 0 else None, "key": key}, os.path.join(out_dir, "r%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames", [(2, 7), (2, 1)])
def test_style_broadcast_and_frame_sharding(tmp_path, world, n_frames):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_frames, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(str(tmp_path), "r%d.pt" % k), weights_only=False) for k in range(world)]
    assert r[0]["meta"] == r[1]["meta"] == (3, True, 32)
    assert torch.equal(r[0]["bytes"], r[1]["bytes"]), "style statistics differ between ranks after the broadcast"
    owned = sorted(i for k in range(world) for i in r[k]["mine"])
    assert owned == list(range(n_frames)), "every frame must have exactly one owner"
    merged = {k: v for d in r[0]["gathered"] for k, v in d.items()}
    assert sorted(merged) == list(range(n_frames))
    for i in range(n_frames):                       # rank-independent result for a given frame
        assert int(merged[i][0, 0]) == (i * 7 + r[0]["key"]) % 251


def test_shard_frames_is_a_partition():
    from vstnet_b200.video import shard_frames
    for world in (1, 2, 4, 8):
        for n in (0, 1, 5, 240):
            parts = [shard_frames(n, r, world) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
