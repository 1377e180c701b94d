"""CPU, world_size 2 over gloo: the multi-GPU host logic of the video path — frame sharding, the single
broadcast of the style statistics, and the ordered frame delivery of video_transfer.py (SharedFrameRing)."""
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
        from vstnet_b200.video import SharedFrameRing, broadcast_style_stats, shard_frames, style_buffer_bytes
        lib = _lib.load()
        pre = None
        nbytes = style_buffer_bytes(lib.vst_cwct_stats_bytes, 32, 3, 1)      # every rank sizes the buffer by itself
        if rank == 0:
            g = torch.Generator().manual_seed(11)
            n = int(lib.vst_cwct_stats_bytes(32, 3))
            pre = {"stats": [torch.randint(0, 256, (n,), generator=g, dtype=torch.uint8)], "L": 3, "masked": True, "C": 32}
        got = broadcast_style_stats(pre, nbytes, torch.device("cpu"))          # exactly one collective
        mine = shard_frames(n_frames, rank, world)
        # every rank "stylizes" its frames with a function of (frame index, style bytes); rank 0 takes them in frame
        # order from the shared-memory ring (2 slots per rank: producers must wait for the consumer)
        key = int(got["stats"][0].to(torch.int64).sum())
        path = os.path.join(out_dir, "ring")
        slots = 2 * world
        if rank == 0:
            ring = SharedFrameRing(path, (4, 6, 3), slots, create=True)
        dist.barrier()
        if rank != 0:
            ring = SharedFrameRing(path, (4, 6, 3), slots, create=False)
        ordered = []
        consumer = None
        if rank == 0:
            import threading

            def consume():
                for i in range(n_frames):
                    ordered.append(ring.get(i, timeout=60).copy())
                    ring.release(i)
            consumer = threading.Thread(target=consume)
            consumer.start()
        for i in mine:
            ring.put(i, np.full((4, 6, 3), (i * 7 + key) % 251, np.uint8), timeout=60)
        if consumer is not None:
            consumer.join()
        dist.barrier()
        ring.close(unlink=(rank == 0))
        torch.save({"meta": (got["L"], got["masked"], got["C"]), "bytes": got["stats"][0].clone(), "mine": mine,
                    "ordered": ordered, "key": key}, os.path.join(out_dir, "r%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames", [(2, 23), (2, 1)])
def test_style_broadcast_and_frame_sharding(tmp_path, world, n_frames):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_frames, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(str(tmp_path), "r%d.pt" % k), weights_only=False) for k in range(world)]
    assert r[0]["meta"] == r[1]["meta"] == (3, True, 32)
    assert torch.equal(r[0]["bytes"], r[1]["bytes"]), "style statistics differ between ranks after the broadcast"
    owned = sorted(i for k in range(world) for i in r[k]["mine"])
    assert owned == list(range(n_frames)), "every frame must have exactly one owner"
    got = r[0]["ordered"]
    assert len(got) == n_frames
    for i in range(n_frames):                       # delivered in frame order, rank-independent content
        assert got[i].shape == (4, 6, 3) and int(got[i][0, 0, 0]) == (i * 7 + r[0]["key"]) % 251
        assert (got[i] == got[i][0, 0, 0]).all()
    assert not os.path.exists(os.path.join(str(tmp_path), "ring"))


def test_style_stats_pack_roundtrip():
    from vstnet_b200.video import STYLE_HEADER_BYTES, pack_style_stats, unpack_style_stats
    g = torch.Generator().manual_seed(3)
    pre = {"stats": [torch.randint(0, 256, (96,), generator=g, dtype=torch.uint8) for _ in range(2)], "L": 5,
           "masked": False, "C": 16}
    buf = pack_style_stats(pre)
    assert buf.numel() == STYLE_HEADER_BYTES + 2 * 96
    back = unpack_style_stats(buf)
    assert (back["L"], back["masked"], back["C"]) == (5, False, 16)
    assert all(torch.equal(a, b) for a, b in zip(back["stats"], pre["stats"]))


def test_shard_frames_is_a_partition():
    from vstnet_b200.video import shard_frames
    for world in (1, 2, 4, 8):
        for n in (0, 1, 5, 240):
            parts = [shard_frames(n, r, world) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
