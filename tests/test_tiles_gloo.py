"""Multi-GPU host logic on CPU: world_size-2 gloo run of the row-band partition and the gather to rank 0.
Each rank "renders" its bands with the CPU oracle (standing in for the device: the product has no CPU path),
rank 0 assembles the frame and compares it with the oracle's full frame."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_band_partition_covers_image_once():
    from svo_raytracer_b200 import tiles
    for height in (1080, 360, 37, 8, 3):
        for world in (1, 2, 3, 4, 8):
            for band in (4, 8, 16):
                seen = np.zeros(height, int)
                for r in range(world):
                    for y0, y1 in tiles.bands_of(r, world, height, band):
                        assert y0 % 4 == 0
                        seen[y0:y1] += 1
                assert (seen == 1).all()
    with pytest.raises(ValueError):
        tiles.row_bands(100, 6)


def test_pack_unpack_roundtrip():
    from svo_raytracer_b200 import tiles
    rng = np.random.default_rng(0)
    for height, world, band in ((45, 2, 8), (1080, 8, 8), (64, 3, 4)):
        img = torch.from_numpy(rng.integers(0, 255, (height, 17, 4)).astype(np.uint8))
        packed = [tiles.pack_bands(img, r, world, band) for r in range(world)]
        assert len({tuple(p.shape) for p in packed}) == 1  # equal-sized sends
        assert torch.equal(tiles.unpack_bands(packed, height, band), img)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import svo_raytracer_b200 as svo
        from oracle import oracle as O
        from svo_raytracer_b200 import tiles
        n, W, H, band = 64, 96, 52, 8
        hm, mm = svo.terrain_inputs(n)
        nodes = svo.build_terrain(hm, mm, n, 32)  # replicated octree: every rank builds/holds the same bytes
        pos, l1, l2, r1, r2 = svo.CAMERAS["B"]
        frame = O.make_frame(pos, l1, l2, r1, r2, frame_number=1, render_mode=2, max_depth=6)
        rgba = np.zeros((H, W, 4), np.uint8)
        depth = np.zeros((H, W), np.float32)
        for y0, y1 in tiles.bands_of(rank, world, H, band):
            out, _ = O.render(nodes, frame, W, H, y0=y0, y1=y1, planes=("rgba8", "depth"))
            rgba[y0:y1], depth[y0:y1] = out["rgba8"][y0:y1], out["depth"][y0:y1]
        full_rgba = tiles.gather_bands(torch.from_numpy(rgba), band)
        full_depth = tiles.gather_bands(torch.from_numpy(depth), band)
        if rank == 0:
            want, _ = O.render(nodes, frame, W, H, planes=("rgba8", "depth"))
            ok = bool(np.array_equal(full_rgba.numpy(), want["rgba8"]) and
                      np.array_equal(full_depth.numpy().view(np.uint32), want["depth"].view(np.uint32)))
            q.put(("ok" if ok else "mismatch", int((full_rgba.numpy() != want["rgba8"]).sum())))
        else:
            assert full_rgba is None and full_depth is None
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gather_matches_full_frame():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    status, bad = q.get(timeout=10)
    assert status == "ok", "assembled frame differs from the full frame in %d bytes" % bad
