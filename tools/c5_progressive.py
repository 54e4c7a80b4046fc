#!/usr/bin/env python
"""BASELINE configs[4]: 8192^3 SVO, 8192x8192 image, progressive 64 samples per pixel, replicated octree, the image split in
interleaved 8-row bands over the ranks; every rank accumulates the running mean of ITS bands locally for all 64 samples
(svo_frame.flags bit 0 = the block the shader has commented out, svotrace.comp:712-719) and the bands are gathered to GPU 0
ONCE at the end with NCCL over NVLink (a SUM reduce of planes that are zero outside a rank's own bands).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/c5_progressive.py [--size 8192] [--res 8192] [--spp 64]
Prints one JSON line on rank 0 (Mrays/s over all ranks, gather time, a parity sample against the CPU oracle)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import svo_raytracer_b200 as svo
from svo_raytracer_b200 import _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=8192)
ap.add_argument("--res", type=int, default=8192)
ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--camera", default="B")
ap.add_argument("--check-rows", type=int, default=16)
a = ap.parse_args()
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W = H = a.res
depth = min(13, int(np.log2(a.size)))
hm, mm = svo.terrain_inputs(a.size, nthreads=max(1, (os.cpu_count() or 8) // world))
ctx = svo.SvoContext(W, H, device=local)
ctx.build_terrain_device(hm, mm, a.size, min(a.size, 1024))
# the context draws into torch-owned planes so that NCCL can reduce them
color = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
depthp = torch.zeros((H, W), dtype=torch.float32, device="cuda")
stream = torch.cuda.Stream()  # (torch's default stream is handle 0, which svo_set_stream reads as "the context's own stream")
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
ctx.bind_plane(L.PLANE_COLOR_RGBA8, color.data_ptr())
ctx.bind_plane(L.PLANE_DEPTH, depthp.data_ptr())
frames = [svo.camera_frame(a.camera, frame_number=s + 1, render_mode=0, max_depth=depth, flags=1) for s in range(a.spp)]
f0 = svo.camera_frame(a.camera, frame_number=1, render_mode=0, max_depth=depth)
rays_frame = ctx.render_stats(f0)["casts"]  # whole frame; a bounce is cast iff the primary ray hit (independent of the sample)
color.zero_(); depthp.zero_()
ctx.render_interleaved(frames[0], rank, world)  # warm-up
color.zero_(); depthp.zero_()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ev[0].record()
for s in range(a.spp):
    ctx.render_interleaved(frames[s], rank, world)
ev[1].record()
if world > 1:  # the ONE exchange: bands -> GPU 0
    dist.reduce(color, dst=0, op=dist.ReduceOp.SUM)
    dist.reduce(depthp, dst=0, op=dist.ReduceOp.SUM)
ev[2].record()
torch.cuda.synchronize()
t_render, t_gather = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
t = torch.tensor([t_render, t_render + t_gather], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    out = {"config": "BASELINE configs[4]: %d^3 SVO, %dx%d, progressive %d spp, %d GPU(s), interleaved 8-row bands, one NCCL gather at the end" % (
        a.size, W, H, a.spp, world), "n_gpus": world, "render_ms": float(t[0]), "render_plus_gather_ms": float(t[1]), "gather_ms": t_gather,
        "rays": int(rays_frame) * a.spp, "Mrays_per_s_render": rays_frame * a.spp / (float(t[0]) * 1e-3) / 1e6,
        "Mrays_per_s_with_gather": rays_frame * a.spp / (float(t[1]) * 1e-3) / 1e6, "gather_bytes": int((color.numel() + depthp.numel() * 4) * (world - 1) / world)}
    # parity sample: a few full-width rows of the accumulated image against the CPU oracle run for all samples
    if a.check_rows > 0:
        from oracle import oracle as O
        nodes = ctx.download()
        y0 = ((H // 2) // 8 + 3) * 8  # bands 515, 516: ranks 3 and 4 of 8 -- pixels that crossed NVLink in the gather
        y1 = y0 + a.check_rows
        pos, l1, l2, r1, r2 = svo.CAMERAS[a.camera]
        prev = np.zeros((H, W, 4), np.uint8)
        t0 = time.time()
        for s in range(a.spp):
            of = O.make_frame(pos, l1, l2, r1, r2, frame_number=s + 1, render_mode=0, max_depth=depth, flags=1)
            planes, _ = O.render(nodes, of, W, H, y0=y0, y1=y1, nthreads=os.cpu_count() or 8, planes=("rgba8", "depth"), prev_rgba8=prev)
            prev = planes["rgba8"]
        got_c = color[y0:y1].cpu().numpy()
        got_d = depthp[y0:y1].cpu().numpy()
        out["parity"] = {"rows": [y0, y1], "pixels": int((y1 - y0) * W), "rgba8_mismatch": int((got_c != prev[y0:y1]).any(axis=-1).sum()),
                         "depth_mismatch": int((got_d.view(np.uint32) != planes["depth"][y0:y1].view(np.uint32)).sum()), "oracle_s": round(time.time() - t0, 1)}
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
