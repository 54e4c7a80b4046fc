#!/usr/bin/env python
"""bench.py -- Mrays/s of the SVO trace path (primary + diffuse bounce) on B200.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W`
prints ONE JSON line on rank 0.  A "step" is one pass of the hot path over one
frame: render mode 0 of the reference shader (primary cast + one diffuse bounce
cast where the primary hit, src/shaders/svotrace.comp:443-560) at 1920x1080.
Rays = intersectOctree calls actually issued.

  value      whole-job Mrays/s with octree and frame parameters resident on the
             device (kernel launches only in the timed region)
  e2e        the same through the C ABI with HOST buffers: the frame struct comes
             from host memory each step and the colour + depth planes are read
             back into pinned host memory each step (what Main.java does per frame)
  roofline   algorithmic bytes of the reference layout (7 B root + every child
             record the reference fetches + 8 B/pixel output) / kernel time,
             against the measured HBM copy peak; plus the random-sector gather
             roofline of SURVEY 8d
  cpu_baseline  the CPU oracle (a port of the reference shader; the reference's
             own Java/GLSL cannot run here) on the host cores, bounded sample

`--impl reference` times that CPU oracle as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1920, 1080
CAM_CYCLE = ("A", "B", "C")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=int(os.environ.get("SVO_BENCH_SIZE", "8192")), help="world edge in voxels")
    ap.add_argument("--fast-math", type=int, default=0, help="1: fma-contracted kernels (not bit-exact)")
    ap.add_argument("--kernel", type=int, default=10, help="SVO_OPT_KERNEL (10 = the library default)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--camera", default="cycle", choices=["cycle", "A", "B", "C"], help="camera pose per step (default: A, B, C cycled)")
    ap.add_argument("--casts", type=int, default=2, help="mode-0 loop count (reference: 2 = primary + 1 diffuse bounce)")
    ap.add_argument("--mode", type=int, default=0, help="render mode (0 = the GI path of the metric; 2 = the engine's default)")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--band-rows", type=int, default=8, help="tiles: image rows per interleaved band (multiple of 8)")
    ap.add_argument("--fence", default="p2p", choices=["p2p", "nccl"], help="tiles: frame-complete fence = NVLink atomics or NCCL all-reduce")
    ap.add_argument("--partition", default="frames", choices=["frames", "tiles"],
                    help="N>1: frames = rank r renders progressive sample s*N+r of each view (weak scaling, no collective); "
                         "tiles = ONE frame per step split in interleaved 8-row bands, peers store straight into rank 0's "
                         "planes over NVLink (CUDA IPC) + one NCCL fence per frame (strong scaling)")
    return ap.parse_args()


def world(size: int, nthreads: int = 0):
    """Synthetic heightmap world, built by the product's own generator (host code; every rank builds its replica)."""
    import svo_raytracer_b200 as svo
    t0 = time.time()
    hm, mm = svo.terrain_inputs(size, nthreads=nthreads)
    nodes = svo.build_terrain(hm, mm, size, min(size, 1024), nthreads=nthreads)
    return nodes, time.time() - t0


CASTS, MODE = 2, 0


def frame_for(step: int, size: int):
    import svo_raytracer_b200 as svo
    depth = min(13, max(1, int(np.log2(size))))  # MAX_DEPTH = log2 N (13 for the reference's 8192^3, svotrace.comp:40)
    return svo.camera_frame(CAM_CYCLE[step % len(CAM_CYCLE)], frame_number=step + 1, render_mode=MODE,
                            max_depth=depth, casts=CASTS, cone_depth=11)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None
        self.t0 = self.t1 = None

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        import datetime
        sm, mx, reasons, all_sm = [], [], set(), []
        for line in self.f:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                clk, cmax = float(c[2]), float(c[3])
            except ValueError:
                continue
            all_sm.append(clk)
            mx.append(cmax)
            if self.t0 is not None and not (self.t0 - 0.02 <= ts <= self.t1 + 0.02):
                continue  # only samples taken DURING the timed region
            sm.append(clk)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[6:10]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:  # region shorter than the sampling period: fall back to the samples around it
            sm = all_sm[-3:]
        self.f.close()
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_arm(nodes, size, steps, warmup, budget_s, cores):
    """The CPU oracle (port of svotrace.comp) on the host cores; bounded sample of the same workload."""
    from oracle import oracle as O
    O.build()
    # bounded sample: every `stride`-th row block of each frame, sized from a pilot run
    def run(step, y0, y1):
        f = frame_for(step, size)
        of = O.make_frame(list(f.camPos), list(f.l1), list(f.l2), list(f.r1), list(f.r2), frame_number=f.frameNumber,
                          render_mode=f.renderMode, max_depth=f.maxDepth, casts=f.casts, cone_depth=f.coneDepth)
        t0 = time.perf_counter()
        _, st = O.render(nodes, of, W, H, y0=y0, y1=y1, nthreads=cores, planes=("rgba8", "depth"))
        return st.casts, time.perf_counter() - t0
    rows = 40
    pilot_rays, pilot_t = 0, 0.0
    for s in range(3):
        r, t = run(s, H // 2 - rows // 2, H // 2 + rows // 2)
        pilot_rays += r
        pilot_t += t
    per_step_budget = budget_s / max(1, steps + warmup)
    rows = int(max(8, min(H, rows * per_step_budget / max(pilot_t / 3, 1e-6))))
    y0 = (H - rows) // 2
    for s in range(warmup):
        run(s, y0, y0 + rows)
    rays, secs = 0, 0.0
    for s in range(steps):
        r, t = run(warmup + s, y0, y0 + rows)
        rays += r
        secs += t
    return rays / secs / 1e6, "rows [%d,%d) of each %dx%d frame, %d steps, cameras A/B/C cycled" % (y0, y0 + rows, W, H, steps), secs / steps * 1e3


def other_baselines():
    """The two baselines BASELINE.json's north star names next to the C port of castRay: probed, reported, never faked."""
    import ctypes.util
    import shutil
    egl = ctypes.util.find_library("EGL")
    java = shutil.which("java")
    return {
        "reference_glsl_via_egl": {"available": False, "libEGL": egl,
                                   "why": "the reference's shader sources are not on this box (they may not be copied into the repo) and no "
                                          "GL 4.3 harness can run them here" + ("" if egl else "; no libEGL either")},
        "java_castRay": {"available": False, "java": java,
                         "why": ("a JVM exists but " if java else "no JVM on this box; ") + "the reference has no CPU castRay: its traversal exists only as "
                                "GLSL, so the scalar multi-threaded C restatement (cpu_baseline) stands in"},
    }


def main():
    a = parse()
    global W, H, CASTS, MODE, CAM_CYCLE
    W, H, CASTS, MODE = a.width, a.height, a.casts, a.mode
    if a.camera != "cycle":
        CAM_CYCLE = (a.camera,) * 3
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    what = "primary + %d diffuse bounce%s" % (CASTS - 1, "s" if CASTS > 2 else "") if MODE == 0 else "render mode %d" % MODE
    workload = "%d^3 synthetic heightmap terrain SVO, %dx%d, render mode %d (%s), %s" % (
        a.size, W, H, MODE, what, "cameras A/B/C cycled" if a.camera == "cycle" else "camera " + a.camera)

    if a.impl == "reference":
        if rank != 0:
            return 0
        nodes, _ = world(a.size)
        v, sample, ms = cpu_arm(nodes, a.size, a.steps, a.warmup, 120.0, cores)
        print(json.dumps({
            "impl": "reference", "metric": "Mrays/s (primary + diffuse bounce)", "value": v, "unit": "Mrays/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": workload, "tree_bytes": int(nodes.size)},
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import torch
    import svo_raytracer_b200 as svo
    from svo_raytracer_b200 import _lib as L

    dist = None
    if world_size > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank if world_size > 1 else 0
    torch.cuda.set_device(dev)

    nodes, build_s = world(a.size, max(1, cores // world_size))
    ctx = svo.SvoContext(W, H, device=dev)
    ctx.set_option(L.OPT_FAST_MATH, a.fast_math)
    ctx.set_option(L.OPT_KERNEL, a.kernel)
    ctx.set_option(L.OPT_BAND_ROWS, a.band_rows)
    t0 = time.time()
    ctx.upload(nodes)
    upload_s = time.time() - t0
    info = ctx.scene_info()

    tiles = world_size > 1 and a.partition == "tiles"
    total = a.warmup + a.steps
    if tiles:
        # ONE frame per step, image bands interleaved over the ranks, replicated octree.  Rank 0 owns the frame
        # buffer; every peer maps it (CUDA IPC over NVLink) and its kernel stores its bands straight into it.
        frames = [frame_for(s, a.size) for s in range(total)]
        PL = (L.PLANE_COLOR_RGBA8, L.PLANE_DEPTH)
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        drain = lambda: None
        if a.fence == "nccl":
            handles = [ctx.ipc_export(p) for p in PL] if rank == 0 else [None, None]
            dist.broadcast_object_list(handles, src=0)
            if rank != 0:
                for plane, h in zip(PL, handles):
                    ctx.bind_plane(plane, ctx.ipc_import(h))
            fence = torch.zeros(1, device="cuda")

            def render_step(s, consume=None, release=False):
                ctx.render_interleaved(frames[s], rank, world_size)
                dist.all_reduce(fence)  # stream-ordered frame-complete fence: after it rank 0's planes hold the whole frame
                if consume is not None:
                    consume()
                if release:
                    dist.all_reduce(fence)  # the frame has been consumed: peers may overwrite the planes
        else:
            # No collective in the data path.  Rank 0 owns TWO colour/depth sets; frame k goes to set k&1, so two
            # frames are in flight.  Fences are counters in GPU memory bumped by remote atomics over NVLink
            # (svo_fence_*): slot 2+(k&1) of rank 0's counter = "bands of frame k stored" (complete at (k//2+1)*N),
            # slot 0 of every peer's counter = "frames consumed by rank 0".
            handles = [ctx.ipc_export(p | s) for s in (0, L.PLANE_BACK) for p in PL] if rank == 0 else [None] * 4
            dist.broadcast_object_list(handles, src=0)
            if rank == 0:
                sets = [[ctx.device_ptr(p | s) for p in PL] for s in (0, L.PLANE_BACK)]
            else:
                ptrs = [ctx.ipc_import(h) for h in handles]
                sets = [ptrs[0:2], ptrs[2:4]]
            fh = [None] * world_size
            dist.all_gather_object(fh, ctx.fence_export())
            if rank == 0:
                peer_fences = [ctx.ipc_import(fh[r]) for r in range(1, world_size)]
            else:
                owner_fence = [ctx.ipc_import(fh[0])]
            state = {"k": 0, "pending": None}

            def bind(k):
                for plane, ptr in zip(PL, sets[k & 1]):
                    ctx.bind_plane(plane, ptr)

            def finish(j, consume):
                ctx.fence_wait((j // 2 + 1) * world_size, slot=2 + (j & 1))  # every GPU has stored its bands of frame j
                if consume is not None:
                    bind(j)
                    consume()
                ctx.fence_signal(peer_fences, slot=0)  # frame j consumed: its plane set may be overwritten

            def render_step(s, consume=None, release=False):
                k = state["k"]
                state["k"] = k + 1
                if rank == 0:
                    bind(k)
                    ctx.render_interleaved(frames[s], 0, world_size)
                    ctx.fence_signal((), slot=2 + (k & 1))
                    if state["pending"] is not None:
                        finish(*state["pending"])
                    state["pending"] = (k, consume)
                else:
                    ctx.fence_wait(max(k - 1, 0), slot=0)  # rank 0 has consumed frames 0..k-2: set k&1 is free
                    bind(k)
                    ctx.render_interleaved(frames[s], rank, world_size)
                    ctx.fence_signal(owner_fence, slot=2 + (k & 1))

            def drain():
                if rank == 0 and state["pending"] is not None:
                    finish(*state["pending"])
                    state["pending"] = None
    else:
        # Units: rank r renders its own progressive sample (frameNumber) of the same views -- independent units, no
        # data-path collective (weak scaling); step s on rank r is sample s*N + r.
        def my_frame(s):
            f = frame_for(s, a.size)
            f.frameNumber = s * world_size + rank + 1
            return f
        frames = [my_frame(s) for s in range(total)]

        drain = lambda: None

        def render_step(s, consume=None, release=False):
            ctx.render(frames[s])
            if consume is not None:
                consume()

    # rays per frame and algorithmic bytes (instrumented kernel, outside the timed region; casts do not depend on the
    # RNG sample: a bounce is cast iff the primary ray hit)
    per_cam = {}
    for ci, cam in enumerate(CAM_CYCLE):
        per_cam[cam] = ctx.render_stats(frames[ci])
    rays_per_step = [per_cam[CAM_CYCLE[s % 3]]["casts"] for s in range(total)]
    alg_bytes_per_step = [per_cam[CAM_CYCLE[s % 3]]["record_bytes"] + 8 * W * H for s in range(total)]
    iters_per_step = [per_cam[CAM_CYCLE[s % 3]]["iters"] for s in range(total)]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(v):
        if dist is None:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing ------------------------------------------------
    clocks = ClockSampler(dev)  # started early: nvidia-smi needs ~100 ms before its first sample
    for s in range(a.warmup):
        render_step(s)
    drain()
    ctx.sync()
    barrier()
    launches0 = ctx.launch_count()
    clocks.begin()
    ctx.timer_begin()
    for s in range(a.warmup, total):
        render_step(s)
    drain()
    dev_ms = ctx.timer_end()
    clocks.end()
    launches = ctx.launch_count() - launches0
    barrier()
    clk = clocks.stop()
    max_ms = reduce_max(dev_ms)
    step_rays = float(sum(rays_per_step[a.warmup:]))
    all_rays = step_rays if tiles else step_rays * world_size  # frames mode: every rank casts a full frame per step
    value = all_rays / (max_ms * 1e-3) / 1e6

    # ---- end to end through the C ABI with host buffers -------------------------
    color_h = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
    depth_h = torch.empty((H, W), dtype=torch.float32).pin_memory()
    reads = rank == 0 or not tiles  # tiles: only rank 0 holds the frame

    def readback():
        ctx.read_plane_into(L.PLANE_COLOR_RGBA8, color_h.data_ptr(), color_h.numel())
        ctx.read_plane_into(L.PLANE_DEPTH, depth_h.data_ptr(), depth_h.numel() * 4)

    def e2e_step_sync(s):
        render_step(s, readback if reads else None, True)  # the 92-byte svo_frame is read from host memory by the call

    def timed(step_fn, finish=None):
        for s in range(a.warmup):
            step_fn(s)
        drain()
        if finish:
            finish()
        barrier()
        t0 = time.perf_counter()
        for s in range(a.warmup, total):
            step_fn(s)
        drain()
        if finish:
            finish()
        barrier()
        return reduce_max(time.perf_counter() - t0)

    e2e_sync_s = timed(e2e_step_sync)
    e2e_s = e2e_sync_s
    if not tiles:
        # the public pipelined path: two colour/depth sets on the device and two pinned sets on the host, frame s+1
        # renders while frame s crosses PCIe (svo_read_planes_async / svo_swap_buffers); every frame still lands in
        # host memory inside the timed region (svo_read_wait before the clock stops)
        color_h2, depth_h2 = torch.empty_like(color_h).pin_memory(), torch.empty_like(depth_h).pin_memory()
        host_sets = ((color_h, depth_h), (color_h2, depth_h2))

        def e2e_step_pipelined(s):
            ctx.render(frames[s])
            ch, dh = host_sets[s & 1]
            ctx.read_planes_async(ch.data_ptr(), dh.data_ptr())
            ctx.swap_buffers()
        e2e_s = timed(e2e_step_pipelined, ctx.read_wait)
    e2e_value = all_rays / e2e_s / 1e6

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline ----------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    timed = range(a.warmup, total)
    launch_ms = dev_ms / max(1, launches)
    alg_bytes = float(np.mean([alg_bytes_per_step[s] for s in timed]))
    achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
    mean_F = sum(iters_per_step[s] for s in timed) / sum(rays_per_step[s] for s in timed)
    gather = {}
    try:
        S = ctx.gather_probe(min(max(info["descriptors"] * 8, 1 << 20), 8 << 30))
        rays_s = sum(rays_per_step[s] for s in timed) / (dev_ms * 1e-3)
        gather = {"sectors_per_s": S, "working_set_bytes": info["descriptors"] * 8, "node_fetches_per_ray": mean_F,
                  "achieved_sector_equiv_per_s": rays_s * mean_F, "frac": rays_s * mean_F / S}
    except svo.SvoError as e:
        gather = {"error": str(e)}
    # DRAM traffic per launch from the committed `ncu --set full` capture of this same command (dram__bytes_read.sum +
    # dram__bytes_write.sum, mean of the three camera frames), profiles/r01_tile_full.json
    traffic = None
    tpath = os.path.join(ROOT, "profiles", {0: "r01_tile_full.json", 10: "r01_tile_wide_full.json"}.get(a.kernel, "none"))
    if os.path.exists(tpath) and a.size == 8192 and world_size == 1 and (W, H, CASTS, MODE) == (1920, 1080, 2, 0):
        caps = json.load(open(tpath))
        traffic = float(np.mean([c["dram_traffic_MB"] for c in caps])) * 1e6
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "kernel": {1: "k_render_persistent", 10: "k_render_tile_stack<wide>"}.get(a.kernel, "k_render_tile"), "algorithmic_bytes_per_launch": alg_bytes,
                "launch_ms": launch_ms, "gather": gather,
                "note": "the kernel is instruction-issue bound, not memory bound (ncu: issue slots 73 %, ALU pipe 74 %, DRAM 5 %): "
                        "upload-time transcoding removed the per-iteration record fetch the algorithmic-byte count assumes"}

    cpu = None
    if not a.no_cpu_baseline and world_size == 1:
        v, sample, _ = cpu_arm(nodes, a.size, 3, 0, a.cpu_seconds, cores)
        cpu = {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample}

    out = {
        "metric": "Mrays/s (primary + diffuse bounce)", "value": value, "unit": "Mrays/s", "n_gpus": world_size, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": max_ms / a.steps, "higher_is_better": True, "scaling": "strong" if tiles else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "partition": a.partition if world_size > 1 else "single GPU", "tree_bytes": int(nodes.size), "descriptors": info["descriptors"], "levels": info["levels"],
                   "l2_policy": "inputs larger than L2 (octree %.2f GB, 3 camera poses cycled); no flush between steps" % (nodes.size / 1e9)
                   if nodes.size > 126e6 else "octree fits L2; camera poses cycled; no flush",
                   "fast_math": a.fast_math, "kernel": a.kernel, "rays_per_step": {c: per_cam[c]["casts"] for c in CAM_CYCLE},
                   "world_build_s": round(build_s, 2), "upload_transcode_s": round(upload_s, 2),
                   "units": ("one frame per step, interleaved 8-row bands per rank, peers store into rank 0's planes over NVLink, "
                             "frame-complete fence = " + ("remote atomics over NVLink" if a.fence == "p2p" else "NCCL all-reduce")) if tiles else
                            "rank r renders progressive sample s*N+r of each view; no data-path collective"},
        "clocks": clk, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 92, "d2h_bytes_per_step": W * H * 8,
                "ms_per_step": e2e_s / a.steps * 1e3,
                "how": "svo_render + svo_read_planes_async + svo_swap_buffers per frame (read-back of frame s overlaps the render of "
                       "frame s+1), svo_read_wait inside the timed region" if not tiles else "rank 0 reads the assembled frame back after every frame-complete fence",
                "blocking_value": all_rays / e2e_sync_s / 1e6},
        "roofline": roofline,
    }
    if cpu is not None:
        out["cpu_baseline"] = cpu
        out["other_baselines"] = other_baselines()
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
