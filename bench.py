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
             against the measured HBM copy peak; plus what actually bounds the
             kernel: issue slots and active lanes (from the ncu capture
             tools/profile_bench.sh makes of this same command) and the L1-miss
             sector rate against the measured gather rates
  cpu_baseline  the reference's own shader compiled for the CPU (oracle/_ref,
             kind "reference") -- or the C restatement (kind "port") where that
             library is absent -- on the host cores, bounded sample; its planes are
             compared with the frames the GPU has just been timed on ("parity")
  N > 1      ONE frame per step split into interleaved 8-row bands over the ranks
             (replicated octree); peers store straight into rank 0's planes over
             NVLink and the render kernel's last CTA bumps the frame-complete fence:
             strong scaling, the north-star layout.  `--partition frames` keeps the
             replica mode (rank r renders its own progressive sample; weak scaling).

`--impl reference` times the CPU arm alone (rank 0; no CUDA library is mapped).
"""
from __future__ import annotations

import argparse
import ctypes as C
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
CASTS, MODE = 2, 0
NO_HIT = 0xFFFFFFFF

# the three camera poses of SURVEY.md section 8d (== svo_raytracer_b200/cameras.py; repeated here so that the reference
# arm does not import the product package)
_FWD = ((-1.6, -0.9, -1.0), (-1.6, 0.9, -1.0), (1.6, -0.9, -1.0), (1.6, 0.9, -1.0))
_DOWN = ((-0.8, -1.0, -0.45), (-0.8, -1.0, 0.45), (0.8, -1.0, -0.45), (0.8, -1.0, 0.45))
CAMERAS = {"A": ((1.5, 1.5, 2.0),) + _FWD, "B": ((1.5, 1.3, 2.0),) + _FWD, "C": ((1.5, 1.6, 1.5),) + _DOWN}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=int(os.environ.get("SVO_BENCH_SIZE", "8192")), help="world edge in voxels")
    ap.add_argument("--fast-math", type=int, default=0, help="1: fma-contracted kernels (not bit-exact)")
    ap.add_argument("--kernel", type=int, default=-1, help="SVO_OPT_KERNEL (-1 = the library default)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--camera", default="cycle", choices=["cycle", "A", "B", "C"], help="camera pose per step (default: A, B, C cycled)")
    ap.add_argument("--casts", type=int, default=2, help="mode-0 loop count (reference: 2 = primary + 1 diffuse bounce)")
    ap.add_argument("--mode", type=int, default=0, help="render mode (0 = the GI path of the metric; 2 = the engine's default)")
    ap.add_argument("--mirror", type=int, default=0, help="svo_frame.mirrorValue (config 3: 4 with --world blobs)")
    ap.add_argument("--world", default="terrain", choices=["terrain"], help="synthetic world")
    ap.add_argument("--accumulate", type=int, default=0, help="1: progressive running mean (svo_frame.flags bit 0), frameNumber = step + 1")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--lanes", type=int, default=0, help="frames in flight per GPU (svo_select_lane; 1..7; 0 = 3 on one GPU, 6 in the tile partition)")
    ap.add_argument("--beam", type=int, default=-1, help="conservative beam pre-pass per frame (svo_beam_conservative + SVO_FRAME_BEAM_FLOOR; the frame does not "
                                                         "change): -1 = on for one GPU / replica mode in render modes 0 and 3, in the tile partition 2 while a rank "
                                                         "draws >= 500k pixels of the frame, else 0; 1 = every rank computes the whole lattice; "
                                                         "2 = tile partition with the lattice shared between the ranks (svo_beam_lattice_rows)")
    ap.add_argument("--band-rows", type=int, default=8, help="tiles: image rows per interleaved band (multiple of 8)")
    ap.add_argument("--rotate", type=int, default=0, help="tiles: 1 = rank r draws bands (r + k) mod N of frame k, so that with several frames in "
                                                          "flight every GPU sees the average band load instead of always the same bands "
                                                          "(measured at N = 8: +1.7 %% on the device, but the host-memory gather drops 10x when the "
                                                          "pages a rank writes change every frame: off)")
    ap.add_argument("--fence", default="p2p", choices=["p2p", "nccl"], help="tiles: frame-complete fence = NVLink atomics or NCCL all-reduce")
    ap.add_argument("--partition", default="auto", choices=["auto", "frames", "tiles"],
                    help="N>1: tiles (default) = ONE frame per step split in interleaved bands, peers store straight into rank 0's "
                         "planes over NVLink (CUDA IPC), frame-complete fence bumped by the render kernel (strong scaling); "
                         "frames = rank r renders progressive sample s*N+r of each view (weak scaling, no exchange)")
    ap.add_argument("--world-cache", default=os.environ.get("SVO_BENCH_WORLD_CACHE", "/dev/shm"),
                    help="directory for the generated world (built once per box, shared by ranks and by the reference arm); '' disables")
    ap.add_argument("--profile-json", default=os.path.join(ROOT, "profiles", "r02_bench_kernel.json"),
                    help="ncu summary of this command's render kernel (tools/profile_bench.sh) for roofline.issue / lanes / traffic")
    return ap.parse_args()


# ---- world ------------------------------------------------------------------------------------------------------
def _worldgen():
    """Host-only world generation (svo_terrain_generate, svo_build_terrain) from its own small library, so that the
    reference arm never maps the CUDA library."""
    path = os.path.join(ROOT, "svo_raytracer_b200", "libsvo_worldgen.so")
    if not os.path.exists(path):
        path = os.path.join(ROOT, "svo_raytracer_b200", "libsvo_b200.so")  # same functions (older build)
    L = C.CDLL(path)
    L.svo_terrain_generate.restype = C.c_int
    L.svo_terrain_generate.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.svo_build_terrain.restype = C.c_int
    L.svo_build_terrain.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_int]
    return L


def build_world(size: int, nthreads: int = 0) -> np.ndarray:
    L = _worldgen()
    hm = np.empty((size, size), np.uint16)
    mm = np.empty((size, size), np.uint8)
    if L.svo_terrain_generate(size, 1, hm.ctypes.data_as(C.c_void_p), mm.ctypes.data_as(C.c_void_p), nthreads) != 0:
        raise RuntimeError("svo_terrain_generate failed")
    need = C.c_uint64()
    chunk = min(size, 1024)
    if L.svo_build_terrain(hm.ctypes.data_as(C.c_void_p), mm.ctypes.data_as(C.c_void_p), size, chunk, None, 0, C.byref(need), nthreads) != 0:
        raise RuntimeError("svo_build_terrain(size) failed")
    out = np.empty(int(need.value), np.uint8)
    if L.svo_build_terrain(hm.ctypes.data_as(C.c_void_p), mm.ctypes.data_as(C.c_void_p), size, chunk, out.ctypes.data_as(C.c_void_p), out.size,
                           C.byref(need), nthreads) != 0:
        raise RuntimeError("svo_build_terrain failed")
    return out


def world(size: int, cache_dir: str, rank: int = 0, barrier=None, nthreads: int = 0):
    """The synthetic heightmap world (node stream in the engine's layout).  Built ONCE per box by whoever gets there first
    and kept in `cache_dir` (default /dev/shm): at N ranks rank 0 builds while the others wait, and later runs on the same
    box -- the other N of a scaling sweep, the reference arm -- map the file.  Returns (nodes, seconds, how)."""
    t0 = time.time()
    path = os.path.join(cache_dir, "svo_bench_world_terrain_%d_seed1.u8" % size) if cache_dir else None
    if path and os.path.exists(path):
        if barrier:
            barrier()
        return np.memmap(path, dtype=np.uint8, mode="r"), time.time() - t0, "cached"
    nodes = None
    if rank == 0:
        nodes = build_world(size, nthreads)
        if path:
            try:
                tmp = path + ".tmp.%d" % os.getpid()
                nodes.tofile(tmp)
                os.replace(tmp, path)
            except OSError:
                path = None
    if barrier:
        barrier()
    if nodes is None:
        nodes = np.memmap(path, dtype=np.uint8, mode="r") if path and os.path.exists(path) else build_world(size, nthreads)
    return nodes, time.time() - t0, "built"


def heightmaps(size: int, cache_dir: str, rank: int = 0, barrier=None, nthreads: int = 0):
    """The synthetic height and material maps (host code), generated once per box like world()."""
    t0 = time.time()
    path = os.path.join(cache_dir, "svo_bench_maps_terrain_%d_seed1.npz" % size) if cache_dir else None

    def load():
        z = np.load(path)
        return z["hm"], z["mm"]
    if path and os.path.exists(path):
        if barrier:
            barrier()
        hm, mm = load()
        return hm, mm, time.time() - t0, "cached"
    hm = mm = None
    if rank == 0:
        L = _worldgen()
        hm = np.empty((size, size), np.uint16)
        mm = np.empty((size, size), np.uint8)
        if L.svo_terrain_generate(size, 1, hm.ctypes.data_as(C.c_void_p), mm.ctypes.data_as(C.c_void_p), nthreads) != 0:
            raise RuntimeError("svo_terrain_generate failed")
        if path:
            try:
                tmp = path + ".tmp.%d.npz" % os.getpid()
                np.savez(tmp, hm=hm, mm=mm)
                os.replace(tmp, path)
            except OSError:
                path = None
    if barrier:
        barrier()
    if hm is None:
        if path and os.path.exists(path):
            hm, mm = load()
        else:
            hm, mm = heightmaps(size, "", 0, None, nthreads)[:2]
    return hm, mm, time.time() - t0, "generated"


def depth_for(size: int) -> int:
    return min(13, max(1, int(np.log2(size))))  # MAX_DEPTH = log2 N (13 for the reference's 8192^3, svotrace.comp:40)


def frame_params(step: int, size: int, a) -> dict:
    cam = CAM_CYCLE[step % len(CAM_CYCLE)]
    pos, l1, l2, r1, r2 = CAMERAS[cam]
    return dict(cam_pos=pos, l1=l1, l2=l2, r1=r1, r2=r2, frame_number=step + 1, render_mode=MODE, max_depth=depth_for(size), casts=CASTS,
                cone_depth=11, mirror_value=a.mirror, flags=1 if a.accumulate else 0)


def config_of(a, world_size: int, tree_bytes: int, partition: str) -> dict:
    """The workload, in the same words for both arms (the driver compares the two dicts)."""
    what = "primary + %d diffuse bounce%s" % (CASTS - 1, "s" if CASTS > 2 else "") if MODE == 0 else "render mode %d" % MODE
    return {
        "workload": "%d^3 synthetic heightmap terrain SVO, %dx%d, render mode %d (%s), %s" % (
            a.size, W, H, MODE, what, "cameras A/B/C cycled" if a.camera == "cycle" else "camera " + a.camera),
        "size": a.size, "width": W, "height": H, "render_mode": MODE, "casts": CASTS, "mirror_value": a.mirror,
        "accumulate": a.accumulate, "cameras": list(CAM_CYCLE) if a.camera == "cycle" else [a.camera], "tree_bytes": int(tree_bytes),
        "partition": partition if world_size > 1 else "single GPU",
        "l2_policy": ("inputs larger than L2 (octree %.2f GB, 3 camera poses cycled); no flush between steps" % (tree_bytes / 1e9))
        if tree_bytes > 126e6 else "octree fits L2; camera poses cycled; no flush",
    }


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md).  NVML is polled from a thread every ~4 ms so
    that even a 15 ms timed region (the driver's --steps 20) yields several samples; `nvidia-smi -lms 20` is the fallback."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.t0 = self.t1 = None
        self.samples = []  # (time, sm MHz, reasons bitmask)
        self.nvml = None
        self.p = None
        try:
            import threading
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.stop_flag = False
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
        except Exception:
            self.nvml = None
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            try:
                self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                           "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
            except OSError:
                self.p = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                clk = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                rs = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") else \
                    int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((time.time(), clk, rs))
            except Exception:
                pass
            time.sleep(0.004)

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.th.join(timeout=2)
            inside = [x for x in self.samples if self.t0 is not None and self.t0 - 0.002 <= x[0] <= self.t1 + 0.002]
            use = inside if inside else self.samples[-3:]
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            reasons = sorted({nm for _, _, rs in use for bit, nm in names.items() if rs & bit})
            return {"sm_mhz": float(np.median([c for _, c, _ in use])) if use else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(inside), "source": "nvml, polled every ~4 ms"}
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
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 20"}


# ---- CPU arm ----------------------------------------------------------------------------------------------------
def cpu_arm(nodes, a, steps, warmup, budget_s, cores, keep_planes=False):
    """The reference's own shader compiled for the CPU (oracle/_ref) -- or, where that library is absent, the C
    restatement -- on the host cores; bounded sample of the same workload.  Returns (Mrays/s, kind, sample text,
    ms per step, {step: (y0, y1, planes)})."""
    from oracle import oracle as O
    O.build()
    R = None
    if a.mirror == 0 and not a.accumulate:  # the shipped shader has neither (both are comments upstream)
        try:
            from oracle import ref as Rmod
            if Rmod.available():
                Rmod.lib()
                R = Rmod
        except Exception:
            R = None
    kind = "reference" if R is not None else "port"
    ray_cache = {}

    def rays_of(of, key, y0, y1):  # rays = casts issued; counted once per (camera, rows), untimed, by the restatement's counters
        if key not in ray_cache:
            _, st = O.render(nodes, of, W, H, y0=y0, y1=y1, nthreads=cores, planes=())
            ray_cache[key] = st.casts
        return ray_cache[key]

    def run(step, y0, y1, keep=False):
        fp = frame_params(step, a.size, a)
        of = O.make_frame(fp["cam_pos"], fp["l1"], fp["l2"], fp["r1"], fp["r2"], frame_number=fp["frame_number"], render_mode=fp["render_mode"],
                          max_depth=fp["max_depth"], casts=fp["casts"], cone_depth=fp["cone_depth"], mirror_value=fp["mirror_value"], flags=0)
        n = rays_of(of, (step % len(CAM_CYCLE), y0, y1), y0, y1)
        t0 = time.perf_counter()
        if R is not None:
            planes = R.render(nodes, of, W, H, y0=y0, y1=y1, nthreads=cores, planes=("rgba8", "depth"))
        else:
            planes, _ = O.render(nodes, of, W, H, y0=y0, y1=y1, nthreads=cores, planes=("rgba8", "depth"))
        return n, time.perf_counter() - t0, (planes if keep else None)

    rows = min(40, H)
    pilot_t = 0.0
    for s in range(3):
        _, t, _ = run(s, H // 2 - rows // 2, H // 2 - rows // 2 + rows)
        pilot_t += t
    per_step_budget = budget_s / max(1, steps + warmup)
    rows = int(max(8, min(H, rows * per_step_budget / max(pilot_t / 3, 1e-6))))
    y0 = (H - rows) // 2
    for s in range(warmup):
        run(s, y0, y0 + rows)
    rays, secs, kept = 0, 0.0, {}
    for s in range(steps):
        r, t, planes = run(warmup + s, y0, y0 + rows, keep_planes)
        rays += r
        secs += t
        if planes is not None:
            kept[warmup + s] = (y0, y0 + rows, planes)
    sample = "rows [%d,%d) of each %dx%d frame, %d steps, cameras %s" % (y0, y0 + rows, W, H, steps, "/".join(CAM_CYCLE[:3]))
    return rays / secs / 1e6, kind, sample, secs / steps * 1e3, kept


def other_baselines():
    """The two baselines BASELINE.json's north star names next to the CPU arm: probed, reported, never faked."""
    import ctypes.util
    import shutil
    egl = ctypes.util.find_library("EGL")
    java = shutil.which("java")
    return {
        "reference_glsl_via_egl": {"available": False, "libEGL": egl,
                                   "why": "no GL 4.3 compute context can be created on this box" + ("" if egl else " (no libEGL)") +
                                          "; the reference's GLSL runs instead on the CPU, compiled by g++ (cpu_baseline kind 'reference')"},
        "java_castRay": {"available": False, "java": java,
                         "why": ("a JVM exists but " if java else "no JVM on this box; ") + "the reference has no CPU castRay: its traversal exists only as "
                                "GLSL, which is what cpu_baseline runs"},
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
    partition = ("tiles" if a.partition == "auto" else a.partition) if max(world_size, a.gpus) > 1 else "single GPU"
    metric = "Mrays/s (primary + diffuse bounce)"

    if a.impl == "reference":
        if rank != 0:
            return 0
        nodes, build_s, how = world(a.size, a.world_cache)
        v, kind, sample, ms, _ = cpu_arm(nodes, a, a.steps, a.warmup, 120.0, cores)
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": v, "unit": "Mrays/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if partition == "tiles" else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_of(a, max(world_size, a.gpus), nodes.size, partition),
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample,
                             "what": "svotrace.comp compiled for the CPU by g++ (oracle/build_ref.py)" if kind == "reference" else "C restatement of svotrace.comp (oracle/svo_oracle.c)"},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "world": {"how": how, "seconds": round(build_s, 2)}}))
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

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = svo.SvoContext(W, H, device=dev)
    ctx.set_option(L.OPT_FAST_MATH, a.fast_math)
    if a.kernel >= 0:
        ctx.set_option(L.OPT_KERNEL, a.kernel)
    kernel_id = ctx.get_option(L.OPT_KERNEL)
    ctx.set_option(L.OPT_BAND_ROWS, a.band_rows)
    # The world: height / material maps on the host (once per box), then every rank builds its replica of the octree ON
    # ITS GPU (svo_build_terrain_device: the node stream is written straight into HBM and transcoded there).
    nodes = None
    rank_barrier = (lambda: dist.barrier()) if dist is not None else None
    hm, mm, maps_s, maps_how = heightmaps(a.size, a.world_cache, rank, rank_barrier, max(1, cores // world_size))
    t0 = time.time()
    try:
        tree_bytes = ctx.build_terrain_device(hm, mm, a.size, min(a.size, 1024))
        ctx.sync()
        world_how = "built on the device (svo_build_terrain_device), maps %s on the host" % maps_how
    except svo.SvoError:
        nodes, _, _ = world(a.size, a.world_cache, rank, rank_barrier, max(1, cores // world_size))
        ctx.upload(np.ascontiguousarray(nodes))
        tree_bytes = int(nodes.size)
        world_how = "host builder + svo_upload (shape not taken by the device builder)"
    upload_s = time.time() - t0
    build_s = maps_s + upload_s
    del hm, mm
    info = ctx.scene_info()


    tiles = world_size > 1 and partition == "tiles"
    if a.beam < 0:
        # one GPU / replicas: on.  Tile partition: the lattice is shared between the ranks (each traces 1/N of it for
        # everybody, one more fence per frame) while a rank's share of the frame is large enough to pay for that fence:
        # measured on 8 B200s (profiles/r02_sharedbeam_*.json) +7 % at N=2, +3 % at N=4 and at N=8 4K, -9 % at N=8 1080p.
        a.beam = (2 if W * H // world_size >= 500_000 and a.fence == "p2p" else 0) if tiles else 1
    BEAM = a.beam >= 1 and MODE in (0, 3) and not a.accumulate and not (tiles and a.fence == "nccl")
    ROTATE = bool(a.rotate) and tiles and not a.accumulate  # (the running mean of --accumulate lives in the planes a rank drew last time)
    BEAM_SHARED = BEAM and tiles and a.beam == 2  # every rank traces 1/N of the lattice for everybody (svo_beam_lattice_rows)

    def frame_for(s):
        fp = frame_params(s, a.size, a)
        if BEAM:
            fp["flags"] |= 2  # SVO_FRAME_BEAM_FLOOR
        return svo.make_frame(fp["cam_pos"], fp["l1"], fp["l2"], fp["r1"], fp["r2"], frame_number=fp["frame_number"], render_mode=fp["render_mode"],
                              max_depth=fp["max_depth"], casts=fp["casts"], cone_depth=fp["cone_depth"], mirror_value=fp["mirror_value"], flags=fp["flags"])

    total = a.warmup + a.steps
    # frames in flight per GPU: a frame's kernel ends with the critical path of its longest rays (~0.1 ms whatever share of the
    # frame the GPU renders); the next frames' tiles fill the SMs meanwhile (lanes = stream + plane set each)
    LANES = max(1, min(7, a.lanes if a.lanes > 0 else (6 if tiles else 3)))
    if a.accumulate:
        LANES = 1  # a running mean lives in ONE plane set
    PL = (L.PLANE_COLOR_RGBA8, L.PLANE_DEPTH)
    if tiles:
        # ONE frame per step, image bands interleaved over the ranks, replicated octree.  Rank 0 owns the frame
        # buffer; every peer maps it (CUDA IPC over NVLink) and its kernel stores its bands straight into it.
        frames = [frame_for(s) for s in range(total + 3)]
        drain = lambda: None
        if a.fence == "nccl":
            nccl_stream = torch.cuda.Stream()  # (torch's default stream is handle 0 = "the context's own stream" to svo_set_stream)
            torch.cuda.set_stream(nccl_stream)
            ctx.set_stream(nccl_stream.cuda_stream)  # the all-reduce runs on torch's stream
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
            # No collective in the data path.  Rank 0 owns LANES colour/depth sets; frame k goes to set k % LANES, so LANES
            # frames are in flight.  Fences are counters in GPU memory bumped by remote atomics over NVLink:
            # slot 1+(k%LANES) of rank 0's counter = "bands of frame k stored" (complete at (k//LANES+1)*N) -- bumped by the
            # LAST CTA of each rank's render kernel (svo_render_interleaved_signal: one launch per rank and frame) --
            # slot 0 of every peer's counter = "frames consumed by rank 0".
            handles = [ctx.ipc_export(p | (l << 8)) for l in range(LANES) for p in PL] if rank == 0 else [None] * (2 * LANES)
            dist.broadcast_object_list(handles, src=0)
            if rank == 0:
                sets = [[ctx.device_ptr(p | (l << 8)) for p in PL] for l in range(LANES)]
            else:
                ptrs = [ctx.ipc_import(h) for h in handles]
                sets = [ptrs[2 * l:2 * l + 2] for l in range(LANES)]
            fh = [None] * world_size
            dist.all_gather_object(fh, ctx.fence_export())
            if rank == 0:
                peer_fences = [ctx.ipc_import(fh[r]) for r in range(1, world_size)]
            else:
                owner_fence = [ctx.ipc_import(fh[0])]
            state = {"k": 0, "pending": None}
            if BEAM_SHARED:
                lat_h = [None] * world_size
                dist.all_gather_object(lat_h, [ctx.ipc_export(L.PLANE_BEAM_LATTICE | (l << 8)) for l in range(LANES)])
                lat_ptrs = [[ctx.device_ptr(L.PLANE_BEAM_LATTICE | (l << 8)) if r == rank else ctx.ipc_import(lat_h[r][l]) for r in range(world_size)]
                            for l in range(LANES)]
                all_fences = [ctx.fence_device_ptr() if r == rank else ctx.ipc_import(fh[r]) for r in range(world_size)]
                lat_rows = H // 4 + 1
                my_rows = (rank * lat_rows // world_size, (rank + 1) * lat_rows // world_size)

            def beam_for(s, k):
                if not BEAM:
                    return
                if not BEAM_SHARED:
                    ctx.beam_conservative(frames[s])  # every rank computes the (small) beam plane of the whole frame for itself
                    return
                l = k % LANES
                ctx.beam_lattice_rows(frames[s], my_rows[0], my_rows[1], lat_ptrs[l], all_fences, slot=8 + l)
                ctx.fence_wait((k // LANES + 1) * world_size, slot=8 + l)  # every rank's rows of this frame's lattice have arrived
                ctx.beam_filter()

            def bind(k):
                for plane, ptr in zip(PL, sets[k % LANES]):
                    ctx.bind_plane(plane, ptr)

            # Frame k lives on lane k % LANES (stream + plane set): the kernels of the next frames start while frame k's last tiles drain.
            def finish(j, consume):
                ctx.select_lane(j % LANES)
                if consume is None:  # device-resident loop: "every GPU has stored frame j" -> "frame j consumed" in one launch
                    ctx.fence_wait_signal((j // LANES + 1) * world_size, 1 + (j % LANES), peer_fences, 0)
                    return
                ctx.fence_wait((j // LANES + 1) * world_size, slot=1 + (j % LANES))  # every GPU has stored its bands of frame j
                bind(j)
                consume()
                ctx.fence_signal(peer_fences, slot=0)  # frame j consumed: its plane set may be overwritten

            def render_step(s, consume=None, release=False):
                k = state["k"]
                state["k"] = k + 1
                ctx.select_lane(k % LANES)
                part = (rank + k) % world_size if ROTATE else rank
                if rank == 0:
                    beam_for(s, k)
                    bind(k)
                    ctx.render_interleaved_signal(frames[s], part, world_size, (), slot=1 + (k % LANES))
                    if state["pending"] is not None:
                        finish(*state["pending"])
                    state["pending"] = (k, consume)
                else:
                    ctx.fence_wait(max(k - LANES + 1, 0), slot=0)  # rank 0 has consumed frames 0..k-LANES: this lane's set is free
                    beam_for(s, k)
                    bind(k)
                    ctx.render_interleaved_signal(frames[s], part, world_size, owner_fence, slot=1 + (k % LANES))

            def drain():
                if rank == 0 and state["pending"] is not None:
                    finish(*state["pending"])
                    state["pending"] = None
    else:
        # Units: rank r renders its own progressive sample (frameNumber) of the same views -- independent units, no
        # data-path exchange (weak scaling); step s on rank r is sample s*N + r.
        def my_frame(s):
            f = frame_for(s)
            if not a.accumulate:
                f.frameNumber = s * world_size + rank + 1
            return f
        frames = [my_frame(s) for s in range(total + 3)]

        drain = lambda: None

        def render_step(s, consume=None, release=False):
            if LANES > 1:
                ctx.select_lane(s % LANES)  # consecutive frames on alternating lanes (stream + plane set): frame s+1 may start while frame s's last tiles drain
            if BEAM:
                ctx.beam_conservative(frames[s])  # per-block lower bounds on the primary hit distance (this lane's beam plane)
            ctx.render(frames[s])
            if consume is not None:
                consume()

    # rays per frame and algorithmic bytes (instrumented kernels, outside the timed region; casts do not depend on the
    # RNG sample: a bounce is cast iff the primary ray hit)
    per_cam, per_cam_exec = {}, {}
    for ci, cam in enumerate(CAM_CYCLE):
        f = frame_for(ci)
        f.flags = 2 if BEAM else 0
        per_cam[cam] = ctx.render_stats(f)
        if BEAM:
            ctx.beam_conservative(f)
        per_cam_exec[cam] = ctx.render_stats_executed(f)
    rays_per_step = [per_cam[CAM_CYCLE[s % 3]]["casts"] for s in range(total)]
    alg_bytes_per_step = [per_cam[CAM_CYCLE[s % 3]]["record_bytes"] + 8 * W * H for s in range(total)]
    iters_per_step = [per_cam[CAM_CYCLE[s % 3]]["iters"] for s in range(total)]
    exec_iters_per_step = [per_cam_exec[CAM_CYCLE[s % 3]]["iters"] for s in range(total)]

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
    ctx.sync()  # raises if a frame-complete fence ran into its watchdog: the frames would be incomplete
    barrier()
    clk = clocks.stop()
    max_ms = reduce_max(dev_ms)
    step_rays = float(sum(rays_per_step[a.warmup:]))
    all_rays = step_rays if (tiles or world_size == 1) else step_rays * world_size  # frames mode: every rank casts a full frame per step
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
        host_sets = [(color_h, depth_h)] + [(torch.empty_like(color_h).pin_memory(), torch.empty_like(depth_h).pin_memory()) for _ in range(max(LANES, 2) - 1)]

        def e2e_step_pipelined(s):
            l = s % len(host_sets)
            ctx.select_lane(l)  # (waits, on the device, until this lane's previous read-back has left its planes)
            if BEAM:
                ctx.beam_conservative(frames[s])
            ctx.render(frames[s])
            ch, dh = host_sets[l]
            ctx.read_planes_async(ch.data_ptr(), dh.data_ptr())
        e2e_s = timed(e2e_step_pipelined, ctx.read_wait)
    e2e_value = all_rays / e2e_s / 1e6

    # ---- tiles: end to end with the frame assembled in HOST memory ----------------------------------------------------
    # Through GPU 0 (above) every frame crosses ONE PCIe link (16.6 MB per 1080p frame: ~0.4 ms, more than the 8-GPU render
    # time).  Here every rank renders its bands into its own planes and copies them over ITS link straight to their place in
    # frame buffers that live in host memory shared by the ranks (pinned by each): svo_read_interleaved_async.
    e2e_host_value, got_host = None, {}
    if tiles and a.fence == "p2p":
        for plane in PL:
            ctx.bind_plane(plane, None)
        set_bytes = W * H * 8
        shm_path = "/dev/shm/svo_bench_frames_%s" % os.environ.get("MASTER_PORT", "0")
        if rank == 0:
            with open(shm_path, "wb") as fh:
                fh.truncate(LANES * set_bytes)
        dist.barrier()
        shm = torch.from_file(shm_path, shared=True, size=LANES * set_bytes, dtype=torch.uint8)
        rc = torch.cuda.cudart().cudaHostRegister(shm.data_ptr(), shm.numel(), 0)
        ok = torch.tensor([1 if int(rc) == 0 else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 1:
            base = shm.data_ptr()
            cptr = [base + l * set_bytes for l in range(LANES)]
            dptr = [base + l * set_bytes + W * H * 4 for l in range(LANES)]
            hstate = {"k": 0}

            def e2e_host_step(s):
                l = hstate["k"] % LANES
                hstate["k"] += 1
                ctx.select_lane(l)  # (waits, on the device, until this lane's previous copy has left its planes)
                part = (rank + hstate["k"]) % world_size if ROTATE else rank
                if BEAM:
                    ctx.beam_conservative(frames[s])
                ctx.render_interleaved(frames[s], part, world_size)
                ctx.read_interleaved_async(part, world_size, cptr[l], dptr[l])
            e2e_host_s = timed(e2e_host_step, ctx.read_wait)
            e2e_host_value = all_rays / e2e_host_s / 1e6
            for ci in range(len(set(CAM_CYCLE))):  # the assembled host frames, for the parity check
                s = a.warmup + ci
                if s >= total:
                    break
                ctx.select_lane(0)
                if BEAM:
                    ctx.beam_conservative(frames[s])
                ctx.render_interleaved(frames[s], rank, world_size)
                ctx.read_interleaved_async(rank, world_size, cptr[0], dptr[0])
                ctx.read_wait()
                barrier()
                if rank == 0:
                    hv = shm[:set_bytes].numpy()
                    got_host[s] = (hv[:W * H * 4].reshape(H, W, 4).copy(), hv[W * H * 4:].view(np.float32).reshape(H, W).copy())
                barrier()
            torch.cuda.cudart().cudaHostUnregister(shm.data_ptr())
        del shm
        barrier()
        if rank == 0:
            try:
                os.unlink(shm_path)
            except OSError:
                pass

    # ---- the frames just timed, kept for the parity check (outside the timed region) ------------------------------
    got_frames = {}
    if not a.accumulate:
        for ci in range(len(set(CAM_CYCLE))):
            s = a.warmup + ci  # a timed step of each camera
            if s >= total:
                break
            render_step(s, readback if reads else None, True)
            drain()
            if reads:
                ctx.sync()
                got_frames[s] = (color_h.numpy().copy(), depth_h.numpy().copy())

    # ---- tiles: where does the step time go? ----------------------------------------------------------------------
    scaling_detail = None
    if tiles:
        barrier()
        per_rank = []
        for ci in range(3):  # this rank's share of each camera's frame, alone on its stream (no fences)
            if BEAM:
                ctx.beam_conservative(frames[ci])
            ctx.render_interleaved(frames[ci], rank, world_size)
            ctx.sync()
            ctx.timer_begin()
            for _ in range(5):
                ctx.render_interleaved(frames[ci], rank, world_size)
            per_rank.append(ctx.timer_end() / 5)
        t = torch.tensor(per_rank, device="cuda", dtype=torch.float64)
        gathered = [torch.zeros_like(t) for _ in range(world_size)]
        dist.all_gather(gathered, t)
        shares = np.array([g.cpu().numpy() for g in gathered])  # [rank, camera]
        barrier()
        single = []
        if rank == 0:  # the whole frame on one GPU with the same kernel: the strong-scaling reference point
            for plane in PL:
                ctx.bind_plane(plane, None)
            for ci in range(3):
                if BEAM:
                    ctx.beam_conservative(frames[ci])
                ctx.render_interleaved(frames[ci], 0, 1)
                ctx.sync()
                ctx.timer_begin()
                for _ in range(5):
                    ctx.render_interleaved(frames[ci], 0, 1)
                single.append(ctx.timer_end() / 5)
        barrier()
        scaling_detail = {
            "share_kernel_ms_by_camera": {c: {"min": float(shares[:, i].min()), "mean": float(shares[:, i].mean()), "max": float(shares[:, i].max())}
                                          for i, c in enumerate(CAM_CYCLE)},
            "share_kernel_ms_max_mean": float(shares.max(axis=0).mean()),
            "single_gpu_frame_ms_by_camera": {c: float(single[i]) for i, c in enumerate(CAM_CYCLE)} if single else None,
            "single_gpu_frame_ms_mean": float(np.mean(single)) if single else None,
            "nvlink_bytes_per_frame": int(8 * W * H * (world_size - 1) // world_size),
            "launches_per_step_rank0": launches / max(1, a.steps),
            "note": "step time - max share kernel time = launch + fence latency and inter-GPU skew; max share x N - single-GPU frame = load "
                    "imbalance between the interleaved bands plus the per-launch tail",
        }

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- parity of the timed frames ------------------------------------------------------------------------------
    parity = None
    cpu = None
    if world_size == 1 and not a.no_cpu_baseline:
        if nodes is None:
            nodes = ctx.download()  # the CPU arm traces the stream the device built
        v, kind, sample, _, kept = cpu_arm(nodes, a, 3, a.warmup, a.cpu_seconds, cores, keep_planes=True)
        cpu = {"value": v, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample,
               "what": "svotrace.comp compiled for the CPU by g++ (oracle/build_ref.py)" if kind == "reference" else "C restatement of svotrace.comp (oracle/svo_oracle.c)"}
        if got_frames:
            px = bad_c = bad_d = 0
            for s, (y0, y1, planes) in kept.items():
                if s not in got_frames:
                    continue
                gc, gd = got_frames[s]
                bad_c += int((gc[y0:y1] != planes["rgba8"][y0:y1]).any(axis=-1).sum())
                bad_d += int((gd[y0:y1].view(np.uint32) != planes["depth"][y0:y1].view(np.uint32)).sum())
                px += (y1 - y0) * W
            parity = {"against": "cpu_baseline planes (kind %s) of the same frames" % kind, "frames": len(kept), "pixels": px,
                      "rgba8_mismatch": bad_c, "depth_mismatch": bad_d}
    elif tiles and got_frames:
        # the frame assembled from N GPUs' bands against the same frame rendered by GPU 0 alone (itself checked against
        # the CPU arm by the N = 1 run and by tests/test_gpu_parity.py::test_full_size_properties)
        for plane in PL:
            ctx.bind_plane(plane, None)
        px = bad_c = bad_d = 0
        def alone(s):  # the whole frame on GPU 0, without the beam floor
            f = frame_for(s)
            f.flags = 0
            ctx.render(f)
            return ctx.read_color_rgba8(), ctx.read_depth()
        for s, (gc, gd) in got_frames.items():
            wc, wd = alone(s)
            bad_c += int((gc != wc).any(axis=-1).sum())
            bad_d += int((gd.view(np.uint32) != wd.view(np.uint32)).sum())
            px += W * H
        for s, (gc, gd) in got_host.items():  # ... and the frames assembled in host memory from every rank's own read-back
            wc, wd = alone(s)
            bad_c += int((gc != wc).any(axis=-1).sum())
            bad_d += int((gd.view(np.uint32) != wd.view(np.uint32)).sum())
            px += W * H
        parity = {"against": "the same frames rendered by GPU 0 alone", "frames": len(got_frames) + len(got_host), "pixels": px, "rgba8_mismatch": bad_c,
                  "depth_mismatch": bad_d, "assembled": "in GPU 0's planes over NVLink (%d frames) and in shared host memory (%d frames)" % (len(got_frames), len(got_host))}

    # ---- roofline ----------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    timed_steps = range(a.warmup, total)
    render_launches = a.steps  # one render kernel per step on this rank (fence kernels of the tile partition are not counted here)
    launch_ms = dev_ms / max(1, render_launches)
    alg_bytes = float(np.mean([alg_bytes_per_step[s] for s in timed_steps])) / (world_size if tiles else 1)
    achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
    n_rays = sum(rays_per_step[s] for s in timed_steps)
    mean_F = sum(iters_per_step[s] for s in timed_steps) / n_rays
    mean_F_exec = sum(exec_iters_per_step[s] for s in timed_steps) / n_rays
    kernel_name = {1: "k_render_persistent", 10: "k_render_tile_stack<wide>", 13: "k_render_tile_balanced", 17: "k_render_tile_queue"}.get(kernel_id, "k_render_tile")
    if tiles:
        kernel_name = "k_render_tile_queue"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "kernel": kernel_name, "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": launch_ms,
                "node_fetches_per_ray_reference": mean_F, "loop_iterations_per_ray_executed": mean_F_exec,
                "note": "the kernel is instruction-issue bound, not memory bound: upload-time transcoding removed the per-iteration record "
                        "fetch the algorithmic-byte count assumes (real DRAM traffic = `traffic`), and the content box skips the empty-space "
                        "iterations between the reference's count and the executed one.  `issue` and `lanes` are the bounds that bind."}
    # ncu capture of this same command's render kernel (tools/profile_bench.sh writes it; one row per camera frame)
    prof = None
    if os.path.exists(a.profile_json) and world_size == 1 and (a.size, W, H, CASTS, MODE) == (8192, 1920, 1080, 2, 0):
        try:
            rows = [r for r in json.load(open(a.profile_json)) if kernel_name.split("<")[0] in r.get("kernel", "")]
            if rows:
                prof = {k: float(np.mean([r[k] for r in rows if k in r])) for k in rows[0] if isinstance(rows[0][k], (int, float))}
                prof["rows"] = len(rows)
        except (ValueError, OSError, KeyError):
            prof = None
    if prof:
        sm_clk = (clk.get("sm_mhz") or 1965.0) * 1e6
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        peak_issue = sms * 4 * sm_clk
        roofline["traffic"] = prof.get("dram_traffic_MB", 0.0) * 1e6
        roofline["profile"] = os.path.relpath(a.profile_json, ROOT)
        if "warp_instructions" in prof:
            ach = prof["warp_instructions"] / (launch_ms * 1e-3)
            roofline["issue"] = {"warp_instructions_per_launch": prof["warp_instructions"], "achieved": ach, "peak": peak_issue,
                                 "unit": "warp instructions/s", "frac": ach / peak_issue,
                                 "issue_slot_util_pct_ncu": prof.get("issue_slot_util_pct"),
                                 "how": "smsp__inst_executed.sum per launch (ncu) / live launch time, against SMs x 4 schedulers x live SM clock"}
        if "active_lanes_per_inst" in prof:
            roofline["lanes"] = {"active_lanes_per_inst": prof["active_lanes_per_inst"], "frac": prof["active_lanes_per_inst"] / 32.0}
    gather = {}
    try:  # SURVEY 8d's gather roofline: random 32-byte sectors, L2-resident and HBM-resident working sets
        S_hbm = ctx.gather_probe(min(max(info["descriptors"] * 8, 1 << 20), 8 << 30))
        S_l2 = ctx.gather_probe(48 << 20)
        rays_s = n_rays / (dev_ms * 1e-3)
        gather = {"sectors_per_s_hbm_resident": S_hbm, "sectors_per_s_l2_resident": S_l2, "working_set_bytes": info["descriptors"] * 8,
                  "reference_fetches_per_s": rays_s * mean_F}
        if prof and "l1_miss_sectors" in prof:
            # sectors the kernel really requests from L2 per launch, split by where they were served, against the two rates
            miss = prof["l1_miss_sectors"]
            l2_hit = prof.get("l2_hit_pct", 60.0) / 100.0
            t_bound = miss * l2_hit / S_l2 + miss * (1.0 - l2_hit) / S_hbm
            gather.update({"l1_miss_sectors_per_launch": miss, "l2_hit_rate": l2_hit, "bound_ms": t_bound * 1e3,
                           "frac": t_bound / (launch_ms * 1e-3),
                           "how": "time the measured gather rates need for the kernel's L1-miss sectors (ncu) / live launch time"})
    except svo.SvoError as e:
        gather = {"error": str(e)}
    roofline["gather"] = gather

    out = {
        "metric": metric, "value": value, "unit": "Mrays/s", "n_gpus": world_size, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": max_ms / a.steps, "higher_is_better": True, "scaling": "strong" if tiles else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_of(a, world_size, tree_bytes, partition),
        "impl_details": {"kernel": kernel_id, "fast_math": a.fast_math, "frames_in_flight": LANES, "band_rotation": ROTATE, "conservative_beam_prepass": ("lattice shared between the ranks" if BEAM_SHARED else bool(BEAM)), "descriptors": info["descriptors"], "levels": info["levels"],
                         "rays_per_step": {c: per_cam[c]["casts"] for c in CAM_CYCLE},
                         "world": {"how": world_how, "seconds": round(build_s, 2), "maps_s": round(maps_s, 2)}, "build_and_transcode_s": round(upload_s, 3),
                         "units": ("one frame per step, interleaved %d-row bands per rank, peers store into rank 0's planes over NVLink, "
                                   "frame-complete fence = %s" % (a.band_rows, "remote atomics over NVLink, bumped by the render kernel's last CTA" if a.fence == "p2p" else "NCCL all-reduce")) if tiles else
                                  ("rank r renders progressive sample s*N+r of each view; no data-path exchange" if world_size > 1 else "one GPU renders every frame")},
        "clocks": clk, "gpu_launches": launches,
        "e2e": {"value": e2e_host_value if e2e_host_value is not None else e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 92, "d2h_bytes_per_step": W * H * 8,
                "ms_per_step": (all_rays / (e2e_host_value * 1e6) if e2e_host_value is not None else e2e_s) / a.steps * 1e3,
                "how": "svo_select_lane + svo_render + svo_read_planes_async per frame (the read-back of frame s overlaps the render of "
                       "the next frames), svo_read_wait inside the timed region" if not tiles else
                       ("every rank renders its bands and copies them over its own PCIe link to their place in frame buffers in host memory shared by "
                        "the ranks (svo_render_interleaved + svo_read_interleaved_async, %d frames in flight); every frame is in host memory when the "
                        "clock stops" % LANES if e2e_host_value is not None else
                        "rank 0 reads the assembled frame back after every frame-complete fence"),
                "via_gpu0_value": e2e_value if tiles else None,
                "via_gpu0_note": "frame gathered in GPU 0's planes over NVLink, then read back over GPU 0's PCIe link alone (16.6 MB per 1080p frame bounds it)" if tiles else None,
                "blocking_value": all_rays / e2e_sync_s / 1e6},
        "roofline": roofline,
    }
    if parity is not None:
        out["parity"] = parity
    if scaling_detail is not None:
        out["scaling_detail"] = scaling_detail
    if cpu is not None:
        out["cpu_baseline"] = cpu
        out["other_baselines"] = other_baselines()
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
