"""CPU check of the DEVICE source: svo_trace.cuh compiled for the host (tests/hostemu) against the oracle.

The build container has no GPU, so the `-m gpu` parity tests cannot run there.  These tests run the very statements
of the sm_100a kernels -- same headers, same template instances, same launch-time selection -- on the CPU and
demand the same bit-exact agreement with the oracle.  What they cannot see is nvcc's code generation and the
__global__ wrappers; `pytest -m gpu` on the B200 covers those through the C ABI."""
import numpy as np
import pytest

from hostemu import emu as E

PLANES = ("rgba8", "depth", "radiance", "hit_id", "iter", "primary_t")


def _same(g, w):
    if g.dtype.kind == "f":
        return (g.view(np.uint32) == w.view(np.uint32)) | (np.isnan(g) & np.isnan(w))
    return g == w


def _assert_planes_equal(got, want, what, planes=PLANES):
    for k in planes:
        same = _same(got[k], want[k])
        assert same.all(), "%s: plane %s differs in %d of %d elements" % (what, k, int((~same).sum()), same.size)


@pytest.fixture(scope="module")
def scene512(terrain512):
    s = E.Scene(terrain512)
    yield s
    s.close()


@pytest.fixture(scope="module")
def scene128(terrain128):
    s = E.Scene(terrain128)
    yield s
    s.close()


@pytest.mark.parametrize("cam", ["A", "B", "C"])
@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
def test_device_source_config1_all_modes(svo, oracle, terrain512, scene512, cam, mode):
    """BASELINE configs[0] (512^3, 640x360): Trav::run (tile kernel) and Trav::step (persistent / wavefront state
    machine), every plane, every render mode."""
    pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
    f = oracle.make_frame(pos, l1, l2, r1, r2, frame_number=3, render_mode=mode)
    want, _ = oracle.render(terrain512, f, 640, 360, nthreads=8)
    _assert_planes_equal(scene512.render(f, 640, 360, path=E.PATH_RUN), want, "run cam %s mode %d" % (cam, mode))
    if cam == "B":
        _assert_planes_equal(scene512.render(f, 640, 360, path=E.PATH_STEP), want, "step cam %s mode %d" % (cam, mode))
    # production instance (content box on, no validation planes): colour and depth unchanged
    if mode != 1:
        _assert_planes_equal(scene512.render(f, 640, 360, box=True, aux=False), want, "box cam %s mode %d" % (cam, mode),
                             planes=("rgba8", "depth"))


def test_device_source_stats_counters(svo, oracle, terrain512, scene512):
    """k_render_stats' counters (the algorithmic bytes of bench.py's roofline) equal the oracle's."""
    for cam, mode in (("A", 0), ("B", 2), ("C", 0)):
        pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
        f = oracle.make_frame(pos, l1, l2, r1, r2, frame_number=2, render_mode=mode)
        _, st = oracle.render(terrain512, f, 640, 360, nthreads=8, planes=("depth",))
        _, got = scene512.render(f, 640, 360, path=E.PATH_STATS)
        assert got == {"casts": st.casts, "iters": st.iters, "record_bytes": st.record_bytes}, (cam, mode)


def test_device_source_ray_stream(svo, oracle, terrain512, scene512):
    rng = np.random.default_rng(42)
    n = 100000
    rays = np.zeros(n, dtype=oracle.RAY_DTYPE)
    rays["o"] = rng.uniform(0.9, 2.1, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    rays["d"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["d"][:50, 0] = 0.0
    rays["d"][50:60] = 0.0
    rays["d"][60:70] = np.nan
    rays["d"][70:80, 1] = np.nan
    rays["o"][80:90] = np.nan
    for depth in (13, 9, 5):
        want, _ = oracle.cast_rays(terrain512, rays, max_depth=depth, nthreads=8)
        got = scene512.cast(rays, max_depth=depth)
        for k in ("id", "value", "iter"):
            assert np.array_equal(got[k], want[k]), (depth, k)
        assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))


def test_device_source_multichunk_rows_beam_accumulate(svo, oracle, terrain128, scene128):
    """Chunk splices + fill level, uneven row bands, the beam pre-pass feeding a frame, progressive accumulation."""
    W, H = 200, 120
    for cam in ("A", "B", "C"):
        pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
        f = oracle.make_frame(pos, l1, l2, r1, r2, frame_number=1, render_mode=0, max_depth=7)
        want, _ = oracle.render(terrain128, f, W, H, nthreads=8)
        got = None
        for y0, y1 in ((0, 37), (37, 38), (38, H)):
            part = scene128.render(f, W, H, y0, y1)
            got = part if got is None else {k: np.where((np.arange(H) >= y0)[(slice(None),) + (None,) * (part[k].ndim - 1)], part[k], got[k]) for k in part}
        _assert_planes_equal(got, want, "bands cam " + cam)
    pos, l1, l2, r1, r2 = svo.CAMERAS["B"]
    fb = oracle.make_frame(pos, l1, l2, r1, r2, frame_number=1, render_mode=2, use_beam=1, max_depth=7)
    beam_want = oracle.beam(terrain128, fb, W, H)
    beam_got = scene128.beam(fb, W, H)
    assert np.array_equal(beam_got.view(np.uint32), beam_want.view(np.uint32))
    want, _ = oracle.render(terrain128, fb, W, H, beam=beam_want, nthreads=8)
    _assert_planes_equal(scene128.render(fb, W, H, beam=beam_got), want, "beam frame")
    prev = None
    pos, l1, l2, r1, r2 = svo.CAMERAS["C"]
    for frame in (1, 2, 3):
        f = oracle.make_frame(pos, l1, l2, r1, r2, frame_number=frame, render_mode=0, max_depth=7, flags=1)
        want, _ = oracle.render(terrain128, f, 160, 90, nthreads=4, planes=("rgba8", "depth"), prev_rgba8=prev)
        got = scene128.render(f, 160, 90, box=True, aux=False, prev_rgba8=prev)
        assert np.array_equal(got["rgba8"], want["rgba8"]), frame
        prev = want["rgba8"]


@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_device_source_random_worlds_random_cameras(svo, oracle, seed):
    """The fuzz of test_gpu_parity.test_random_worlds_random_cameras on the host build of the device source."""
    rng = np.random.default_rng(seed)
    n = 32 if seed % 2 else 64
    vox = np.zeros((n, n, n), np.uint8)
    pts = rng.integers(0, n, size=(rng.integers(50, 600), 3))
    vox[pts[:, 2], pts[:, 1], pts[:, 0]] = rng.integers(1, 5, size=len(pts))
    for _ in range(3):
        lo = rng.integers(0, n - 10, 3)
        sz = rng.integers(3, 10, 3)
        vox[lo[2]:lo[2] + sz[2], lo[1]:lo[1] + sz[1], lo[0]:lo[0] + sz[0]] = rng.integers(1, 4)
    nodes, _ = oracle.build_dense(vox)
    W, H = 96, 64
    depth = int(np.log2(n))
    sc = E.Scene(nodes)
    for trial in range(6):
        pos = rng.uniform(0.7, 2.3, 3) if trial % 2 else rng.uniform(1.1, 1.9, 3)
        fwd = rng.normal(size=3)
        fwd /= np.linalg.norm(fwd)
        up = np.cross(fwd, rng.normal(size=3))
        up /= np.linalg.norm(up)
        right = np.cross(fwd, up)
        corners = [fwd + sx * 1.2 * right + sy * 0.8 * up for sx in (-1, 1) for sy in (-1, 1)]
        kw = dict(frame_number=int(rng.integers(1, 50)), render_mode=int(rng.choice([0, 0, 2, 2, 1, 3])),
                  max_depth=int(rng.integers(max(1, depth - 2), depth + 1)), casts=int(rng.integers(1, 4)),
                  cone_depth=int(rng.integers(1, depth + 1)), mirror_value=int(rng.choice([0, 4])))
        f = oracle.make_frame(pos, *corners, **kw)
        want, _ = oracle.render(nodes, f, W, H, nthreads=4)
        _assert_planes_equal(sc.render(f, W, H, path=E.PATH_RUN), want, "seed %d trial %d %s" % (seed, trial, kw))
        _assert_planes_equal(sc.render(f, W, H, path=E.PATH_STEP), want, "step seed %d trial %d %s" % (seed, trial, kw))
        if kw["render_mode"] != 1:
            _assert_planes_equal(sc.render(f, W, H, box=True, aux=False), want, "box seed %d trial %d %s" % (seed, trial, kw),
                                 planes=("rgba8", "depth"))
        # the __global__ kernels with warp-level protocols, on the SIMT emulator: lane refill (7, 8), octant binning (6)
        for kernel in (7, 8, 6, 9, 10, 13, 15, 16):
            _assert_planes_equal(sc.launch_render(f, W, H, kernel=kernel, aux=True), want, "kernel %d seed %d trial %d %s" % (kernel, seed, trial, kw))
            _assert_planes_equal(sc.launch_render(f, W, H, kernel=kernel, aux=False, box=True), want,
                                 "kernel %d box seed %d trial %d %s" % (kernel, seed, trial, kw), planes=("rgba8", "depth"))
    sc.close()


def test_device_source_iteration_cap_boundary(oracle):
    """Casts ending exactly at, just below and beyond the 1500-iteration cap: the device tests the cap on the POP
    path only (Trav), so the boundary is its own case."""
    import svo_stream as S
    nodes = S.serialise(S.tube(10))
    rays = S.tube_rays(10, 100000)
    want, _ = oracle.cast_rays(nodes, rays, 13, nthreads=4)
    assert (want["iter"] == 1500).sum() > 0 and (want["iter"] == 1501).sum() > 0
    sc = E.Scene(nodes)
    got = sc.cast(rays, 13)
    for k in ("id", "iter", "value"):
        assert np.array_equal(got[k], want[k]), k
    got = sc.launch_cast(rays[:6000], 13, kernel=1)  # persistent stream kernel: Trav::step + cap_fixup at every exit
    for k in ("id", "iter", "value"):
        assert np.array_equal(got[k], want[k][:6000]), k
    assert (want["iter"][:6000] == 1500).any() and (want["iter"][:6000] == 1501).any()
    W, H = 96, 64
    h = 2.0 ** -10
    cam = ((1.9, 1 + h / 2, 1 + h / 2), (-1, -1e-4, -1e-4), (-1, 1e-4, -1e-4), (-1, -1e-4, 1e-4), (-1, 1e-4, 1e-4))
    seen = set()
    for x0 in (1.9, 1.732, 1.7305, 1.73):
        pos = (x0,) + cam[0][1:]
        for mode in (1, 0):
            f = oracle.make_frame(pos, *cam[1:], frame_number=3, render_mode=mode, max_depth=13)
            wantp, _ = oracle.render(nodes, f, W, H, nthreads=4)
            _assert_planes_equal(sc.render(f, W, H, path=E.PATH_RUN), wantp, "cap run %s mode %d" % (x0, mode))
            _assert_planes_equal(sc.render(f, W, H, path=E.PATH_STEP), wantp, "cap step %s mode %d" % (x0, mode))
            _assert_planes_equal(sc.launch_render(f, W, H, kernel=7, aux=True), wantp, "cap refill kernel %s mode %d" % (x0, mode))
            _assert_planes_equal(sc.launch_render(f, W, H, kernel=13, aux=True), wantp, "cap balanced kernel %s mode %d" % (x0, mode))
            seen |= set(np.unique(wantp["iter"]).tolist())
    assert {1497, 1500, 1501} <= seen
    sc.close()


def test_device_math_equals_oracle_math(oracle):
    rng = np.random.default_rng(7)
    L = oracle.lib()
    cases = {0: np.concatenate([rng.uniform(-7, 7, 3000), rng.uniform(-2e6, 2e6, 3000), [0.0, -0.0, np.inf, np.nan, 1e9, 3e38]]),
             1: np.concatenate([rng.uniform(-7, 7, 3000), rng.uniform(-2e6, 2e6, 3000), [0.0, -0.0, np.inf, np.nan]]),
             2: np.concatenate([rng.uniform(-1, 1, 5000), [-1.0, 1.0, 0.5, -0.5, 1.0000001, np.nan]]),
             3: np.concatenate([rng.uniform(-20, 5, 5000), rng.uniform(-110, 90, 1000), [0.0, np.nan, -np.inf, np.inf]])}
    names = {0: "sin", 1: "cos", 2: "acos", 3: "exp"}
    for fn, xs in cases.items():
        f = getattr(L, "svo_oracle_" + names[fn])
        for v in xs.astype(np.float32):
            a, b = np.float32(E.math(fn, v)), np.float32(f(float(v)))
            assert a.view(np.uint32) == b.view(np.uint32) or (np.isnan(a) and np.isnan(b)), (names[fn], v)


def test_simt_model_invariants(svo, oracle, terrain128, scene128):
    """The divergence model of tools/simt_model.py: its counters equal the oracle's, and the bounds order as they must
    (32 busy lanes <= slowest lane per warp <= every real loop organisation)."""
    W, H = 200, 120
    pos, l1, l2, r1, r2 = svo.CAMERAS["C"]
    f = oracle.make_frame(pos, l1, l2, r1, r2, frame_number=1, render_mode=0, max_depth=7)
    costs = [14, 40, 18, 34, 2, 8, 10, 700]
    r = scene128.simt(f, W, H, costs, box=False)
    _, st = oracle.render(terrain128, f, W, H, nthreads=4, planes=("depth",))
    # every iteration of every cast is replayed, none invented; NaN rays (zero-normal bounces) leave before the loop on
    # the device where the reference spins 1500 iterations
    assert r["casts"] == st.casts and r["iters"] + 1500 * r["early"] == st.iters
    assert r["pushes"] + r["advances"] + r["pops"] + r["casts"] - r["early"] == r["iters"]  # one exit iteration per cast
    assert r["ideal"] <= r["longest_lane"] <= min(r[k] for k in ("if_if", "while_while", "ww_1_1", "ww_inf_1", "ww_1_inf", "ww_4_2"))
    for G in (4, 16, 64):
        assert r["regroup%d_by_length" % G] <= r["regroup%d_as_is" % G] * 1.0001
    rb = scene128.simt(f, W, H, costs, box=True)
    assert rb["casts"] == r["casts"] and rb["iters"] <= r["iters"]  # the content box only ever removes iterations
    # the warp tile the kernel uses (8x4) is what the model replays by default
    assert scene128.simt(f, W, H, costs, box=False, tile_w=8)["if_if"] == r["if_if"]


KERNEL_IDS = {0: "tile", 5: "tile64", 1: "persistent", 2: "wavefront", 6: "binned", 4: "smem", 7: "refill4", 8: "refill2", 9: "smemstack", 10: "widestack",
              11: "regs72", 13: "balanced", 14: "wide_bands", 15: "split", 16: "split_presetup", 17: "tile_queue", 19: "balanced_mask7"}


@pytest.mark.parametrize("kernel", list(KERNEL_IDS), ids=list(KERNEL_IDS.values()))
def test_global_kernels_on_simt_emulator(svo, oracle, terrain128, scene128, kernel):
    """The __global__ functions of svo_kernels.cu and the product's launch_render() (variant dispatch, grid arithmetic,
    band interleaving) on the coroutine SIMT emulator: warp ballots / shuffles / atomics of the persistent kernel, the
    shared-memory counting sort and block-wide votes of the binned kernel, the shared-memory descriptor prefix --
    every plane bit-exact against the oracle, image sizes that do not divide into CTA tiles, uneven row bands."""
    W, H = 200, 120
    for cam, mode in (("B", 0), ("C", 2), ("A", 0)) if kernel not in (1, 2) else (("B", 0), ("C", 2)):
        pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
        f = oracle.make_frame(pos, l1, l2, r1, r2, frame_number=2, render_mode=mode, max_depth=7, casts=3 if cam == "A" else 2)
        want, _ = oracle.render(terrain128, f, W, H, nthreads=8)
        if kernel != 4:  # variant 4 exists only as the production instance (no validation planes)
            got = None
            for y0, y1 in ((0, 37), (37, 38), (38, H)):
                got = scene128.launch_render(f, W, H, y0, y1, kernel=kernel, aux=True, into=got)
            _assert_planes_equal(got, want, "kernel %d cam %s aux" % (kernel, cam))
        got = scene128.launch_render(f, W, H, kernel=kernel, aux=False, box=True)
        _assert_planes_equal(got, want, "kernel %d cam %s production" % (kernel, cam), planes=("rgba8", "depth"))
    if kernel == 17:  # every launch bumped both fence words once, from its last CTA
        assert E.lib().emu_fence_word(0) == E.lib().emu_fence_word(1) > 0
    if kernel in (0, 5, 14, 17):  # interleaved bands of the multi-GPU tile partition, one launch per part
        pos, l1, l2, r1, r2 = svo.CAMERAS["B"]
        f = oracle.make_frame(pos, l1, l2, r1, r2, frame_number=2, render_mode=0, max_depth=7)
        want, _ = oracle.render(terrain128, f, W, H, nthreads=8)
        for parts, rows in ((3, 8), (2, 16)):
            got = None
            for part in range(parts):
                got = scene128.launch_render(f, W, H, kernel=kernel, aux=True, band_stride=parts, band_offset=part, band_rows=rows, into=got)
            _assert_planes_equal(got, want, "kernel %d interleaved %d x %d rows" % (kernel, parts, rows))


@pytest.mark.parametrize("kernel", [0, 1, 2], ids=["gridstride", "persistent", "gridstride_wide"])
def test_ray_stream_kernels_on_simt_emulator(oracle, terrain128, scene128, kernel):
    """k_cast_stream and k_cast_stream_persistent (warp-level ray fetch from a global counter, lane refill) through the
    product's launch_cast: every hit record in the caller's slot, with and without a binned order, stream lengths that are
    not multiples of the warp / chunk size, zero / axis-parallel / NaN rays that end before the loop."""
    rng = np.random.default_rng(3)
    for n in (1, 31, 700, 5000):
        rays = np.zeros(n, dtype=oracle.RAY_DTYPE)
        rays["o"] = rng.uniform(0.9, 2.1, (n, 3)).astype(np.float32)
        d = rng.normal(size=(n, 3))
        rays["d"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        rays["d"][::37, 0] = 0.0
        rays["d"][5::101] = 0.0
        rays["d"][7::113] = np.nan
        want, _ = oracle.cast_rays(terrain128, rays, max_depth=7, nthreads=4)
        for order in (None, rng.permutation(n)):
            got = scene128.launch_cast(rays, max_depth=7, kernel=kernel, order=order)
            for k in ("id", "value", "iter"):
                assert np.array_equal(got[k], want[k]), (kernel, n, k, order is not None)
            assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32)), (kernel, n)


def test_simt_model_ray_stream(oracle, terrain128, scene128):
    """The stream front end of the divergence model: counters equal the oracle's, bounds order as they must."""
    rng = np.random.default_rng(8)
    n = 3000
    rays = np.zeros(n, dtype=oracle.RAY_DTYPE)
    rays["o"] = rng.uniform(1.05, 1.95, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    rays["d"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    _, st = oracle.cast_rays(terrain128, rays, max_depth=7, nthreads=4)
    r = scene128.simt_stream(rays, [14, 40, 18, 31, 2, 8, 10, 300, 230, 6.5], max_depth=7)
    assert r["casts"] == n and r["iters"] == st.iters
    assert r["ideal"] <= r["longest_lane"] <= r["grid_stride"]


def test_simt_emulator_collective_semantics():
    """The emulator itself: collectives complete among the threads that have not exited, exited lanes vote 0, shuffles /
    reductions / block-wide votes / shared memory / atomics behave like sm_70+ for full-mask calls."""
    out = E.selftest(blocks=3, os_threads=2)
    for b in range(3):
        for t in range(128):
            o = out[b, t]
            lane, warp = t & 31, t >> 5
            if t % 3 == 2:
                assert o[0] == 0xDEAD and not o[1:].any()
                continue
            alive = [l for l in range(32) if (warp * 32 + l) % 3 != 2]
            assert o[0] == sum(1 << l for l in alive if l & 1)
            assert o[1] == (warp * 32 + alive[0]) * 7 + 1
            assert o[2] == sum(alive)
            assert o[3] == 1
            assert o[4] == len(alive)          # shared-memory counter: one atomicAdd per live thread of the warp
            assert o[5] == max(l % 5 for l in alive)  # loop with a vote per trip: everyone stays until the slowest lane is done


@pytest.mark.parametrize("kernel", [15, 16], ids=["split", "split_presetup"])
def test_split_kernels_on_simt_emulator(svo, oracle, terrain128, scene128, kernel):
    """Kernel variant 15: k_split_primary queues the last cast of every mode-0 path, k_split_bounce traces the queue with lane
    refill and finishes the pixels from 80-byte records.  Colour and depth bit-exact for one and several bounces, the mirror
    rule, row bands, image sizes that do not divide into tiles; frames it does not cover fall back to the default kernel."""
    W, H = 200, 120
    for cam, casts, mirror in (("B", 2, 0), ("C", 2, 0), ("A", 3, 0), ("C", 4, 1), ("B", 2, 3)):
        pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
        f = oracle.make_frame(pos, l1, l2, r1, r2, frame_number=9, render_mode=0, max_depth=7, casts=casts, mirror_value=mirror)
        want, _ = oracle.render(terrain128, f, W, H, nthreads=8, planes=("rgba8", "depth"))
        for box in (True, False):
            got = scene128.launch_render(f, W, H, kernel=kernel, aux=False, box=box)
            assert scene128.last_launches == 2
            _assert_planes_equal(got, want, "split cam %s casts %d mirror %d box %s" % (cam, casts, mirror, box), planes=("rgba8", "depth"))
        got = None
        for y0, y1 in ((0, 37), (37, 38), (38, H)):
            got = scene128.launch_render(f, W, H, y0, y1, kernel=kernel, aux=False, box=True, into=got)
        _assert_planes_equal(got, want, "split bands cam %s" % cam, planes=("rgba8", "depth"))
    pos, l1, l2, r1, r2 = svo.CAMERAS["B"]
    for kw in (dict(render_mode=2), dict(render_mode=0, casts=1)):  # not covered: one launch, the default kernel
        f = oracle.make_frame(pos, l1, l2, r1, r2, frame_number=9, max_depth=7, **kw)
        want, _ = oracle.render(terrain128, f, W, H, nthreads=8, planes=("rgba8", "depth"))
        got = scene128.launch_render(f, W, H, kernel=kernel, aux=False, box=True)
        assert scene128.last_launches == 1
        _assert_planes_equal(got, want, "split fallback %s" % kw, planes=("rgba8", "depth"))
    # progressive accumulation reads the previous frame's colour at the pixel: through the rebuilt Pixel too
    prev = None
    pos, l1, l2, r1, r2 = svo.CAMERAS["C"]
    for frame in (1, 2, 3):
        f = oracle.make_frame(pos, l1, l2, r1, r2, frame_number=frame, render_mode=0, max_depth=7, flags=1)
        want, _ = oracle.render(terrain128, f, 160, 90, nthreads=4, planes=("rgba8", "depth"), prev_rgba8=prev)
        got = scene128.launch_render(f, 160, 90, kernel=kernel, aux=False, box=True, prev_rgba8=prev)
        assert np.array_equal(got["rgba8"], want["rgba8"]), frame
        prev = want["rgba8"]


@pytest.mark.parametrize("cam", ["A", "B", "C"])
def test_conservative_beam_prepass_on_simt_emulator(svo, oracle, terrain128, scene128, cam):
    """svo_beam_conservative (k_beam_lattice + k_beam_minfilter): every block's value is a lower bound on the primary hit
    distance of its 16 pixels (+inf only where all 16 miss), it is not vacuous, and a frame whose primary casts start
    there (flags bit 1) is bit-identical in colour and depth to the frame without it while running fewer iterations."""
    W, H = 160, 96
    pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
    kw = dict(frame_number=3, render_mode=0, max_depth=7)
    f0 = oracle.make_frame(pos, l1, l2, r1, r2, **kw)
    want, _ = oracle.render(terrain128, f0, W, H, nthreads=8)
    beam = scene128.beam_conservative(f0, W, H)
    t = np.where(want["hit_id"] != 0xFFFFFFFF, want["primary_t"], np.inf).astype(np.float32)
    tmin = t.reshape(H // 4, 4, W // 4, 4).min(axis=(1, 3))
    assert (beam <= tmin).all(), "not conservative in %d blocks" % int((beam > tmin).sum())
    hit_blocks = np.isfinite(tmin)
    if hit_blocks.any():
        assert (beam[hit_blocks] > 0).mean() > 0.5 and np.median(beam[hit_blocks] / tmin[hit_blocks]) > 0.3  # a useful bound even at this tiny resolution (the margins are in lattice spacings)
    for parts in (1, 2, 3, 8, 40):  # the tile partition's shared pre-pass: lattice rows traced part by part (40 > rows: empty parts still signal)
        assert np.array_equal(scene128.beam_in_parts(f0, W, H, parts).view(np.uint32), beam.view(np.uint32)), parts
    f1 = oracle.make_frame(pos, l1, l2, r1, r2, flags=2, **kw)
    for kernel in (13, 17, 0):
        got = scene128.launch_render(f1, W, H, kernel=kernel, aux=False, box=True, beam=beam)
        _assert_planes_equal(got, want, "beam floor kernel %d cam %s" % (kernel, cam), planes=("rgba8", "depth"))
    # validation instance (aux planes): the floor is ignored, iteration counts are the reference's
    got = scene128.launch_render(f1, W, H, kernel=13, aux=True, beam=beam)
    _assert_planes_equal(got, want, "beam floor ignored with aux planes")
    # mode 2 (the engine's default: primary + shadow ray)
    kw2 = dict(frame_number=3, render_mode=2, max_depth=7)
    want2, _ = oracle.render(terrain128, oracle.make_frame(pos, l1, l2, r1, r2, **kw2), W, H, nthreads=8, planes=("rgba8", "depth"))
    got2 = scene128.launch_render(oracle.make_frame(pos, l1, l2, r1, r2, flags=2, **kw2), W, H, kernel=13, aux=False, box=True, beam=beam)
    _assert_planes_equal(got2, want2, "beam floor mode 2", planes=("rgba8", "depth"))
