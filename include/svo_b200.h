/*
 * svo_b200.h -- C ABI of libsvo_b200.so: the B200 (sm_100a) replacement for
 * the reference renderer's GPU dispatch boundary for ONE path, the sparse
 * voxel octree trace (src/shaders/svotrace.comp) and its beam pre-pass
 * (src/shaders/svobeam.comp).
 *
 * Every entry point names the reference interface it replaces (paths relative
 * to the reference repo root).  Plain pointers and sizes only; no C++ or
 * torch types.  All functions return 0 on success and a non-zero CUresult-
 * style code otherwise (never abort the host VM); svo_last_error() gives the
 * text.  The library is called from one host thread at a time, like the
 * reference's GL context (Main.java runs everything on the GLFW thread).
 *
 * There is no CPU fallback: every call fails with SVO_ERR_NO_DEVICE when no
 * CUDA device is usable.
 */
#ifndef SVO_B200_H
#define SVO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVO_ABI_VERSION 1
#define SVO_NO_HIT 0xFFFFFFFFu
#define SVO_FRAME_ACCUMULATE 1
#define SVO_FRAME_BEAM_FLOOR 2

enum svo_status {
  SVO_OK = 0,
  SVO_ERR_INVALID = 1,    /* bad argument (mirrors CUDA_ERROR_INVALID_VALUE) */
  SVO_ERR_OOM = 2,        /* CUDA_ERROR_OUT_OF_MEMORY */
  SVO_ERR_NO_DEVICE = 100,/* CUDA_ERROR_NO_DEVICE: no fallback exists */
  SVO_ERR_CUDA = 999,     /* any other CUDA failure; text in svo_last_error */
  SVO_ERR_FORMAT = 1000,  /* node stream is not a tree this path can walk */
  SVO_ERR_NO_SCENE = 1001,/* render/cast before svo_upload */
  SVO_ERR_FENCE = 1002    /* svo_sync: a svo_fence_wait gave up (~2 s watchdog): a peer GPU fell behind or died */
};

typedef struct svo_ctx svo_ctx;

/* Per-frame parameters.  Replaces the uniforms written at
 * src/engine/Main.java:269-283 (locations 8 camPos; 1-4 l1,l2,r1,r2; 5
 * frameNumber; 6 renderMode; 11 useBeamOptimization; location 9 bufferEnd
 * is the upload length and lives in svo_upload).  The last four members
 * expose what the shader hard-codes: MAX_DEPTH 13 (svotrace.comp:40), the
 * bounce-loop count 2 (:444), the cone LOD cut 11 (:275-277) and the
 * commented-out mirror material (:500-504; 0 = off = as shipped). */
typedef struct svo_frame {
  float camPos[3];
  float l1[3], l2[3], r1[3], r2[3];
  int32_t frameNumber;
  int32_t renderMode;
  int32_t useBeam;
  int32_t maxDepth;
  int32_t casts;
  int32_t coneDepth;
  int32_t mirrorValue;
  int32_t flags; /* bit 1 (SVO_FRAME_BEAM_FLOOR): the beam plane holds svo_beam_conservative's lower bounds; primary casts start there.
                  * bit 0 (SVO_FRAME_ACCUMULATE): progressive running mean over frameNumber, the block the shader
                  * has commented out at svotrace.comp:712-719; other bits must be 0 */
} svo_frame;

/* ray-stream records (new capability, BASELINE.json configs[3]) */
typedef struct svo_ray { float o[3]; float d[3]; } svo_ray;
typedef struct svo_hit { uint32_t id; float t; uint32_t value; uint32_t iter; } svo_hit;

/* output planes, all width*height elements, row-major, row 0 = gl row 0 */
enum svo_plane {
  SVO_PLANE_COLOR_RGBA8 = 0, /* framebufferImage, image unit 0 (Main.java:66-70) */
  SVO_PLANE_DEPTH = 1,       /* depthbufferImage, image unit 1 (Main.java:73-77) */
  SVO_PLANE_BEAM = 2,        /* beambufferImage (W/4 x H/4), image unit 2 (Main.java:79-86) */
  SVO_PLANE_HIT_ID = 3,      /* new: primary res.pointer (svotrace.comp:294,381; store commented out at :728) */
  SVO_PLANE_ITER = 4,        /* new: primary loop iterations (render mode 1 shows it as a heat map, :428) */
  SVO_PLANE_PRIMARY_T = 5,   /* new: primary res.t */
  SVO_PLANE_RADIANCE = 6,    /* new: finalcolor before the rgba8 store, float4 */
  SVO_PLANE_BEAM_LATTICE = 7,/* svo_device_ptr / svo_ipc_export only: the (W/4+1) x (H/4+1) lattice buffer of the conservative beam
                              * pre-pass of lane `plane >> 8` (shared between the GPUs of the tile partition, svo_beam_lattice_rows) */
  SVO_PLANE_BACK = 0x100     /* OR-ed to COLOR_RGBA8 / DEPTH in svo_device_ptr and svo_ipc_export: the second set
                              * (svo_swap_buffers, lane 1); without it they name the first set.  In general
                              * plane | (lane << 8) names the set of lane 1..6 (svo_select_lane) */
};

enum svo_option {
  SVO_OPT_AUX_PLANES = 1,     /* 0/1: also write planes 3..6 (validation outputs); default 0 */
  SVO_OPT_FAST_MATH = 2,      /* 0: separately rounded arithmetic = the --fmad=false validation semantics (default, the bit-exact
                               * contract); 1: t arithmetic contracted into FFMA (not bit-exact, measured slower) */
  SVO_OPT_KERNEL = 3,         /* kernel variant.  10 (default, fastest): tile kernel, one thread per pixel, 16-byte
                               * stack entries that carry the parent's descriptor; 0: the same with 8-byte entries (also what
                               * svo_render_interleaved and SVO_OPT_FAST_MATH run); 1 persistent megakernel, 2 wavefront, 4 tile +
                               * shared-memory upper levels, 5 64-thread CTAs, 6 tile + per-CTA octant binning, 7 / 8 tile +
                               * warp-local lane refill for bounce rays (4 / 2 pixels per thread), 9 parent stack in shared
                               * memory, 11 / 12 72 / 80 registers per thread, 13 variant 10 with the loop's integer work moved to
                               * the FMA pipe, 14 variant 10 in the band-interleaved launch too, 15 the last cast of
                               * mode-0 pixels in a persistent kernel of its own with lane refill, 16 the same with the cast's set-up
                               * done by the primary kernel, 17 variant 10 as persistent warps that take whole 8x4 tiles from a queue (what
                               * svo_render_interleaved runs unless 0 or 14 is selected).  All bit-exact; the others are measured ablations */
  SVO_OPT_L2_PERSIST = 4,     /* 0/1: L2 access-policy window (persisting) over the upper octree levels, applied at the next
                               * upload; default 0 (measured: no effect, the path is not memory bound) */
  SVO_OPT_RAY_SORT = 5,       /* trace ray streams of >= 65536 rays in binned order (results stay in the caller's order): 0 off; 1 = (direction
                               * octant, origin Morton code); 2 = (octant, 64^3 origin cell, direction bin, finer origin bits); default 2 */
  SVO_OPT_CONTENT_BOUNDS = 6, /* 0/1: end casts that cannot hit anything once they are outside the bounding box of the octree's
                               * non-empty leaves (computed at upload).  Outputs are unchanged; only the iteration count of
                               * MISSING casts differs, so it is ignored in render mode 1 and with SVO_OPT_AUX_PLANES.  default 1 */
  SVO_OPT_BAND_ROWS = 7,      /* image rows per band of svo_render_interleaved (multiple of 8); default 8 */
  SVO_OPT_GPU_TRANSCODE = 8,  /* 0/1: build the traversal descriptors from the uploaded stream on the device (default 1) or on
                               * the host; same result bit for bit */
  SVO_OPT_STREAM_KERNEL = 9   /* svo_cast / svo_cast_device: 0 (default) one grid-stride thread per ray; 1 persistent threads with
                               * warp-level ray fetch -- lanes whose ray has finished are re-armed from the stream; 2 the grid-stride kernel with
                               * 16-byte stack entries (1, 2 not yet measured) */
};

/* -- lifetime: replaces Main.preRun's image/shader setup (Main.java:62-109) and
 *    Renderer.addShader (Renderer.java:43-54) ------------------------------- */
int svo_abi_version(void);
int svo_device_count(int *count);
int svo_create(svo_ctx **out, int device, int width, int height);
void svo_destroy(svo_ctx *ctx);
const char *svo_last_error(const svo_ctx *ctx); /* ctx may be NULL (creation errors) */
int svo_set_option(svo_ctx *ctx, int option, int64_t value);
int svo_get_option(const svo_ctx *ctx, int option, int64_t *value);
/* Run on a caller-owned CUDA stream (cudaStream_t / CUstream as void*); NULL
 * restores the context's own stream. */
int svo_set_stream(svo_ctx *ctx, void *cuda_stream);

/* -- octree upload: replaces Renderer.addSSBO(7, ByteBuffer)
 *    (Renderer.java:123-129, Main.java:122).  `nodes` is the engine's node
 *    buffer (Octree.java:63-67), valid bytes [0, nbytes) with nbytes =
 *    Octree.memOffset.  The library copies; the caller keeps ownership. */
int svo_upload(svo_ctx *ctx, const uint8_t *nodes, uint64_t nbytes);
/* replaces Renderer.updateSSBO(7, buf, start, end) (Renderer.java:136-146,
 * Main.java:349-350): `nodes` is the base of the whole buffer, bytes
 * [start,end) changed.  end may exceed the previous length (appended nodes). */
int svo_upload_range(svo_ctx *ctx, const uint8_t *nodes, uint64_t start, uint64_t end);
/* How the last svo_upload_range was absorbed: out[0] descriptors whose record or child block really changed, [1] subtrees
 * re-walked, [2] descriptors appended, [3] 1 if it fell back to a whole transcode (edit too wide, arrays full, or a quarter
 * of the array already patched-in).  The range is stored and compared with the old bytes in one pass; only the subtrees
 * the changed bytes belong to are transcoded again (csrc/svo_gpu_transcode.cu, gpu_patch).  All zero: nothing changed. */
int svo_upload_stats(const svo_ctx *ctx, uint64_t out[4]);
/* Layout-independent fingerprint of the scene's descriptor tree (depth-first from the root): out[0] reachable descriptors,
 * [1] hash of their masks / reference offsets / depths, [2] deepest level, [3] descriptors stored.  Equal for a scene
 * patched by svo_upload_range and the same stream uploaded whole. */
int svo_scene_canonical(svo_ctx *ctx, uint64_t out[4]);
/* World generation ON THE DEVICE: the stream svo_build_terrain writes, built by kernels (min/max pyramids, level-
 * synchronous classify / size / offset / emit sweeps; csrc/svo_gpu_build.cu) straight into HBM and made the
 * context's scene as svo_upload would -- the node stream never crosses PCIe.  Replaces, for heightmap worlds,
 * Octree.constructCompleteOctree (Octree.java:192-353) + constructInnerOctree (:511-608) + genSurfaceNormal
 * (:620-649) + checkBigNodeExposed (:651-670) + chunkgen-heightmap.comp:13-31 and the addSSBO that follows
 * (Main.java:122).  height / mat: n x n host arrays (row = z).  *out_bytes = Octree.memOffset. */
int svo_build_terrain_device(svo_ctx *ctx, const uint16_t *height, const uint8_t *mat, int n, int chunk, uint64_t *out_bytes);
/* The current scene's node stream back to the host (Octree.writeBufferToFile's payload, Octree.java:974-993);
 * dst_cap >= info[0] of svo_scene_info. */
int svo_download(svo_ctx *ctx, uint8_t *dst, uint64_t dst_cap);
/* counts after upload: info[0] node-stream bytes, [1] interior descriptors,
 * [2] octree levels, [3] device bytes used by the scene */
int svo_scene_info(const svo_ctx *ctx, uint64_t info[4]);

/* -- dispatch: replaces Renderer.dispatchCompute(traceShader, 240, 135, 1) +
 *    glMemoryBarrier (Renderer.java:118-121, Main.java:285).  Asynchronous. */
int svo_render(svo_ctx *ctx, const svo_frame *frame);
/* rows [y0,y1) only: the image-tile partition for multi-GPU */
int svo_render_rows(svo_ctx *ctx, const svo_frame *frame, int y0, int y1);
/* Bands part, part+parts, part+2*parts, ... of the frame (SVO_OPT_BAND_ROWS image rows each, default 8) in ONE
 * launch: the interleaved image partition of the multi-GPU mode (rank = part, world size = parts). */
int svo_render_interleaved(svo_ctx *ctx, const svo_frame *frame, int part, int parts);
/* The same launch followed, in stream order, by svo_fence_signal(fence_ptrs, n, slot) (n = 0: this context's own
 * fence; n = -1: no signal).  With the tile-queue kernel (SVO_OPT_KERNEL 17, what the interleaved launch runs by
 * default) the signal is part of the render kernel -- its last CTA bumps the fences -- so a rank's whole share of a
 * frame, stores into the owner's planes over NVLink and frame-complete fence included, is ONE kernel launch. */
int svo_render_interleaved_signal(svo_ctx *ctx, const svo_frame *frame, int part, int parts, void *const *fence_ptrs, int n, int slot);
/* replaces dispatchCompute(beamShader, W/8/4, H/8/4, 1) (Main.java:257-266) */
int svo_beam(svo_ctx *ctx, const svo_frame *frame);
/* The beam optimisation done conservatively (what svobeam.comp set out to do; its own version is neither a lower bound
 * nor in the fine pass's units): fills the beam plane with, per 4x4 pixel block, a PROVEN lower bound on the
 * primary hit distance of its 16 pixels (+inf: all 16 miss) from normalised rays through the block corners with an LOD
 * stop at 2*sqrt(2) lattice spacings, a min filter over +-4 lattice steps and the node-diagonal margin
 * (csrc/svo_kernels.cu, k_beam_lattice / k_beam_minfilter).  A frame rendered with flags | SVO_FRAME_BEAM_FLOOR (and
 * useBeam = 0) starts every primary cast's walk there: colour, depth and hit are unchanged bit for bit, only loop
 * iterations are saved.  It applies to render modes 0 and 3; it is ignored where the primary cast's iteration count is
 * observable: mode 1 (heat map), mode 2 (the penumbra term reads the primary's stale count when the shadow ray misses,
 * svotrace.comp:616-619) and with SVO_OPT_AUX_PLANES. */
int svo_beam_conservative(svo_ctx *ctx, const svo_frame *frame);
/* The same pre-pass in two halves, for the multi-GPU tile partition, where every rank needs the whole beam plane but should not
 * trace the whole lattice: svo_beam_lattice_rows traces lattice rows [row0, row1) only and stores them into the ndst lattice
 * buffers of dst_ptrs (its own and, over NVLink, its peers': svo_device_ptr / svo_ipc_import of SVO_PLANE_BEAM_LATTICE |
 * lane << 8; ndst = 0: its own), and the launch's last CTA bumps slot `slot` of the nsig fences; after svo_fence_wait for all
 * ranks' rows, svo_beam_filter turns the current lane's lattice into its beam plane (min filter + margin). */
int svo_beam_lattice_rows(svo_ctx *ctx, const svo_frame *frame, int row0, int row1, void *const *dst_ptrs, int ndst,
                          void *const *fence_ptrs, int nsig, int slot);
int svo_beam_filter(svo_ctx *ctx);
int svo_sync(svo_ctx *ctx);
/* Seven lanes -- a CUDA stream and a colour/depth plane set each (set 1 = the SVO_PLANE_BACK set).  svo_select_lane makes
 * `lane` (0..6) current: later calls enqueue on its stream, svo_render draws into its set, reads take it from there.
 * Work on different lanes may overlap on the GPU: rendering frame k+1 on the other lane lets its first tiles fill the SMs
 * that frame k's last, longest tiles leave idle (measured: a 1080p frame carries ~0.13 ms of such tail -- the critical
 * path of its longest rays -- whatever share of the frame a GPU renders, so the 8-GPU tile partition keeps several frames in
 * flight).  Frames on one lane stay ordered; svo_sync, svo_timer_*, uploads, the beam passes and svo_cast order both lanes.  svo_swap_buffers is
 * svo_select_lane(other) + the wait for that set's last read-back.  With a caller-owned stream (svo_set_stream), with
 * SVO_OPT_AUX_PLANES or with a kernel variant whose workspace exists once (1, 2, 15, 16) both lanes share one stream. */
int svo_select_lane(svo_ctx *ctx, int lane);

/* -- readback: replaces glGetTexImage(depth) every frame (Main.java:132-146).
 *    dst is host memory (pinned or pageable), width*height elements. */
int svo_read_plane(svo_ctx *ctx, int plane, void *dst, uint64_t dst_bytes);
int svo_read_plane_rows(svo_ctx *ctx, int plane, int y0, int y1, void *dst, uint64_t dst_bytes);
int svo_read_color_rgba8(svo_ctx *ctx, uint8_t *dst);
int svo_read_depth(svo_ctx *ctx, float *dst);
int svo_read_depth_at(svo_ctx *ctx, int x, int y, float *dst); /* the crosshair pick, Main.java:144-146 */
int svo_read_hit_id(svo_ctx *ctx, uint32_t *dst);
int svo_read_iter(svo_ctx *ctx, uint32_t *dst);
int svo_read_primary_t(svo_ctx *ctx, float *dst);
int svo_read_radiance_f32(svo_ctx *ctx, float *dst);
/* Pipelined read-back (the engine's glGetTexImage every frame, Main.java:132-146, without stalling the GPU):
 * the context keeps two colour/depth sets.  svo_read_planes_async enqueues, on a copy stream and after everything
 * rendered so far, the device->host copies of the CURRENT set into (pinned) host memory; svo_swap_buffers makes the
 * other set the render target (the next render waits, on the device, for that set's last copy); svo_read_wait
 * blocks the host until all enqueued copies have landed.  Per frame: svo_render; svo_read_planes_async;
 * svo_swap_buffers -- frame s+1 renders while frame s crosses PCIe. */
int svo_read_planes_async(svo_ctx *ctx, uint8_t *rgba8_dst, float *depth_dst);
/* The multi-GPU variant of svo_read_planes_async: copies the rows of bands part, part + parts, ... (what
 * svo_render_interleaved(part, parts) drew into this context's current set) into FULL-FRAME host buffers, each row at its
 * place in the frame -- one strided copy per plane.  With the frame buffers in host memory shared by the ranks (pinned by
 * each), every GPU delivers its bands over its own PCIe link and the frame assembles in host memory: the read-back of
 * an N-GPU frame is not bounded by GPU 0's link (16.6 MB per 1080p frame).  Completion: svo_read_wait. */
int svo_read_interleaved_async(svo_ctx *ctx, int part, int parts, uint8_t *rgba8_frame, float *depth_frame);
int svo_swap_buffers(svo_ctx *ctx);
int svo_read_wait(svo_ctx *ctx);
/* device address of a plane (for CUDA/NCCL/peer interop); NULL if absent */
void *svo_device_ptr(svo_ctx *ctx, int plane);
/* redirect a plane to caller-owned device memory (e.g. a peer GPU's frame
 * buffer mapped over NVLink); NULL restores the context's own allocation */
int svo_bind_plane(svo_ctx *ctx, int plane, void *device_ptr);

/* Multi-GPU frame buffers over NVLink (one process per GPU).  svo_ipc_export fills a 72-byte
 * handle (cudaIpcMemHandle_t of the memory block + the plane's offset inside it) for one of this
 * context's own planes; the owning process sends it to its peers by any means; svo_ipc_import maps
 * it in the peer process and returns a device pointer that svo_bind_plane accepts, so a peer's
 * kernel stores its image tiles straight into the owner's plane (no gather step).  svo_ipc_close
 * unmaps everything the context imported. */
#define SVO_IPC_HANDLE_BYTES 72
int svo_ipc_export(svo_ctx *ctx, int plane, uint8_t handle[SVO_IPC_HANDLE_BYTES]);
int svo_ipc_import(svo_ctx *ctx, const uint8_t handle[SVO_IPC_HANDLE_BYTES], void **device_ptr);
int svo_ipc_close(svo_ctx *ctx, void *device_ptr);
/* Stream-ordered fences between the GPUs of the tile partition, without a collective.  Every context owns a
 * counter; svo_fence_export gives its IPC handle (-> svo_ipc_import in the other processes).
 * svo_fence_signal(ctx, ptrs, n) enqueues one kernel that, after everything already in the stream, adds 1 to each
 * of the n counters (n = 0: the context's own counter); svo_fence_wait(ctx, target) enqueues a kernel that waits
 * until the context's own counter has reached `target` (modulo 2^32; gives up after ~2 s so that a dead peer
 * cannot hang the GPU).  Every counter has 16 slots.  Protocol of bench.py --partition tiles: peers signal the frame owner when their bands
 * are stored ("frame complete" = frames * n_gpus), the owner signals the peers when it has consumed the frame. */
int svo_fence_export(svo_ctx *ctx, uint8_t handle[SVO_IPC_HANDLE_BYTES]);
void *svo_fence_device_ptr(svo_ctx *ctx); /* this context's own counter, for fence lists that include the caller itself */
int svo_fence_signal(svo_ctx *ctx, void *const *fence_ptrs, int n, int slot);
int svo_fence_wait(svo_ctx *ctx, int slot, uint32_t target);
/* svo_fence_wait(slot, target) followed by svo_fence_signal(fence_ptrs, n, signal_slot) in ONE launch: what the frame's
 * owner runs per frame in the device-resident loop ("every GPU has stored frame k" -> "frame k consumed"). */
int svo_fence_wait_signal(svo_ctx *ctx, int slot, uint32_t target, void *const *fence_ptrs, int n, int signal_slot);
int svo_fence_reset(svo_ctx *ctx);

/* -- ray streams (new): n independent intersectOctree calls.
 *    svo_cast: host buffers in/out.  svo_cast_device: device buffers. */
int svo_cast(svo_ctx *ctx, const svo_ray *rays, uint64_t n, svo_hit *out, int maxDepth);
int svo_cast_device(svo_ctx *ctx, const void *d_rays, uint64_t n, void *d_out, int maxDepth);

/* -- measurement: CUDA events on the stream the kernels run on */
int svo_timer_begin(svo_ctx *ctx);
int svo_timer_end(svo_ctx *ctx, float *elapsed_ms); /* synchronises */
/* kernels launched by this context since creation (for bench.py gpu_launches) */
int svo_launch_count(const svo_ctx *ctx, uint64_t *count);
/* Instrumented render (validation kernels + counters), synchronous.  Writes all
 * planes like svo_render with SVO_OPT_AUX_PLANES and returns counters[0] =
 * intersectOctree calls, [1] = loop iterations (= child records the reference
 * fetches, svotrace.comp:294), [2] = bytes of those records in the reference
 * layout plus 7 per cast for the root (SURVEY 8d "algorithmic bytes"). */
int svo_render_stats(svo_ctx *ctx, const svo_frame *frame, uint64_t counters[3]);
/* The same counters for what the PRODUCTION kernel executes on this frame (content box on, casts that end before the
 * loop not spun to the cap): [1] is the number of loop iterations actually run -- bench.py reports it next to the
 * reference's count.  Writes the colour and depth planes like svo_render. */
int svo_render_stats_executed(svo_ctx *ctx, const svo_frame *frame, uint64_t counters[3]);
/* Gather roofline: random 32-byte-sector read rate over a working
 * set of `working_set_bytes`, measured with CUDA events.  Returns sectors/s. */
int svo_gather_probe(svo_ctx *ctx, uint64_t working_set_bytes, int loads_per_thread, double *sectors_per_s);

/* -- device-side deterministic math probe (tests only: compares the kernel's
 *    sin/cos/acos/exp/rand with the oracle bit for bit).  fn: 0 sin, 1 cos,
 *    2 acos, 3 exp, 4 rand(x, y). */
int svo_math_probe(svo_ctx *ctx, int fn, const float *x, const float *y, float *out, uint64_t n);

/* -- world generation (replaces Octree.constructCompleteOctree,
 *    Octree.java:192-353, for heightmap worlds).  Host-side, multi-threaded,
 *    byte-identical to the reference builder's stream.  height: n*n u16 (row =
 *    z), mat: n*n u8.  Returns bytes needed via *out_bytes; call with out=NULL
 *    to size.  chunk = CHUNK_SIZE (reference 1024). */
int svo_build_terrain(const uint16_t *height, const uint8_t *mat, int n, int chunk, uint8_t *out,
                      uint64_t cap, uint64_t *out_bytes, int nthreads);

/* Host-only view of the upload-time transcode (tests; no device needed): out[0] descriptors, [1] levels,
 * [2] FNV-1a hash of (descriptor.x, descriptor.y, reference child-block offset) over all descriptors, [3] 1 if
 * nothing in the tree can be hit, [4..6] x/y/z bounds of the non-empty leaves as (lo << 32 | hi) in units of
 * 2^-24 of the cube edge, [7] hash of the per-depth bounds.  desc_out (optional) receives the triples. */
int svo_transcode_probe(const uint8_t *nodes, uint64_t nbytes, int nthreads, uint64_t out[8], uint32_t *desc_out, uint64_t desc_cap);

/* The same 8 words as svo_transcode_probe, computed from what the context actually holds on the device. */
int svo_scene_probe(svo_ctx *ctx, uint64_t out[8]);

/* Deterministic synthetic inputs for benchmarks and tests (the reference's
 * 8192^2 heightmap / material PNGs are absent upstream): n*n u16 heights with
 * the value span of assets/heightmaps/nz.png and n*n u8 materials in {1,2,3}.
 * Host-side, multi-threaded, bit-reproducible for a given (n, seed). */
int svo_terrain_generate(int n, int seed, uint16_t *height, uint8_t *mat, int nthreads);

#ifdef __cplusplus
}
#endif
#endif /* SVO_B200_H */
