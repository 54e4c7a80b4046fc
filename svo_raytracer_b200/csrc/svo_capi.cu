// svo_capi.cu -- the C ABI declared in include/svo_b200.h.
//
// This is the replacement for the reference's GL dispatch boundary,
// src/engine/Renderer.java (addShader :43-54, addSSBO :123-129, updateSSBO
// :136-146, dispatchCompute :118-121) plus the raw GL43C calls in
// src/engine/Main.java (image setup :62-86, uniforms :269-283, depth readback
// :132-146).  No CPU fallback exists: without a CUDA device every call fails.
#include "../../include/svo_b200.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "svo_gpu_build.h"
#include "svo_kernels.h"
#include "svo_transcode.h"

using namespace svo;

static_assert(sizeof(svo_frame) == sizeof(FrameParams), "svo_frame and FrameParams must have one layout");
static_assert(sizeof(svo_ray) == 24 && sizeof(svo_hit) == 16, "ray-stream records are packed");

static thread_local std::string g_error;

constexpr int kLanes = 7;  // svo_select_lane: streams + colour/depth plane sets that may be in flight together

struct svo_ctx {
  int device = 0;
  int W = 0, H = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  // Two lanes (svo_select_lane): a stream and a colour/depth plane set each.  Work enqueued on different lanes may overlap on
  // the GPU: frame k+1's first tiles fill the SMs that frame k's last, longest tiles leave idle (the tail of a 1080p frame is
  // ~0.13 ms of a 0.4-1.4 ms kernel).  `stream` is always the current lane's stream.
  cudaStream_t own_lane_stream[kLanes] = {nullptr, nullptr, nullptr, nullptr}, lane_stream[kLanes] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_lane[kLanes] = {nullptr, nullptr, nullptr, nullptr};
  int lane = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int sm_count = 0;
  // scene
  uint8_t *d_raw = nullptr;
  uint64_t raw_cap = 0, nbytes = 0;
  uint2 *d_desc = nullptr;
  uint32_t *d_refbase = nullptr;
  uint2 *d_meta = nullptr;     // per descriptor: own record offset, parent index (incremental transcode)
  uint8_t *d_flag = nullptr;   // per descriptor: scratch of the incremental transcode, all zero between calls
  uint64_t desc_cap = 0;
  uint64_t ndesc_live = 0;     // descriptors reachable after the last whole transcode (the rest up to ndesc: patched-in, or garbage)
  uint8_t *d_stage = nullptr, *d_bitmap = nullptr;  // svo_upload_range: the caller's bytes before they are stored, changed-byte bitmap
  uint64_t stage_cap = 0, nbytes_before_range = 0;
  void *d_patch_arena = nullptr;  // scratch of the incremental transcode (work lists, counters): allocated once
  static constexpr size_t kPatchArena = 48u << 20;
  uint64_t patch_stats[4] = {0, 0, 0, 0};  // last svo_upload_range: dirty nodes, re-walked roots, descriptors appended, 1 = fell back to a whole transcode
  uint32_t ndesc = 0, nlevels = 0;
  uint32_t first_word_zero = 1;
  bool have_scene = false;
  // planes
  void *own[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  void *bound[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // second colour/depth set for pipelined read-back (svo_swap_buffers / svo_read_planes_async)
  void *back[kLanes][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};  // sets of lanes 1.. ([0] unused: set 0 = own[])
  int render_set = 0;  // which colour/depth set renders draw into and reads take from: 0 = own[], s = back[s]
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_rendered = nullptr, ev_copied[kLanes] = {nullptr, nullptr, nullptr, nullptr};
  // ray-stream scratch
  void *d_rays = nullptr, *d_hits = nullptr;
  uint64_t cast_cap = 0;
  uint32_t *d_sort = nullptr;  // 4 * sort_cap words (keys, keys_alt, idx, order) + CUB temp
  void *d_sort_temp = nullptr;
  uint64_t sort_cap = 0;
  size_t sort_temp_bytes = 0;
  // options
  int opt_aux = 0, opt_fast = 0, opt_kernel = 13, opt_l2 = 0, opt_sort = 2, opt_bounds = 1, opt_band_rows = 8, opt_gpu_transcode = 1, opt_stream_kernel = 0;
  CellBox leaf_box, depth_box[24];  // where casts can end in a hit (svo_transcode.h)
  unsigned int *d_tile_counter = nullptr;
  unsigned int *d_fence = nullptr;  // frame-complete counter peers signal over NVLink (svo_fence_*)
  bool fence_waits_unchecked = false;  // a k_fence_wait ran since svo_sync last looked at the watchdog latch (d_fence[7])
  struct IpcMap { cudaIpcMemHandle_t handle; void *base; };
  std::vector<IpcMap> ipc_maps;  // peer blocks opened by svo_ipc_import
  WaveWorkspace ws{};  // wavefront variant, allocated on first use
  void *ws_block = nullptr;
  float *d_beam_lattice[kLanes] = {nullptr, nullptr, nullptr, nullptr};  // conservative beam pre-pass: (W/4+1) x (H/4+1) lattice-ray distances, per lane
  float *lane_beam[kLanes] = {nullptr, nullptr, nullptr, nullptr};       // beam planes of lanes 1.. (lane 0: own[SVO_PLANE_BEAM]); allocated on first use
  void *split_block = nullptr;  // variant 15: record queue
  SplitQueue split = {};
  int ctas_per_sm = 8;
  uint64_t launches = 0;
  std::string err;
};

static int ensure_pipeline(svo_ctx *c, int set = 1);

// kernels whose workspace exists once per context (wavefront queues, split queue, the persistent kernel's counter) and the
// validation planes (one set) cannot run on two lanes at once: both lanes then share lane 0's stream
static bool lanes_share_stream(const svo_ctx *c) {
  return c->opt_aux || c->opt_kernel == 1 || c->opt_kernel == 2 || c->opt_kernel == 15 || c->opt_kernel == 16 || c->lane_stream[0] == c->lane_stream[1];
}
static void refresh_stream(svo_ctx *c) { c->stream = lanes_share_stream(c) ? c->lane_stream[0] : c->lane_stream[c->lane]; }
// order everything enqueued on the other lane before what the current lane enqueues next
static cudaError_t join_lanes(svo_ctx *c) {
  if (c->lane_stream[0] == c->lane_stream[1]) return cudaSuccess;
  for (int l = 0; l < kLanes; l++) {
    if (c->lane_stream[l] == c->stream) continue;
    cudaError_t e = cudaEventRecord(c->ev_lane[l], c->lane_stream[l]);
    if (e != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(c->stream, c->ev_lane[l], 0)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}
static cudaError_t sync_lanes(svo_ctx *c) {
  cudaError_t e = cudaStreamSynchronize(c->lane_stream[0]);
  for (int l = 1; l < kLanes && e == cudaSuccess; l++)
    if (c->lane_stream[l] != c->lane_stream[0]) e = cudaStreamSynchronize(c->lane_stream[l]);
  return e;
}

namespace {

int fail(svo_ctx *c, int code, const std::string &msg) {
  g_error = msg;
  if (c) c->err = msg;
  return code;
}
int cuda_fail(svo_ctx *c, cudaError_t e, const char *what) {
  std::string m = std::string(what) + ": " + cudaGetErrorString(e);
  cudaGetLastError();  // clear sticky-less errors
  if (e == cudaErrorMemoryAllocation) return fail(c, SVO_ERR_OOM, m);
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return fail(c, SVO_ERR_NO_DEVICE, m);
  return fail(c, SVO_ERR_CUDA, m);
}
#define SVO_CUDA(c, call)                                   \
  do {                                                      \
    cudaError_t e_ = (call);                                \
    if (e_ != cudaSuccess) return cuda_fail((c), e_, #call); \
  } while (0)

// device scratch freed on every return path
struct DevBuf {
  void *p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
  template <class T> T *as() const { return (T *)p; }
};

size_t plane_elem_bytes(int plane) {
  switch (plane) {
    case SVO_PLANE_COLOR_RGBA8: return 4;
    case SVO_PLANE_RADIANCE: return 16;
    default: return 4;
  }
}
size_t plane_elems(const svo_ctx *c, int plane) {
  if (plane == SVO_PLANE_BEAM) return (size_t)(c->W / 4) * (size_t)(c->H / 4);
  return (size_t)c->W * (size_t)c->H;
}
void *plane_ptr(const svo_ctx *c, int plane) {
  if (c->bound[plane]) return c->bound[plane];
  if (plane <= SVO_PLANE_DEPTH && c->render_set >= 1 && c->back[c->render_set][plane]) return c->back[c->render_set][plane];
  if (plane == SVO_PLANE_BEAM && c->render_set >= 1 && c->lane_beam[c->render_set]) return c->lane_beam[c->render_set];
  return c->own[plane];
}

int ensure_aux(svo_ctx *c) {
  for (int p = SVO_PLANE_HIT_ID; p <= SVO_PLANE_RADIANCE; p++) {
    if (!c->own[p]) {
      size_t bytes = plane_elems(c, p) * plane_elem_bytes(p);
      SVO_CUDA(c, cudaMalloc(&c->own[p], bytes ? bytes : 16));
      SVO_CUDA(c, cudaMemsetAsync(c->own[p], 0, bytes, c->stream));
    }
  }
  return SVO_OK;
}

// one allocation carved into the wavefront planes
int ensure_wavefront(svo_ctx *c) {
  if (c->ws_block) return SVO_OK;
  const uint64_t slots = (uint64_t)((c->W + 7) / 8) * (uint64_t)((c->H + 3) / 4) * 32u;
  const uint64_t planes = 2 * 2 + 2 + 2 * 6;  // ray A/B x2 queues, hit A/B, state 6 x2 queues
  const uint64_t bytes = planes * slots * sizeof(uint4) + 4096;
  SVO_CUDA(c, cudaMalloc(&c->ws_block, bytes));
  uint4 *p = (uint4 *)c->ws_block;
  auto take = [&]() { uint4 *r = p; p += slots; return r; };
  for (int q = 0; q < 2; q++) { c->ws.rayA[q] = take(); c->ws.rayB[q] = take(); }
  c->ws.hitA = take();
  c->ws.hitB = take();
  for (int q = 0; q < 2; q++)
    for (int k = 0; k < 6; k++) c->ws.state[q][k] = take();
  c->ws.counters = (unsigned *)p;
  c->ws.slots = slots;
  return SVO_OK;
}

int ensure_split(svo_ctx *c) {
  if (c->split_block) return SVO_OK;
  const uint64_t cap = (uint64_t)c->W * (uint64_t)c->H;
  SVO_CUDA(c, cudaMalloc(&c->split_block, 6 * cap * sizeof(uint4) + 256));
  uint4 *p = (uint4 *)c->split_block;
  for (int k = 0; k < 6; k++) c->split.q[k] = p + (size_t)k * cap;
  c->split.counters = (unsigned int *)(p + 6 * cap);
  c->split.capacity = cap;
  return SVO_OK;
}

SceneView scene_view(const svo_ctx *c, const svo_frame *frame = nullptr) {
  SceneView v;
  // content box of the frame (svo_transcode.h)
  content_box(c->leaf_box, c->depth_box, frame ? frame->maxDepth : -1, frame ? frame->coneDepth : -1, v.box_lo, v.box_hi);
  v.desc = c->d_desc;
  v.refbase = c->d_refbase;
  v.raw = c->d_raw;
  v.nbytes = c->nbytes;
  v.ndesc = c->ndesc;
  v.first_word_zero = c->first_word_zero;
  v.zero = 0u;
  v.one = 1u;
  v.two = 2u;
  v.four = 4u;
  v.exp_unit = 1u << 23;
  v.top = nullptr;
  v.ntop = 0;
  return v;
}
LaunchCfg launch_cfg(const svo_ctx *c) {
  LaunchCfg l;
  l.fast = c->opt_fast != 0;
  l.aux = c->opt_aux != 0;
  l.box = false;
  l.kernel = c->opt_kernel;
  l.stream_kernel = c->opt_stream_kernel;
  l.sm_count = c->sm_count;
  l.scene_levels = (int)c->nlevels;
  l.band_stride = 0;
  l.band_offset = 0;
  l.band_ctas = c->opt_band_rows / 8;
  l.ctas_per_sm = c->ctas_per_sm;
  l.tile_counter = c->d_tile_counter;
  l.tile_queue = c->d_tile_counter + 2 + 2 * c->lane;  // two words per lane, zero between launches (the last CTA out resets them)
  l.fences.n = 0;
  for (int i = 0; i < 16; i++) l.fences.p[i] = nullptr;
  l.split = c->split;
  return l;
}
Planes planes_of(const svo_ctx *c) {
  Planes p;
  p.rgba8 = (uchar4 *)plane_ptr(c, SVO_PLANE_COLOR_RGBA8);
  p.depth = (float *)plane_ptr(c, SVO_PLANE_DEPTH);
  p.beam = (const float *)plane_ptr(c, SVO_PLANE_BEAM);
  p.hit_id = (uint32_t *)plane_ptr(c, SVO_PLANE_HIT_ID);
  p.iter = (uint32_t *)plane_ptr(c, SVO_PLANE_ITER);
  p.primary_t = (float *)plane_ptr(c, SVO_PLANE_PRIMARY_T);
  p.radiance = (float4 *)plane_ptr(c, SVO_PLANE_RADIANCE);
  return p;
}

int check_frame(svo_ctx *c, const svo_frame *f) {
  if (!f) return fail(c, SVO_ERR_INVALID, "frame is NULL");
  if (f->maxDepth < 1 || f->maxDepth > 23) return fail(c, SVO_ERR_INVALID, "maxDepth must be in [1,23]");
  if (f->coneDepth < 1 || f->coneDepth > 23) return fail(c, SVO_ERR_INVALID, "coneDepth must be in [1,23]");
  if (f->casts < 0 || f->casts > 64) return fail(c, SVO_ERR_INVALID, "casts must be in [0,64]");
  if ((f->flags & ~(SVO_FRAME_ACCUMULATE | SVO_FRAME_BEAM_FLOOR)) != 0) return fail(c, SVO_ERR_INVALID, "unknown bits in flags");
  if ((f->flags & SVO_FRAME_BEAM_FLOOR) && f->useBeam)
    return fail(c, SVO_ERR_INVALID, "the beam plane holds either upstream's beam distances (useBeam) or conservative bounds (SVO_FRAME_BEAM_FLOOR), not both");
  return SVO_OK;
}

// cudaIpcGetMemHandle names the whole cudaMalloc block that contains a pointer and cudaIpcOpenMemHandle returns
// that block's base (small allocations share a block), so the handle blob carries the offset as well:
// bytes [0,64) cudaIpcMemHandle_t, bytes [64,72) offset of the exported pointer inside its block.
int export_handle(svo_ctx *c, void *ptr, uint8_t handle[72]) {
  typedef int (*GetRange)(unsigned long long *, size_t *, unsigned long long);
  static GetRange get_range = nullptr;
  if (!get_range) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn)
      return fail(c, SVO_ERR_CUDA, "cuMemGetAddressRange is not available");
    get_range = (GetRange)fn;
  }
  unsigned long long base = 0;
  size_t size = 0;
  if (get_range(&base, &size, (unsigned long long)(uintptr_t)ptr) != 0) return fail(c, SVO_ERR_CUDA, "cuMemGetAddressRange failed");
  cudaIpcMemHandle_t hnd;
  SVO_CUDA(c, cudaIpcGetMemHandle(&hnd, (void *)(uintptr_t)base));
  memcpy(handle, &hnd, 64);
  const uint64_t off = (uint64_t)(uintptr_t)ptr - base;
  memcpy(handle + 64, &off, 8);
  return SVO_OK;
}

// The content-box shortcut changes only the iteration count of casts that miss; it is used where nothing shows
// that count: not in render mode 1 (iteration heat map, svotrace.comp:561-571) and not with the validation planes.
bool box_allowed(const svo_ctx *c, const svo_frame *f) { return c->opt_bounds && !c->opt_aux && f->renderMode != 1; }

// (re)build the device descriptor arrays from the node stream in d_raw
int retranscode(svo_ctx *c) {
  uint64_t nd = 0;
  uint32_t nlevels = 0;
  CellBox leaf_box, depth_box[24];
  bool done = false;
  c->have_scene = false;  // until the descriptor arrays match d_raw again (a refused stream leaves the context without a scene)
  // the kernels may still be reading the previous arrays
  SVO_CUDA(c, sync_lanes(c));
  // final arrays: 25 % head-room, which the incremental transcode (svo_upload_range) appends into
  auto ensure_arrays = [&](uint64_t n) -> int {
    if (n + n / 8 <= c->desc_cap && c->d_meta && c->d_flag) return SVO_OK;
    if (c->d_desc) cudaFree(c->d_desc);
    if (c->d_refbase) cudaFree(c->d_refbase);
    if (c->d_meta) cudaFree(c->d_meta);
    if (c->d_flag) cudaFree(c->d_flag);
    c->d_desc = nullptr; c->d_refbase = nullptr; c->d_meta = nullptr; c->d_flag = nullptr;
    c->desc_cap = 0;
    const uint64_t want = n + n / 4 + 4096;
    SVO_CUDA(c, cudaMalloc((void **)&c->d_desc, want * sizeof(uint2)));
    SVO_CUDA(c, cudaMalloc((void **)&c->d_refbase, want * sizeof(uint32_t)));
    SVO_CUDA(c, cudaMalloc((void **)&c->d_meta, want * sizeof(uint2)));
    SVO_CUDA(c, cudaMalloc((void **)&c->d_flag, want));
    c->desc_cap = want;
    return SVO_OK;
  };
  if (c->opt_gpu_transcode) {
    // a tree has at most one interior record per 7 bytes; anything larger is aliased/cyclic and goes to the host
    // path, which reports it
    const uint64_t cap = c->nbytes / 7 + 4096;
    DevBuf tmp_desc, tmp_ref, tmp_meta;
    SVO_CUDA(c, tmp_desc.alloc(cap * sizeof(uint2)));
    SVO_CUDA(c, tmp_ref.alloc(cap * sizeof(uint32_t)));
    SVO_CUDA(c, tmp_meta.alloc(cap * sizeof(uint2)));
    bool overflow = false;
    SVO_CUDA(c, gpu_transcode(c->d_raw, c->nbytes, tmp_desc.as<uint2>(), tmp_ref.as<uint32_t>(), tmp_meta.as<uint2>(), cap, &nd, &nlevels, &leaf_box,
                              depth_box, &overflow, c->stream));
    c->launches += 2 + 3 * (uint64_t)nlevels;
    if (!overflow) {
      int rc = ensure_arrays(nd);
      if (rc) return rc;
      SVO_CUDA(c, cudaMemcpyAsync(c->d_desc, tmp_desc.p, nd * sizeof(uint2), cudaMemcpyDeviceToDevice, c->stream));
      SVO_CUDA(c, cudaMemcpyAsync(c->d_refbase, tmp_ref.p, nd * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
      SVO_CUDA(c, cudaMemcpyAsync(c->d_meta, tmp_meta.p, nd * sizeof(uint2), cudaMemcpyDeviceToDevice, c->stream));
      SVO_CUDA(c, cudaMemsetAsync(c->d_flag, 0, c->desc_cap, c->stream));
      SVO_CUDA(c, cudaStreamSynchronize(c->stream));
      done = true;
    }
  }
  if (!done) {  // host path (SVO_OPT_GPU_TRANSCODE 0, or a stream the device pass refused)
    std::vector<uint8_t> h_raw(c->nbytes);
    if (c->nbytes) SVO_CUDA(c, cudaMemcpy(h_raw.data(), c->d_raw, c->nbytes, cudaMemcpyDeviceToHost));
    Transcoded t;
    std::string err;
    if (!transcode_stream(h_raw.data(), c->nbytes, t, err)) return fail(c, SVO_ERR_FORMAT, err);
    nd = t.desc.size();
    int rc = ensure_arrays(nd);
    if (rc) return rc;
    SVO_CUDA(c, cudaMemcpyAsync(c->d_desc, t.desc.data(), nd * sizeof(uint2), cudaMemcpyHostToDevice, c->stream));
    SVO_CUDA(c, cudaMemcpyAsync(c->d_refbase, t.refbase.data(), nd * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    SVO_CUDA(c, cudaMemcpyAsync(c->d_meta, t.meta.data(), nd * sizeof(uint2), cudaMemcpyHostToDevice, c->stream));
    SVO_CUDA(c, cudaMemsetAsync(c->d_flag, 0, c->desc_cap, c->stream));
    SVO_CUDA(c, cudaStreamSynchronize(c->stream));  // t goes out of scope
    nlevels = (uint32_t)t.level_start.size();
    leaf_box = t.leaf_box;
    for (int d = 0; d < 24; d++) depth_box[d] = t.depth_box[d];
  }
  c->ndesc_live = nd;
  c->ndesc = (uint32_t)nd;
  c->nlevels = nlevels;
  uint32_t w0 = 0;
  if (c->nbytes) SVO_CUDA(c, cudaMemcpy(&w0, c->d_raw, c->nbytes < 4 ? c->nbytes : 4, cudaMemcpyDeviceToHost));
  c->first_word_zero = (w0 == 0);
  c->leaf_box = leaf_box;
  for (int d = 0; d < 24; d++) c->depth_box[d] = depth_box[d];
  c->have_scene = true;

  // L2 access-policy window over the hot upper levels (a prefix of the BFS array)
  if (c->opt_l2) {
    cudaDeviceProp prop;
    SVO_CUDA(c, cudaGetDeviceProperties(&prop, c->device));
    size_t win = nd * sizeof(uint2);
    if ((size_t)prop.accessPolicyMaxWindowSize < win) win = (size_t)prop.accessPolicyMaxWindowSize;
    size_t persist = (size_t)prop.persistingL2CacheMaxSize;
    if (persist > 0 && win > 0) {
      if (win > persist) win = persist;
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist);
      cudaStreamAttrValue attr;
      memset(&attr, 0, sizeof attr);
      attr.accessPolicyWindow.base_ptr = c->d_desc;
      attr.accessPolicyWindow.num_bytes = win;
      attr.accessPolicyWindow.hitRatio = 1.0f;
      attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
      cudaGetLastError();
    }
  }
  return SVO_OK;
}

}  // namespace

// ---- pipelined read-back: render frame s+1 while frame s travels to the host -----------------------------------
static int ensure_pipeline(svo_ctx *c, int set) {
  if (!c->copy_stream) {
    SVO_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    SVO_CUDA(c, cudaEventCreateWithFlags(&c->ev_rendered, cudaEventDisableTiming));
    for (int p = 0; p < kLanes; p++) SVO_CUDA(c, cudaEventCreateWithFlags(&c->ev_copied[p], cudaEventDisableTiming));
  }
  if (set >= 1 && set < kLanes && !c->back[set][0]) {
    for (int p = 0; p < 2; p++) {
      const size_t bytes = plane_elems(c, p) * plane_elem_bytes(p);
      SVO_CUDA(c, cudaMalloc(&c->back[set][p], bytes));
      SVO_CUDA(c, cudaMemsetAsync(c->back[set][p], 0, bytes, c->stream));
    }
    SVO_CUDA(c, cudaStreamSynchronize(c->stream));  // other lanes may use the set next
  }
  return SVO_OK;
}

extern "C" {

int svo_abi_version(void) { return SVO_ABI_VERSION; }

int svo_device_count(int *count) {
  if (!count) return fail(nullptr, SVO_ERR_INVALID, "count is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    return cuda_fail(nullptr, e, "cudaGetDeviceCount");
  }
  *count = n;
  return SVO_OK;
}

int svo_create(svo_ctx **out, int device, int width, int height) {
  if (!out) return fail(nullptr, SVO_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (width <= 0 || height <= 0 || (uint64_t)width * (uint64_t)height > (1ull << 31))
    return fail(nullptr, SVO_ERR_INVALID, "bad image size");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, SVO_ERR_NO_DEVICE,
                std::string("no CUDA device (this path has no CPU fallback): ") + (e != cudaSuccess ? cudaGetErrorString(e) : "0 devices"));
  if (device < 0 || device >= n) return fail(nullptr, SVO_ERR_INVALID, "device index out of range");
  svo_ctx *c = new svo_ctx();
  c->device = device;
  c->W = width;
  c->H = height;
  int rc = SVO_OK;
  do {
    if ((e = cudaSetDevice(device)) != cudaSuccess) { rc = cuda_fail(nullptr, e, "cudaSetDevice"); break; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { rc = cuda_fail(nullptr, e, "cudaGetDeviceProperties"); break; }
    c->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess) { rc = cuda_fail(nullptr, e, "cudaStreamCreate"); break; }
    c->own_lane_stream[0] = c->own_stream;
    for (int l = 1; l < kLanes && e == cudaSuccess; l++) e = cudaStreamCreateWithFlags(&c->own_lane_stream[l], cudaStreamNonBlocking);
    if (e != cudaSuccess) { rc = cuda_fail(nullptr, e, "cudaStreamCreate"); break; }
    for (int l = 0; l < kLanes; l++) c->lane_stream[l] = c->own_lane_stream[l];
    c->stream = c->own_stream;
    for (int l = 0; l < kLanes; l++)
      if ((e = cudaEventCreateWithFlags(&c->ev_lane[l], cudaEventDisableTiming)) != cudaSuccess) break;
    if (e != cudaSuccess) { rc = cuda_fail(nullptr, e, "cudaEventCreate"); break; }
    if ((e = cudaEventCreate(&c->ev0)) != cudaSuccess || (e = cudaEventCreate(&c->ev1)) != cudaSuccess) { rc = cuda_fail(nullptr, e, "cudaEventCreate"); break; }
    if ((e = cudaMalloc((void **)&c->d_tile_counter, 256)) != cudaSuccess) { rc = cuda_fail(nullptr, e, "cudaMalloc(counter)"); break; }
    cudaMemsetAsync(c->d_tile_counter, 0, 256, c->stream);
    if ((e = cudaMalloc((void **)&c->d_fence, 512)) != cudaSuccess) { rc = cuda_fail(nullptr, e, "cudaMalloc(fence)"); break; }
    cudaMemsetAsync(c->d_fence, 0, 512, c->stream);
    for (int p = SVO_PLANE_COLOR_RGBA8; p <= SVO_PLANE_BEAM; p++) {
      size_t bytes = plane_elems(c, p) * plane_elem_bytes(p);
      if ((e = cudaMalloc(&c->own[p], bytes ? bytes : 16)) != cudaSuccess) { rc = cuda_fail(nullptr, e, "cudaMalloc(plane)"); break; }
      cudaMemsetAsync(c->own[p], 0, bytes, c->stream);
    }
  } while (0);
  if (rc != SVO_OK) {
    svo_destroy(c);
    return rc;
  }
  *out = c;
  return SVO_OK;
}

void svo_destroy(svo_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->own_stream) cudaStreamSynchronize(c->own_stream);
  for (int l = 1; l < kLanes; l++)
    if (c->own_lane_stream[l]) cudaStreamSynchronize(c->own_lane_stream[l]);
  for (auto &m : c->ipc_maps) cudaIpcCloseMemHandle(m.base);
  for (int p = 0; p < 7; p++)
    if (c->own[p]) cudaFree(c->own[p]);
  if (c->d_raw) cudaFree(c->d_raw);
  if (c->d_desc) cudaFree(c->d_desc);
  if (c->d_refbase) cudaFree(c->d_refbase);
  if (c->d_meta) cudaFree(c->d_meta);
  if (c->d_flag) cudaFree(c->d_flag);
  if (c->d_stage) cudaFree(c->d_stage);
  if (c->d_bitmap) cudaFree(c->d_bitmap);
  if (c->d_patch_arena) cudaFree(c->d_patch_arena);
  if (c->d_rays) cudaFree(c->d_rays);
  if (c->d_hits) cudaFree(c->d_hits);
  for (int l = 1; l < kLanes; l++)
    for (int p = 0; p < 2; p++)
      if (c->back[l][p]) cudaFree(c->back[l][p]);
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  if (c->ev_rendered) cudaEventDestroy(c->ev_rendered);
  for (int p = 0; p < kLanes; p++)
    if (c->ev_copied[p]) cudaEventDestroy(c->ev_copied[p]);
  if (c->d_sort) cudaFree(c->d_sort);
  if (c->d_sort_temp) cudaFree(c->d_sort_temp);
  if (c->d_tile_counter) cudaFree(c->d_tile_counter);
  if (c->d_fence) cudaFree(c->d_fence);
  if (c->ws_block) cudaFree(c->ws_block);
  if (c->split_block) cudaFree(c->split_block);
  for (int l = 0; l < kLanes; l++) {
    if (c->d_beam_lattice[l]) cudaFree(c->d_beam_lattice[l]);
    if (c->lane_beam[l]) cudaFree(c->lane_beam[l]);
  }
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  for (int l = 1; l < kLanes; l++)
    if (c->own_lane_stream[l]) cudaStreamDestroy(c->own_lane_stream[l]);
  for (int l = 0; l < kLanes; l++)
    if (c->ev_lane[l]) cudaEventDestroy(c->ev_lane[l]);
  cudaGetLastError();
  delete c;
}

const char *svo_last_error(const svo_ctx *c) { return c ? c->err.c_str() : g_error.c_str(); }

int svo_set_option(svo_ctx *c, int option, int64_t value) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  switch (option) {
    case SVO_OPT_AUX_PLANES:
      if (cudaSetDevice(c->device) == cudaSuccess) sync_lanes(c);
      c->opt_aux = value != 0;
      refresh_stream(c);
      return SVO_OK;
    case SVO_OPT_FAST_MATH: c->opt_fast = value != 0; return SVO_OK;
    case SVO_OPT_KERNEL:
      if (value != 0 && value != 1 && value != 2 && (value < 4 || value > 26)) return fail(c, SVO_ERR_INVALID, "unknown kernel variant");
      if (cudaSetDevice(c->device) == cudaSuccess) sync_lanes(c);
      c->opt_kernel = (int)value;
      refresh_stream(c);
      return SVO_OK;
    case SVO_OPT_L2_PERSIST: c->opt_l2 = value != 0; return SVO_OK;
    case SVO_OPT_RAY_SORT:
      if (value < 0 || value > 2) return fail(c, SVO_ERR_INVALID, "SVO_OPT_RAY_SORT is 0, 1 or 2");
      c->opt_sort = (int)value;
      return SVO_OK;
    case SVO_OPT_CONTENT_BOUNDS: c->opt_bounds = value != 0; return SVO_OK;
    case SVO_OPT_GPU_TRANSCODE: c->opt_gpu_transcode = value != 0; return SVO_OK;
    case SVO_OPT_STREAM_KERNEL:
      if (value < 0 || value > 2) return fail(c, SVO_ERR_INVALID, "unknown ray-stream kernel");
      c->opt_stream_kernel = (int)value;
      return SVO_OK;
    case SVO_OPT_BAND_ROWS:
      if (value < 8 || value > 4096 || value % 8) return fail(c, SVO_ERR_INVALID, "band rows must be a multiple of 8");
      c->opt_band_rows = (int)value;
      return SVO_OK;
    default: return fail(c, SVO_ERR_INVALID, "unknown option");
  }
}
int svo_get_option(const svo_ctx *c, int option, int64_t *value) {
  if (!c || !value) return fail(nullptr, SVO_ERR_INVALID, "NULL argument");
  switch (option) {
    case SVO_OPT_AUX_PLANES: *value = c->opt_aux; return SVO_OK;
    case SVO_OPT_FAST_MATH: *value = c->opt_fast; return SVO_OK;
    case SVO_OPT_KERNEL: *value = c->opt_kernel; return SVO_OK;
    case SVO_OPT_L2_PERSIST: *value = c->opt_l2; return SVO_OK;
    case SVO_OPT_RAY_SORT: *value = c->opt_sort; return SVO_OK;
    case SVO_OPT_CONTENT_BOUNDS: *value = c->opt_bounds; return SVO_OK;
    case SVO_OPT_BAND_ROWS: *value = c->opt_band_rows; return SVO_OK;
    case SVO_OPT_GPU_TRANSCODE: *value = c->opt_gpu_transcode; return SVO_OK;
    case SVO_OPT_STREAM_KERNEL: *value = c->opt_stream_kernel; return SVO_OK;
    default: return fail(const_cast<svo_ctx *>(c), SVO_ERR_INVALID, "unknown option");
  }
}

int svo_set_stream(svo_ctx *c, void *cuda_stream) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  SVO_CUDA(c, cudaSetDevice(c->device));
  SVO_CUDA(c, sync_lanes(c));
  // a caller-owned stream carries both lanes (no overlap between frames); NULL restores the context's two streams
  for (int l = 0; l < kLanes; l++) c->lane_stream[l] = cuda_stream ? (cudaStream_t)cuda_stream : c->own_lane_stream[l];
  refresh_stream(c);
  return SVO_OK;
}

int svo_upload(svo_ctx *c, const uint8_t *nodes, uint64_t nbytes) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  if (!nodes && nbytes) return fail(c, SVO_ERR_INVALID, "nodes is NULL");
  if (nbytes >= (1ull << 32)) return fail(c, SVO_ERR_INVALID, "node stream must be < 4 GiB");
  SVO_CUDA(c, cudaSetDevice(c->device));
  c->have_scene = false;
  SVO_CUDA(c, join_lanes(c));
  if (nbytes + 16 > c->raw_cap) {
    SVO_CUDA(c, sync_lanes(c));
    if (c->d_raw) cudaFree(c->d_raw);
    c->d_raw = nullptr;
    c->raw_cap = 0;
    const uint64_t cap = nbytes + nbytes / 16 + 4096;
    SVO_CUDA(c, cudaMalloc((void **)&c->d_raw, cap));
    c->raw_cap = cap;
  }
  c->nbytes = nbytes;
  if (nbytes) SVO_CUDA(c, cudaMemcpyAsync(c->d_raw, nodes, nbytes, cudaMemcpyHostToDevice, c->stream));
  return retranscode(c);
}

int svo_upload_range(svo_ctx *c, const uint8_t *nodes, uint64_t start, uint64_t end) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  if (!c->have_scene) return fail(c, SVO_ERR_NO_SCENE, "svo_upload_range before svo_upload");
  // Renderer.updateSSBO prints "Update SSBO error: Invalid parameters." and returns (Renderer.java:137-140)
  if (!nodes || start >= end) return fail(c, SVO_ERR_INVALID, "Update SSBO error: Invalid parameters.");
  if (end >= (1ull << 32)) return fail(c, SVO_ERR_INVALID, "node stream must be < 4 GiB");
  SVO_CUDA(c, cudaSetDevice(c->device));
  SVO_CUDA(c, join_lanes(c));
  if (end + 16 > c->raw_cap) {  // appended nodes outgrew the allocation: move the stream
    uint8_t *nr = nullptr;
    const uint64_t cap = end + end / 4 + 4096;
    SVO_CUDA(c, cudaMalloc((void **)&nr, cap));
    SVO_CUDA(c, sync_lanes(c));
    if (c->nbytes) SVO_CUDA(c, cudaMemcpy(nr, c->d_raw, c->nbytes, cudaMemcpyDeviceToDevice));
    cudaFree(c->d_raw);
    c->d_raw = nr;
    c->raw_cap = cap;
  }
  c->nbytes_before_range = c->nbytes;
  if (end > c->nbytes) {
    // bytes between the old end and `start` were never uploaded: they read as zero
    if (start > c->nbytes) SVO_CUDA(c, cudaMemsetAsync(c->d_raw + c->nbytes, 0, start - c->nbytes, c->stream));
    c->nbytes = end;
  }
  // Store the bytes and find out which of them really changed (the engine's ranges run from the first to the last touched
  // record: mostly unchanged bytes), then re-walk only the subtrees those bytes belong to (gpu_patch).
  const uint64_t len = end - start, old_nbytes = c->nbytes_before_range;
  if (len + 16 > c->stage_cap) {
    SVO_CUDA(c, sync_lanes(c));
    if (c->d_stage) cudaFree(c->d_stage);
    if (c->d_bitmap) cudaFree(c->d_bitmap);
    c->d_stage = c->d_bitmap = nullptr;
    c->stage_cap = 0;
    const uint64_t cap = len + len / 2 + 4096;
    SVO_CUDA(c, cudaMalloc((void **)&c->d_stage, cap));
    SVO_CUDA(c, cudaMalloc((void **)&c->d_bitmap, cap / 8 + 16));
    c->stage_cap = cap;
  }
  if (!c->d_patch_arena) SVO_CUDA(c, cudaMalloc(&c->d_patch_arena, svo_ctx::kPatchArena));
  SVO_CUDA(c, cudaMemcpyAsync(c->d_stage, nodes + start, len, cudaMemcpyHostToDevice, c->stream));
  uint64_t span[2] = {0, 0};
  SVO_CUDA(c, gpu_diff_apply(c->d_raw, c->d_stage, start, end, old_nbytes, c->d_bitmap, span, c->d_patch_arena, svo_ctx::kPatchArena, c->stream));
  c->launches++;
  for (int k = 0; k < 4; k++) c->patch_stats[k] = 0;
  if (span[0] >= span[1]) return SVO_OK;  // nothing changed: the descriptors are still right
  bool fallback = !c->opt_gpu_transcode || !c->d_meta || !c->d_flag;
  if (!fallback) {
    uint64_t nd = c->ndesc;
    c->have_scene = false;
    SVO_CUDA(c, gpu_patch(c->d_raw, c->nbytes, c->d_bitmap, start, end, span, c->d_desc, c->d_refbase, c->d_meta, c->d_flag, c->desc_cap, &nd,
                          &c->leaf_box, c->depth_box, &fallback, c->patch_stats, c->d_patch_arena, svo_ctx::kPatchArena, c->stream));
    c->launches += 8;
    if (!fallback) {
      c->ndesc = (uint32_t)nd;
      // whole transcodes recompute what patches only ever grow (content boxes) and drop what they leave behind (unreachable
      // descriptors): do one when the patched-in part has grown to a quarter of the array
      if (c->ndesc - c->ndesc_live > c->ndesc_live / 4 + 4096) fallback = true;
    }
  }
  if (fallback) {
    c->patch_stats[3] = 1;
    return retranscode(c);
  }
  if (span[0] < 4) {  // octreeBuffer[0] == 0 feeds the debug overlay (svotrace.comp:696)
    uint32_t w0 = 0;
    SVO_CUDA(c, cudaMemcpy(&w0, c->d_raw, c->nbytes < 4 ? c->nbytes : 4, cudaMemcpyDeviceToHost));
    c->first_word_zero = (w0 == 0);
  }
  c->have_scene = true;
  return SVO_OK;
}

int svo_build_terrain_device(svo_ctx *c, const uint16_t *height, const uint8_t *mat, int n, int chunk, uint64_t *out_bytes) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  if (!height || !mat || n < 2 || (n & (n - 1)) || chunk < 2 || (chunk & (chunk - 1))) return fail(c, SVO_ERR_INVALID, "bad heightmap / size / chunk");
  SVO_CUDA(c, cudaSetDevice(c->device));
  SVO_CUDA(c, sync_lanes(c));  // kernels may still be reading the previous scene
  uint8_t *stream = nullptr;
  uint64_t nbytes = 0, cap = 0, nl = 0;
  bool unsupported = false;
  SVO_CUDA(c, gpu_build_terrain(height, mat, n, chunk, &stream, &nbytes, &cap, &unsupported, c->stream, &nl));
  if (unsupported)
    return fail(c, SVO_ERR_INVALID, "the device builder takes chunk >= 4, at most 65535 sub-octrees and streams < 4 GiB; "
                                    "use svo_build_terrain + svo_upload for this shape");
  c->launches += nl;
  c->have_scene = false;
  if (c->d_raw) cudaFree(c->d_raw);
  c->d_raw = stream;
  c->raw_cap = cap;
  c->nbytes = nbytes;
  if (out_bytes) *out_bytes = nbytes;
  return retranscode(c);
}

int svo_download(svo_ctx *c, uint8_t *dst, uint64_t cap) {
  if (!c || !dst) return fail(nullptr, SVO_ERR_INVALID, "NULL argument");
  if (!c->have_scene) return fail(c, SVO_ERR_NO_SCENE, "svo_download before a scene exists");
  if (cap < c->nbytes) return fail(c, SVO_ERR_INVALID, "dst too small");
  SVO_CUDA(c, cudaSetDevice(c->device));
  if (c->nbytes) SVO_CUDA(c, cudaMemcpyAsync(dst, c->d_raw, c->nbytes, cudaMemcpyDeviceToHost, c->stream));
  SVO_CUDA(c, cudaStreamSynchronize(c->stream));
  return SVO_OK;
}

int svo_scene_info(const svo_ctx *c, uint64_t info[4]) {
  if (!c || !info) return fail(nullptr, SVO_ERR_INVALID, "NULL argument");
  info[0] = c->nbytes;
  info[1] = c->ndesc;
  info[2] = c->nlevels;
  info[3] = c->raw_cap + c->desc_cap * (sizeof(uint2) + sizeof(uint32_t) + sizeof(uint2) + 1);
  return SVO_OK;
}

int svo_render_rows(svo_ctx *c, const svo_frame *frame, int y0, int y1) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  if (!c->have_scene) return fail(c, SVO_ERR_NO_SCENE, "svo_render before svo_upload");
  int rc = check_frame(c, frame);
  if (rc) return rc;
  if (y0 < 0 || y1 > c->H || y0 > y1) return fail(c, SVO_ERR_INVALID, "row range outside the image");
  SVO_CUDA(c, cudaSetDevice(c->device));
  if (c->opt_aux && (rc = ensure_aux(c)) != SVO_OK) return rc;
  FrameParams fp;
  memcpy(&fp, frame, sizeof fp);
  if (c->opt_kernel == 2) {
    if (c->W > 65535 || c->H > 65535) return fail(c, SVO_ERR_INVALID, "the wavefront kernel packs pixel coordinates in 16 bits: width and height must be <= 65535");
    if ((rc = ensure_wavefront(c)) != SVO_OK) return rc;
    SVO_CUDA(c, launch_render_wavefront(launch_cfg(c), scene_view(c), fp, planes_of(c), c->W, c->H, y0, y1, c->ws, c->stream));
    c->launches += (uint64_t)wavefront_launches(fp);
    return SVO_OK;
  }
  if ((c->opt_kernel == 15 || c->opt_kernel == 16) && (rc = ensure_split(c)) != SVO_OK) return rc;
  LaunchCfg cfg = launch_cfg(c);
  cfg.box = box_allowed(c, frame);
  SVO_CUDA(c, launch_render(cfg, scene_view(c, frame), fp, planes_of(c), c->W, c->H, y0, y1, c->stream));
  c->launches += (uint64_t)render_launches(cfg, fp);
  return SVO_OK;
}
int svo_render(svo_ctx *c, const svo_frame *frame) { return svo_render_rows(c, frame, 0, c ? c->H : 0); }

int svo_render_interleaved_signal(svo_ctx *c, const svo_frame *frame, int part, int parts, void *const *fence_ptrs, int n, int slot) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  if (!c->have_scene) return fail(c, SVO_ERR_NO_SCENE, "svo_render before svo_upload");
  int rc = check_frame(c, frame);
  if (rc) return rc;
  if (parts < 1 || part < 0 || part >= parts) return fail(c, SVO_ERR_INVALID, "part must be in [0, parts)");
  if (n < -1 || n > 16 || (n > 0 && !fence_ptrs)) return fail(c, SVO_ERR_INVALID, "bad fence list (at most 16)");
  if (n >= 0 && (slot < 0 || slot >= 16)) return fail(c, SVO_ERR_INVALID, "fence slot must be in [0,16)");
  SVO_CUDA(c, cudaSetDevice(c->device));
  if (c->opt_aux && (rc = ensure_aux(c)) != SVO_OK) return rc;
  FrameParams fp;
  memcpy(&fp, frame, sizeof fp);
  LaunchCfg cfg = launch_cfg(c);
  // the band-interleaved partition is a feature of the tile kernels: 17 (tile queue, the default) / 14 (wide stack
  // entries, static CTAs) / 0 (8-byte stack entries; also the only one with SVO_OPT_FAST_MATH)
  cfg.kernel = c->opt_fast ? 0 : (c->opt_kernel == 0 || c->opt_kernel == 14) ? c->opt_kernel : 17;
  cfg.band_stride = parts;
  cfg.band_offset = part;
  cfg.box = box_allowed(c, frame);
  FenceList fl;
  fl.n = 0;
  for (int i = 0; i < 16; i++) fl.p[i] = nullptr;
  if (n >= 0) {  // n == 0: this context's own fence (as svo_fence_signal)
    fl.n = n > 0 ? n : 1;
    if (n == 0) fl.p[0] = c->d_fence + 8 * slot;
    for (int i = 0; i < n; i++) fl.p[i] = (unsigned int *)fence_ptrs[i] + 8 * slot;
  }
  if (cfg.kernel == 17) cfg.fences = fl;  // signalled by the render kernel's last CTA
  SVO_CUDA(c, launch_render(cfg, scene_view(c, frame), fp, planes_of(c), c->W, c->H, 0, c->H, c->stream));
  c->launches++;
  if (cfg.kernel != 17 && fl.n > 0) {
    SVO_CUDA(c, launch_fence_signal(fl, c->stream));
    c->launches++;
  }
  return SVO_OK;
}
int svo_render_interleaved(svo_ctx *c, const svo_frame *frame, int part, int parts) {
  return svo_render_interleaved_signal(c, frame, part, parts, nullptr, -1, 0);
}

int svo_beam(svo_ctx *c, const svo_frame *frame) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  if (!c->have_scene) return fail(c, SVO_ERR_NO_SCENE, "svo_beam before svo_upload");
  int rc = check_frame(c, frame);
  if (rc) return rc;
  SVO_CUDA(c, cudaSetDevice(c->device));
  SVO_CUDA(c, join_lanes(c));  // one beam plane per context
  FrameParams fp;
  memcpy(&fp, frame, sizeof fp);
  SVO_CUDA(c, launch_beam(launch_cfg(c), scene_view(c), fp, (float *)plane_ptr(c, SVO_PLANE_BEAM), c->W, c->H, c->stream));
  c->launches++;
  return SVO_OK;
}

static int ensure_lattice(svo_ctx *c, int l) {
  const size_t lat_bytes = (size_t)(c->W / 4 + 1) * (size_t)(c->H / 4 + 1) * sizeof(float) + 16;
  if (!c->d_beam_lattice[l]) {
    SVO_CUDA(c, cudaMalloc((void **)&c->d_beam_lattice[l], lat_bytes));
    SVO_CUDA(c, cudaMemsetAsync(c->d_beam_lattice[l], 0, lat_bytes, c->stream));
  }
  if (l >= 1 && !c->lane_beam[l]) {
    const size_t bytes = plane_elems(c, SVO_PLANE_BEAM) * sizeof(float);
    SVO_CUDA(c, cudaMalloc((void **)&c->lane_beam[l], bytes ? bytes : 16));
    SVO_CUDA(c, cudaMemsetAsync(c->lane_beam[l], 0, bytes, c->stream));
  }
  return SVO_OK;
}

// The conservative beam pre-pass in two halves, for the tile partition: every rank traces 1/N of the lattice and stores it
// into every rank's lattice buffer over NVLink (dst_ptrs: svo_device_ptr / svo_ipc_import of SVO_PLANE_BEAM_LATTICE of the
// frame's lane), the launch's last CTA bumps fence_ptrs[*] + slot; after waiting for N such bumps every rank filters locally.
int svo_beam_lattice_rows(svo_ctx *c, const svo_frame *frame, int row0, int row1, void *const *dst_ptrs, int ndst, void *const *fence_ptrs, int nsig,
                          int slot) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  if (!c->have_scene) return fail(c, SVO_ERR_NO_SCENE, "svo_beam_lattice_rows before svo_upload");
  int rc = check_frame(c, frame);
  if (rc) return rc;
  if (ndst < 0 || ndst > 16 || nsig < 0 || nsig > 16 || (ndst > 0 && !dst_ptrs) || (nsig > 0 && !fence_ptrs)) return fail(c, SVO_ERR_INVALID, "bad pointer list (at most 16)");
  if (slot < 0 || slot >= 16) return fail(c, SVO_ERR_INVALID, "fence slot must be in [0,16)");
  SVO_CUDA(c, cudaSetDevice(c->device));
  const int l = c->render_set;
  if ((rc = ensure_lattice(c, l)) != SVO_OK) return rc;
  FenceList dst, sig;
  dst.n = ndst;
  sig.n = nsig;
  for (int i = 0; i < 16; i++) {
    dst.p[i] = i < ndst ? (unsigned int *)dst_ptrs[i] : nullptr;
    sig.p[i] = i < nsig ? (unsigned int *)fence_ptrs[i] + 8 * slot : nullptr;
  }
  FrameParams fp;
  memcpy(&fp, frame, sizeof fp);
  SVO_CUDA(c, launch_beam_lattice_rows(scene_view(c), fp, c->d_beam_lattice[l], c->W, c->H, row0, row1, dst, sig, c->d_tile_counter + 32 + 2 * l, c->stream));
  c->launches++;
  return SVO_OK;
}
int svo_beam_filter(svo_ctx *c) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  SVO_CUDA(c, cudaSetDevice(c->device));
  const int l = c->render_set;
  int rc = ensure_lattice(c, l);
  if (rc) return rc;
  SVO_CUDA(c, launch_beam_filter(c->d_beam_lattice[l], (float *)plane_ptr(c, SVO_PLANE_BEAM), c->W, c->H, c->stream));
  c->launches++;
  return SVO_OK;
}

int svo_beam_conservative(svo_ctx *c, const svo_frame *frame) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  if (!c->have_scene) return fail(c, SVO_ERR_NO_SCENE, "svo_beam_conservative before svo_upload");
  int rc = check_frame(c, frame);
  if (rc) return rc;
  SVO_CUDA(c, cudaSetDevice(c->device));
  // every lane has its own lattice scratch and beam plane: the pre-pass of frame k+1 may run while frame k is still being drawn
  const int l = c->render_set;
  if ((rc = ensure_lattice(c, l)) != SVO_OK) return rc;
  FrameParams fp;
  memcpy(&fp, frame, sizeof fp);
  SVO_CUDA(c, launch_beam_conservative(scene_view(c), fp, c->d_beam_lattice[l], (float *)plane_ptr(c, SVO_PLANE_BEAM), c->W, c->H, c->stream));
  c->launches += 2;
  return SVO_OK;
}

int svo_sync(svo_ctx *c) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  SVO_CUDA(c, cudaSetDevice(c->device));
  SVO_CUDA(c, sync_lanes(c));
  if (c->fence_waits_unchecked) {  // did a frame-complete wait give up (k_fence_wait's watchdog)?  Frames since then are incomplete.
    unsigned int dead = 0;
    SVO_CUDA(c, cudaMemcpy(&dead, c->d_fence + 7, sizeof dead, cudaMemcpyDeviceToHost));
    c->fence_waits_unchecked = false;
    if (dead == 0xDEADu)
      return fail(c, SVO_ERR_FENCE, "a frame-complete fence timed out: a peer GPU fell behind or died; frames rendered since are incomplete "
                                    "(svo_fence_reset clears the latch)");
  }
  return SVO_OK;
}

int svo_read_plane_rows(svo_ctx *c, int plane, int y0, int y1, void *dst, uint64_t dst_bytes) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  if (plane < 0 || plane > SVO_PLANE_RADIANCE || !dst) return fail(c, SVO_ERR_INVALID, "bad plane or NULL dst");
  const int rows = plane == SVO_PLANE_BEAM ? c->H / 4 : c->H, cols = plane == SVO_PLANE_BEAM ? c->W / 4 : c->W;
  if (y0 < 0 || y1 > rows || y0 > y1) return fail(c, SVO_ERR_INVALID, "row range outside the plane");
  const void *src = plane_ptr(c, plane);
  if (!src) return fail(c, SVO_ERR_INVALID, "plane not allocated (enable SVO_OPT_AUX_PLANES before rendering)");
  const size_t eb = plane_elem_bytes(plane);
  const size_t bytes = (size_t)(y1 - y0) * (size_t)cols * eb;
  if (dst_bytes < bytes) return fail(c, SVO_ERR_INVALID, "dst too small");
  SVO_CUDA(c, cudaSetDevice(c->device));
  if (bytes) SVO_CUDA(c, cudaMemcpyAsync(dst, (const char *)src + (size_t)y0 * (size_t)cols * eb, bytes, cudaMemcpyDeviceToHost, c->stream));
  SVO_CUDA(c, cudaStreamSynchronize(c->stream));
  return SVO_OK;
}
int svo_read_plane(svo_ctx *c, int plane, void *dst, uint64_t dst_bytes) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  return svo_read_plane_rows(c, plane, 0, plane == SVO_PLANE_BEAM ? c->H / 4 : c->H, dst, dst_bytes);
}
int svo_read_color_rgba8(svo_ctx *c, uint8_t *dst) { return svo_read_plane(c, SVO_PLANE_COLOR_RGBA8, dst, ~0ull); }
int svo_read_depth(svo_ctx *c, float *dst) { return svo_read_plane(c, SVO_PLANE_DEPTH, dst, ~0ull); }
int svo_read_hit_id(svo_ctx *c, uint32_t *dst) { return svo_read_plane(c, SVO_PLANE_HIT_ID, dst, ~0ull); }
int svo_read_iter(svo_ctx *c, uint32_t *dst) { return svo_read_plane(c, SVO_PLANE_ITER, dst, ~0ull); }
int svo_read_primary_t(svo_ctx *c, float *dst) { return svo_read_plane(c, SVO_PLANE_PRIMARY_T, dst, ~0ull); }
int svo_read_radiance_f32(svo_ctx *c, float *dst) { return svo_read_plane(c, SVO_PLANE_RADIANCE, dst, ~0ull); }

int svo_read_depth_at(svo_ctx *c, int x, int y, float *dst) {
  if (!c || !dst) return fail(nullptr, SVO_ERR_INVALID, "NULL argument");
  if (x < 0 || y < 0 || x >= c->W || y >= c->H) return fail(c, SVO_ERR_INVALID, "pixel outside the image");
  SVO_CUDA(c, cudaSetDevice(c->device));
  const float *src = (const float *)plane_ptr(c, SVO_PLANE_DEPTH) + (size_t)y * (size_t)c->W + (size_t)x;
  SVO_CUDA(c, cudaMemcpyAsync(dst, src, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  SVO_CUDA(c, cudaStreamSynchronize(c->stream));
  return SVO_OK;
}

int svo_read_planes_async(svo_ctx *c, uint8_t *rgba8_dst, float *depth_dst) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  SVO_CUDA(c, cudaSetDevice(c->device));
  int rc = ensure_pipeline(c);
  if (rc) return rc;
  SVO_CUDA(c, cudaEventRecord(c->ev_rendered, c->stream));        // everything rendered so far
  SVO_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_rendered, 0));
  const size_t n = (size_t)c->W * (size_t)c->H;
  if (rgba8_dst) SVO_CUDA(c, cudaMemcpyAsync(rgba8_dst, plane_ptr(c, SVO_PLANE_COLOR_RGBA8), n * 4, cudaMemcpyDeviceToHost, c->copy_stream));
  if (depth_dst) SVO_CUDA(c, cudaMemcpyAsync(depth_dst, plane_ptr(c, SVO_PLANE_DEPTH), n * 4, cudaMemcpyDeviceToHost, c->copy_stream));
  SVO_CUDA(c, cudaEventRecord(c->ev_copied[c->render_set], c->copy_stream));
  return SVO_OK;
}

// The rows of bands part, part + parts, ... (what svo_render_interleaved(part, parts) draws) from the current lane's
// colour / depth set into FULL-FRAME host buffers, at their place in the frame: one strided copy per plane.
int svo_read_interleaved_async(svo_ctx *c, int part, int parts, uint8_t *rgba8_frame, float *depth_frame) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  if (parts < 1 || part < 0 || part >= parts) return fail(c, SVO_ERR_INVALID, "part must be in [0, parts)");
  SVO_CUDA(c, cudaSetDevice(c->device));
  int rc = ensure_pipeline(c, 0);
  if (rc) return rc;
  SVO_CUDA(c, cudaEventRecord(c->ev_rendered, c->stream));
  SVO_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_rendered, 0));
  const int band = c->opt_band_rows, nbands = (c->H + band - 1) / band;
  const int mine = nbands > part ? (nbands - part + parts - 1) / parts : 0;  // bands part, part + parts, ...
  for (int p = 0; p < 2; p++) {
    char *dst = p == 0 ? (char *)rgba8_frame : (char *)depth_frame;
    if (!dst || mine == 0) continue;
    const char *src = (const char *)plane_ptr(c, p);
    const size_t row = (size_t)c->W * 4u, band_bytes = row * (size_t)band, pitch = band_bytes * (size_t)parts, off = band_bytes * (size_t)part;
    // all of this part's bands but the last are whole; the last one may be cut by the image's bottom edge
    const int last = part + (mine - 1) * parts;
    const int last_rows = (last + 1) * band <= c->H ? band : c->H - last * band;
    const int whole = last_rows == band ? mine : mine - 1;
    if (whole > 0) SVO_CUDA(c, cudaMemcpy2DAsync(dst + off, pitch, src + off, pitch, band_bytes, (size_t)whole, cudaMemcpyDeviceToHost, c->copy_stream));
    if (whole < mine) {
      const size_t o = band_bytes * (size_t)last;
      SVO_CUDA(c, cudaMemcpyAsync(dst + o, src + o, row * (size_t)last_rows, cudaMemcpyDeviceToHost, c->copy_stream));
    }
  }
  SVO_CUDA(c, cudaEventRecord(c->ev_copied[c->render_set], c->copy_stream));
  return SVO_OK;
}

int svo_swap_buffers(svo_ctx *c) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  SVO_CUDA(c, cudaSetDevice(c->device));
  int rc = ensure_pipeline(c);
  if (rc) return rc;
  c->render_set = c->render_set == 0 ? 1 : 0;  // (lanes 2, 3 are reached with svo_select_lane only)
  c->lane = c->render_set;  // the other set is drawn on the other lane: consecutive frames may overlap on the GPU
  refresh_stream(c);
  // the next render overwrites this set: wait (on the device) until its last read-back has left it
  SVO_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copied[c->render_set], 0));
  return SVO_OK;
}

int svo_select_lane(svo_ctx *c, int lane) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  if (lane < 0 || lane >= kLanes) return fail(c, SVO_ERR_INVALID, "lane must be in [0, 7)");
  if (lane == c->lane) return SVO_OK;
  SVO_CUDA(c, cudaSetDevice(c->device));
  int rc = ensure_pipeline(c, lane);
  if (rc) return rc;
  c->lane = lane;
  c->render_set = lane;
  refresh_stream(c);
  SVO_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copied[c->render_set], 0));
  return SVO_OK;
}

int svo_read_wait(svo_ctx *c) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  SVO_CUDA(c, cudaSetDevice(c->device));
  if (c->copy_stream) SVO_CUDA(c, cudaStreamSynchronize(c->copy_stream));
  return SVO_OK;
}

void *svo_device_ptr(svo_ctx *c, int plane) {
  if (c && (plane & 0xFF) == SVO_PLANE_BEAM_LATTICE && (plane >> 8) >= 0 && (plane >> 8) < kLanes) {
    cudaSetDevice(c->device);
    return ensure_lattice(c, plane >> 8) == SVO_OK ? (void *)c->d_beam_lattice[plane >> 8] : nullptr;
  }
  if (c && (plane >> 8) >= 1 && (plane >> 8) < kLanes && (plane & 0xFF) <= SVO_PLANE_DEPTH) {  // plane | (set << 8): the sets of lanes 1..3
    cudaSetDevice(c->device);
    return ensure_pipeline(c, plane >> 8) == SVO_OK ? c->back[plane >> 8][plane & 0xFF] : nullptr;
  }
  if (!c || plane < 0 || plane > SVO_PLANE_RADIANCE) return nullptr;
  if (plane <= SVO_PLANE_DEPTH) return c->own[plane];  // the front set, whatever is bound or current
  if (plane >= SVO_PLANE_HIT_ID && !c->own[plane] && !c->bound[plane]) {
    cudaSetDevice(c->device);
    if (ensure_aux(c) != SVO_OK) return nullptr;
  }
  return plane_ptr(c, plane);
}

int svo_bind_plane(svo_ctx *c, int plane, void *device_ptr) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  if (plane < 0 || plane > SVO_PLANE_RADIANCE) return fail(c, SVO_ERR_INVALID, "bad plane");
  c->bound[plane] = device_ptr;
  return SVO_OK;
}

int svo_ipc_export(svo_ctx *c, int plane, uint8_t handle[72]) {
  if (!c || !handle) return fail(nullptr, SVO_ERR_INVALID, "NULL argument");
  if ((plane & 0xFF) == SVO_PLANE_BEAM_LATTICE && (plane >> 8) >= 0 && (plane >> 8) < kLanes) {
    SVO_CUDA(c, cudaSetDevice(c->device));
    int rc = ensure_lattice(c, plane >> 8);
    if (rc) return rc;
    SVO_CUDA(c, cudaStreamSynchronize(c->stream));
    return export_handle(c, c->d_beam_lattice[plane >> 8], handle);
  }
  if ((plane >> 8) >= 1 && (plane >> 8) < kLanes && (plane & 0xFF) <= SVO_PLANE_DEPTH) {
    SVO_CUDA(c, cudaSetDevice(c->device));
    int rc = ensure_pipeline(c, plane >> 8);
    if (rc) return rc;
    return export_handle(c, c->back[plane >> 8][plane & 0xFF], handle);
  }
  if (plane < 0 || plane > SVO_PLANE_RADIANCE) return fail(c, SVO_ERR_INVALID, "bad plane");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  SVO_CUDA(c, cudaSetDevice(c->device));
  if (plane >= SVO_PLANE_HIT_ID) {
    int rc = ensure_aux(c);
    if (rc) return rc;
  }
  if (!c->own[plane]) return fail(c, SVO_ERR_INVALID, "plane not allocated");
  return export_handle(c, c->own[plane], handle);
}
int svo_fence_export(svo_ctx *c, uint8_t handle[72]) {
  if (!c || !handle) return fail(nullptr, SVO_ERR_INVALID, "NULL argument");
  SVO_CUDA(c, cudaSetDevice(c->device));
  return export_handle(c, c->d_fence, handle);
}
void *svo_fence_device_ptr(svo_ctx *c) { return c ? (void *)c->d_fence : nullptr; }
int svo_fence_signal(svo_ctx *c, void *const *fence_ptrs, int n, int slot) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  if (n < 0 || n > 16 || (n > 0 && !fence_ptrs)) return fail(c, SVO_ERR_INVALID, "bad fence list (at most 16)");
  if (slot < 0 || slot >= 16) return fail(c, SVO_ERR_INVALID, "fence slot must be in [0,16)");
  SVO_CUDA(c, cudaSetDevice(c->device));
  FenceList fl;
  fl.n = n > 0 ? n : 1;
  for (int i = 0; i < 16; i++) fl.p[i] = nullptr;
  if (n == 0) fl.p[0] = c->d_fence + 8 * slot;
  for (int i = 0; i < n; i++) fl.p[i] = (unsigned int *)fence_ptrs[i] + 8 * slot;  // slots are 32 bytes apart
  SVO_CUDA(c, launch_fence_signal(fl, c->stream));
  c->launches++;
  return SVO_OK;
}
int svo_fence_wait(svo_ctx *c, int slot, uint32_t target) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  if (slot < 0 || slot >= 16) return fail(c, SVO_ERR_INVALID, "fence slot must be in [0,16)");
  SVO_CUDA(c, cudaSetDevice(c->device));
  SVO_CUDA(c, launch_fence_wait(c->d_fence + 8 * slot, c->d_fence + 7, target, c->stream));
  c->fence_waits_unchecked = true;
  c->launches++;
  return SVO_OK;
}
int svo_fence_wait_signal(svo_ctx *c, int slot, uint32_t target, void *const *fence_ptrs, int n, int signal_slot) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  if (slot < 0 || slot >= 16 || signal_slot < 0 || signal_slot >= 16) return fail(c, SVO_ERR_INVALID, "fence slot must be in [0,16)");
  if (n < 1 || n > 16 || !fence_ptrs) return fail(c, SVO_ERR_INVALID, "bad fence list (1..16)");
  SVO_CUDA(c, cudaSetDevice(c->device));
  FenceList fl;
  fl.n = n;
  for (int i = 0; i < 16; i++) fl.p[i] = i < n ? (unsigned int *)fence_ptrs[i] + 8 * signal_slot : nullptr;
  SVO_CUDA(c, launch_fence_wait_signal(c->d_fence + 8 * slot, c->d_fence + 7, target, fl, c->stream));
  c->fence_waits_unchecked = true;
  c->launches++;
  return SVO_OK;
}
int svo_fence_reset(svo_ctx *c) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  SVO_CUDA(c, cudaSetDevice(c->device));
  SVO_CUDA(c, sync_lanes(c));
  SVO_CUDA(c, cudaMemsetAsync(c->d_fence, 0, 512, c->stream));
  SVO_CUDA(c, cudaStreamSynchronize(c->stream));
  return SVO_OK;
}

int svo_ipc_import(svo_ctx *c, const uint8_t handle[72], void **device_ptr) {
  if (!c || !handle || !device_ptr) return fail(nullptr, SVO_ERR_INVALID, "NULL argument");
  SVO_CUDA(c, cudaSetDevice(c->device));
  cudaIpcMemHandle_t hnd;
  memcpy(&hnd, handle, 64);
  uint64_t off = 0;
  memcpy(&off, handle + 64, 8);
  // one block may be imported several times (planes and fences can share it): open it once per process
  for (auto &m : c->ipc_maps)
    if (memcmp(&m.handle, &hnd, 64) == 0) {
      *device_ptr = (char *)m.base + off;
      return SVO_OK;
    }
  void *base = nullptr;
  SVO_CUDA(c, cudaIpcOpenMemHandle(&base, hnd, cudaIpcMemLazyEnablePeerAccess));
  svo_ctx::IpcMap m;
  m.handle = hnd;
  m.base = base;
  c->ipc_maps.push_back(m);
  *device_ptr = (char *)base + off;
  return SVO_OK;
}
int svo_ipc_close(svo_ctx *c, void *device_ptr) {
  // unmaps every block this context imported (device_ptr is accepted for symmetry with svo_ipc_import)
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  (void)device_ptr;
  SVO_CUDA(c, cudaSetDevice(c->device));
  SVO_CUDA(c, sync_lanes(c));
  for (auto &m : c->ipc_maps) cudaIpcCloseMemHandle(m.base);
  c->ipc_maps.clear();
  cudaGetLastError();
  return SVO_OK;
}

int svo_cast_device(svo_ctx *c, const void *d_rays, uint64_t n, void *d_out, int maxDepth) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  if (!c->have_scene) return fail(c, SVO_ERR_NO_SCENE, "svo_cast before svo_upload");
  if (maxDepth < 1 || maxDepth > 23) return fail(c, SVO_ERR_INVALID, "maxDepth must be in [1,23]");
  if (n && (!d_rays || !d_out)) return fail(c, SVO_ERR_INVALID, "NULL ray or hit buffer");
  SVO_CUDA(c, cudaSetDevice(c->device));
  SVO_CUDA(c, join_lanes(c));  // sort scratch and the persistent stream kernel's counter exist once
  const uint32_t *order = nullptr;
  if (c->opt_sort && n >= 65536 && n < (1ull << 31)) {  // bin by octant + origin Morton code; results go back to the caller's order
    if (n > c->sort_cap) {
      SVO_CUDA(c, cudaStreamSynchronize(c->stream));
      if (c->d_sort) cudaFree(c->d_sort);
      if (c->d_sort_temp) cudaFree(c->d_sort_temp);
      c->d_sort = nullptr;
      c->d_sort_temp = nullptr;
      c->sort_cap = 0;
      c->sort_temp_bytes = ray_sort_temp_bytes(n);
      SVO_CUDA(c, cudaMalloc((void **)&c->d_sort, 4 * n * sizeof(uint32_t)));
      SVO_CUDA(c, cudaMalloc(&c->d_sort_temp, c->sort_temp_bytes ? c->sort_temp_bytes : 16));
      c->sort_cap = n;
    }
    uint32_t *keys = c->d_sort, *keys_alt = keys + c->sort_cap, *idx = keys_alt + c->sort_cap, *ord = idx + c->sort_cap;
    SVO_CUDA(c, launch_ray_sort(d_rays, n, keys, keys_alt, idx, ord, c->d_sort_temp, c->sort_temp_bytes, c->opt_sort, c->stream));
    c->launches += 2;
    order = ord;
  }
  SVO_CUDA(c, launch_cast(launch_cfg(c), scene_view(c), d_rays, order, n, d_out, maxDepth, c->stream));
  if (n) c->launches++;
  return SVO_OK;
}

int svo_cast(svo_ctx *c, const svo_ray *rays, uint64_t n, svo_hit *out, int maxDepth) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  if (!c->have_scene) return fail(c, SVO_ERR_NO_SCENE, "svo_cast before svo_upload");
  if (n == 0) return SVO_OK;
  if (!rays || !out) return fail(c, SVO_ERR_INVALID, "NULL ray or hit buffer");
  SVO_CUDA(c, cudaSetDevice(c->device));
  if (n > c->cast_cap) {
    SVO_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->d_rays) cudaFree(c->d_rays);
    if (c->d_hits) cudaFree(c->d_hits);
    c->d_rays = c->d_hits = nullptr;
    c->cast_cap = 0;
    SVO_CUDA(c, cudaMalloc(&c->d_rays, n * sizeof(svo_ray)));
    SVO_CUDA(c, cudaMalloc(&c->d_hits, n * sizeof(svo_hit)));
    c->cast_cap = n;
  }
  SVO_CUDA(c, cudaMemcpyAsync(c->d_rays, rays, n * sizeof(svo_ray), cudaMemcpyHostToDevice, c->stream));
  int rc = svo_cast_device(c, c->d_rays, n, c->d_hits, maxDepth);
  if (rc) return rc;
  SVO_CUDA(c, cudaMemcpyAsync(out, c->d_hits, n * sizeof(svo_hit), cudaMemcpyDeviceToHost, c->stream));
  SVO_CUDA(c, cudaStreamSynchronize(c->stream));
  return SVO_OK;
}

int svo_timer_begin(svo_ctx *c) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  SVO_CUDA(c, cudaSetDevice(c->device));
  SVO_CUDA(c, join_lanes(c));  // the interval starts after everything enqueued so far, on either lane ...
  SVO_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  // ... and nothing enqueued on the other lane from now on may start before it
  for (int l = 0; l < kLanes; l++)
    if (c->lane_stream[l] != c->stream) SVO_CUDA(c, cudaStreamWaitEvent(c->lane_stream[l], c->ev0, 0));
  return SVO_OK;
}
int svo_timer_end(svo_ctx *c, float *ms) {
  if (!c || !ms) return fail(nullptr, SVO_ERR_INVALID, "NULL argument");
  SVO_CUDA(c, cudaSetDevice(c->device));
  SVO_CUDA(c, join_lanes(c));  // ... and ends when both lanes have drained
  SVO_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  SVO_CUDA(c, cudaEventSynchronize(c->ev1));
  SVO_CUDA(c, cudaEventElapsedTime(ms, c->ev0, c->ev1));
  return SVO_OK;
}
int svo_launch_count(const svo_ctx *c, uint64_t *count) {
  if (!c || !count) return fail(nullptr, SVO_ERR_INVALID, "NULL argument");
  *count = c->launches;
  return SVO_OK;
}

int svo_transcode_probe(const uint8_t *nodes, uint64_t nbytes, int nthreads, uint64_t out[8], uint32_t *desc_out, uint64_t desc_cap) {
  if ((!nodes && nbytes) || !out) return fail(nullptr, SVO_ERR_INVALID, "NULL argument");
  Transcoded t;
  std::string err;
  if (!transcode_stream(nodes, nbytes, t, err, nthreads)) return fail(nullptr, SVO_ERR_FORMAT, err);
  uint64_t h = 1469598103934665603ull;  // FNV-1a over descriptors and reference offsets
  auto mix = [&](uint32_t v) { for (int i = 0; i < 4; i++) { h ^= (v >> (8 * i)) & 0xFFu; h *= 1099511628211ull; } };
  for (size_t i = 0; i < t.desc.size(); i++) { mix(t.desc[i].x); mix(t.desc[i].y); mix(t.refbase[i]); }
  out[0] = t.desc.size();
  out[1] = t.level_start.size();
  out[2] = h;
  CellBox b = t.leaf_box;
  for (const CellBox &d : t.depth_box) b.add(d);
  out[3] = b.empty();
  out[4] = ((uint64_t)t.leaf_box.lo[0] << 32) | t.leaf_box.hi[0];
  out[5] = ((uint64_t)t.leaf_box.lo[1] << 32) | t.leaf_box.hi[1];
  out[6] = ((uint64_t)t.leaf_box.lo[2] << 32) | t.leaf_box.hi[2];
  out[7] = 0;
  for (const CellBox &d : t.depth_box) out[7] = out[7] * 31 + d.lo[1] + 7 * d.hi[1];
  if (desc_out)
    for (size_t i = 0; i < t.desc.size() && 3 * i + 2 < desc_cap; i++) {
      desc_out[3 * i] = t.desc[i].x; desc_out[3 * i + 1] = t.desc[i].y; desc_out[3 * i + 2] = t.refbase[i];
    }
  return SVO_OK;
}

int svo_scene_probe(svo_ctx *c, uint64_t out[8]) {
  if (!c || !out) return fail(nullptr, SVO_ERR_INVALID, "NULL argument");
  if (!c->have_scene) return fail(c, SVO_ERR_NO_SCENE, "svo_scene_probe before svo_upload");
  SVO_CUDA(c, cudaSetDevice(c->device));
  SVO_CUDA(c, cudaStreamSynchronize(c->stream));
  std::vector<uint2> desc(c->ndesc);
  std::vector<uint32_t> ref(c->ndesc);
  SVO_CUDA(c, cudaMemcpy(desc.data(), c->d_desc, (size_t)c->ndesc * sizeof(uint2), cudaMemcpyDeviceToHost));
  SVO_CUDA(c, cudaMemcpy(ref.data(), c->d_refbase, (size_t)c->ndesc * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  uint64_t h = 1469598103934665603ull;
  auto mix = [&](uint32_t v) { for (int i = 0; i < 4; i++) { h ^= (v >> (8 * i)) & 0xFFu; h *= 1099511628211ull; } };
  for (size_t i = 0; i < desc.size(); i++) { mix(desc[i].x); mix(desc[i].y); mix(ref[i]); }
  out[0] = c->ndesc;
  out[1] = c->nlevels;
  out[2] = h;
  CellBox b = c->leaf_box;
  for (const CellBox &d : c->depth_box) b.add(d);
  out[3] = b.empty();
  out[4] = ((uint64_t)c->leaf_box.lo[0] << 32) | c->leaf_box.hi[0];
  out[5] = ((uint64_t)c->leaf_box.lo[1] << 32) | c->leaf_box.hi[1];
  out[6] = ((uint64_t)c->leaf_box.lo[2] << 32) | c->leaf_box.hi[2];
  out[7] = 0;
  for (const CellBox &d : c->depth_box) out[7] = out[7] * 31 + d.lo[1] + 7 * d.hi[1];
  return SVO_OK;
}

static int render_stats(svo_ctx *c, const svo_frame *frame, uint64_t counters[3], bool executed) {
  if (!c || !counters) return fail(nullptr, SVO_ERR_INVALID, "NULL argument");
  if (!c->have_scene) return fail(c, SVO_ERR_NO_SCENE, "svo_render_stats before svo_upload");
  int rc = check_frame(c, frame);
  if (rc) return rc;
  SVO_CUDA(c, cudaSetDevice(c->device));
  if ((rc = ensure_aux(c)) != SVO_OK) return rc;
  SVO_CUDA(c, join_lanes(c));
  DevBuf dbuf;
  SVO_CUDA(c, dbuf.alloc(3 * sizeof(unsigned long long)));
  unsigned long long *d = dbuf.as<unsigned long long>();
  SVO_CUDA(c, cudaMemsetAsync(d, 0, 3 * sizeof(unsigned long long), c->stream));
  FrameParams fp;
  memcpy(&fp, frame, sizeof fp);
  SVO_CUDA(c, launch_render_stats(scene_view(c, executed ? frame : nullptr), fp, planes_of(c), c->W, c->H, 0, c->H, d, c->stream, executed));
  c->launches++;
  unsigned long long h[3];
  SVO_CUDA(c, cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  SVO_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 3; i++) counters[i] = h[i];
  return SVO_OK;
}

// Layout-independent fingerprint of the descriptor tree: a depth-first walk from the root over the downloaded arrays.
// Two scenes that hold the same octree give the same numbers however their descriptors are ordered (whole transcode:
// breadth-first; after svo_upload_range: patched subtrees live at the tail).
int svo_scene_canonical(svo_ctx *c, uint64_t out[4]) {
  if (!c || !out) return fail(nullptr, SVO_ERR_INVALID, "NULL argument");
  if (!c->have_scene) return fail(c, SVO_ERR_NO_SCENE, "svo_scene_canonical before svo_upload");
  SVO_CUDA(c, cudaSetDevice(c->device));
  SVO_CUDA(c, sync_lanes(c));
  std::vector<uint2> desc(c->ndesc);
  std::vector<uint32_t> ref(c->ndesc);
  SVO_CUDA(c, cudaMemcpy(desc.data(), c->d_desc, (size_t)c->ndesc * sizeof(uint2), cudaMemcpyDeviceToHost));
  SVO_CUDA(c, cudaMemcpy(ref.data(), c->d_refbase, (size_t)c->ndesc * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  uint64_t h = 1469598103934665603ull, reachable = 0, max_depth = 0;
  auto mix = [&](uint32_t v) { for (int i = 0; i < 4; i++) { h ^= (v >> (8 * i)) & 0xFFu; h *= 1099511628211ull; } };
  std::vector<std::pair<uint32_t, uint32_t>> stack;  // (descriptor index, depth)
  if (c->ndesc) stack.push_back({0u, 0u});
  while (!stack.empty()) {
    const uint32_t i = stack.back().first, d = stack.back().second;
    stack.pop_back();
    if (i >= c->ndesc || reachable > c->ndesc) return fail(c, SVO_ERR_FORMAT, "descriptor tree is broken (index out of range or cycle)");
    reachable++;
    if (d > max_depth) max_depth = d;
    mix(desc[i].y);
    mix(ref[i]);
    mix(d);
    const uint32_t has = desc[i].y >> 24;
    uint32_t k = (uint32_t)__builtin_popcount(has);
    for (int cidx = 7; cidx >= 0; cidx--)  // push in reverse so that children pop in child order
      if ((has >> cidx) & 1u) stack.push_back({desc[i].x + --k, d + 1u});
  }
  out[0] = reachable;
  out[1] = h;
  out[2] = max_depth;
  out[3] = c->ndesc;
  return SVO_OK;
}
int svo_upload_stats(const svo_ctx *c, uint64_t out[4]) {
  if (!c || !out) return fail(nullptr, SVO_ERR_INVALID, "NULL argument");
  for (int k = 0; k < 4; k++) out[k] = c->patch_stats[k];
  return SVO_OK;
}

int svo_render_stats(svo_ctx *c, const svo_frame *frame, uint64_t counters[3]) { return render_stats(c, frame, counters, false); }
int svo_render_stats_executed(svo_ctx *c, const svo_frame *frame, uint64_t counters[3]) { return render_stats(c, frame, counters, true); }

int svo_gather_probe(svo_ctx *c, uint64_t working_set_bytes, int loads_per_thread, double *sectors_per_s) {
  if (!c || !sectors_per_s) return fail(nullptr, SVO_ERR_INVALID, "NULL argument");
  if (working_set_bytes < 4096 || loads_per_thread < 8) return fail(c, SVO_ERR_INVALID, "working set or load count too small");
  SVO_CUDA(c, cudaSetDevice(c->device));
  const uint64_t words = working_set_bytes / 8;
  DevBuf b0, b1;
  SVO_CUDA(c, b0.alloc(words * 8));
  SVO_CUDA(c, b1.alloc(64));
  void *buf = b0.p;
  uint32_t *sink = b1.as<uint32_t>();
  SVO_CUDA(c, cudaMemsetAsync(buf, 1, words * 8, c->stream));
  const int loads = (loads_per_thread + 7) / 8 * 8;
  const int blocks = c->sm_count * 8;
  SVO_CUDA(c, launch_gather_probe(buf, words, loads, blocks, sink, c->stream));  // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    SVO_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    SVO_CUDA(c, launch_gather_probe(buf, words, loads, blocks, sink, c->stream));
    SVO_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    SVO_CUDA(c, cudaEventSynchronize(c->ev1));
    float ms = 0;
    SVO_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    if (ms < best) best = ms;
  }
  c->launches += 6;
  *sectors_per_s = (double)blocks * 256.0 * (double)loads / ((double)best * 1e-3);
  return SVO_OK;
}

int svo_math_probe(svo_ctx *c, int fn, const float *x, const float *y, float *out, uint64_t n) {
  if (!c) return fail(nullptr, SVO_ERR_INVALID, "ctx is NULL");
  if (n == 0) return SVO_OK;
  if (!x || !out || (fn == 4 && !y)) return fail(c, SVO_ERR_INVALID, "NULL argument");
  SVO_CUDA(c, cudaSetDevice(c->device));
  DevBuf bx, by, bo;
  SVO_CUDA(c, bx.alloc(n * 4));
  SVO_CUDA(c, by.alloc(n * 4));
  SVO_CUDA(c, bo.alloc(n * 4));
  float *dx = bx.as<float>(), *dy = by.as<float>(), *dout = bo.as<float>();
  SVO_CUDA(c, cudaMemcpyAsync(dx, x, n * 4, cudaMemcpyHostToDevice, c->stream));
  SVO_CUDA(c, cudaMemcpyAsync(dy, y ? y : x, n * 4, cudaMemcpyHostToDevice, c->stream));
  SVO_CUDA(c, launch_math_probe(fn, dx, dy, dout, n, c->stream));
  c->launches++;
  SVO_CUDA(c, cudaMemcpyAsync(out, dout, n * 4, cudaMemcpyDeviceToHost, c->stream));
  SVO_CUDA(c, cudaStreamSynchronize(c->stream));
  return SVO_OK;
}

}  // extern "C"
