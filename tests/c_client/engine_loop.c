/* engine_loop.c -- TEST INFRASTRUCTURE.  A plain-C client of include/svo_b200.h that makes the call sequence of the
 * reference engine's frame loop: Main.preRun (images + addSSBO, Main.java:62-122), updateEarly (depth read-back for the
 * crosshair :132-146, uniforms :269-283, dispatchCompute :285), placeSDF (two updateSSBO ranges, :338-353).  The Java binding
 * of INTEGRATION.md cannot be compiled in this image (no JVM); this is the same boundary bound by a C compiler: it proves that
 * the header is valid C (not only C++), that every entry point the engine needs links, and -- on a GPU box -- that the sequence
 * runs.  Exit codes: 0 ok, 77 no CUDA device (the library has no CPU path), anything else = failure. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "svo_b200.h"

#define CHECK(call)                                                                              \
  do {                                                                                           \
    int rc_ = (call);                                                                            \
    if (rc_ != SVO_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, svo_last_error(ctx)); return 1; } \
  } while (0)

int main(void) {
  svo_ctx *ctx = NULL;
  int ndev = 0;
  if (svo_abi_version() != SVO_ABI_VERSION) return 2;
  if (svo_device_count(&ndev) != SVO_OK || ndev == 0) {
    int rc = svo_create(&ctx, 0, 640, 360);
    return rc == SVO_ERR_NO_DEVICE ? 77 : 3; /* must fail loudly, never fall back */
  }
  /* world: a 64^3 heightmap world (WorldGenerator.java), built on the device */
  enum { N = 64 };
  static uint16_t height[N * N];
  static uint8_t mat[N * N];
  if (svo_terrain_generate(N, 1, height, mat, 1) != SVO_OK) return 4;
  CHECK(svo_create(&ctx, 0, 640, 360));                         /* Main.preRun: images, shader */
  uint64_t mem_offset = 0;
  CHECK(svo_build_terrain_device(ctx, height, mat, N, 32, &mem_offset));
  uint8_t *nodes = (uint8_t *)malloc(mem_offset + 4096);
  if (!nodes) return 5;
  memset(nodes, 0, mem_offset + 4096);
  CHECK(svo_download(ctx, nodes, mem_offset));
  CHECK(svo_upload(ctx, nodes, mem_offset));                    /* renderer.addSSBO(7, octree.getByteBuffer()) */
  svo_frame f;
  memset(&f, 0, sizeof f);
  const float pos[3] = {1.5f, 1.3f, 2.0f}, l1[3] = {-1.6f, -0.9f, -1.0f}, l2[3] = {-1.6f, 0.9f, -1.0f}, r1[3] = {1.6f, -0.9f, -1.0f},
              r2[3] = {1.6f, 0.9f, -1.0f};                     /* Camera.java:13-18 */
  memcpy(f.camPos, pos, sizeof pos); memcpy(f.l1, l1, sizeof l1); memcpy(f.l2, l2, sizeof l2);
  memcpy(f.r1, r1, sizeof r1); memcpy(f.r2, r2, sizeof r2);
  f.renderMode = 2; f.maxDepth = 6; f.casts = 2; f.coneDepth = 11;
  float crosshair = -2.0f;
  static uint8_t rgba[640 * 360 * 4];
  for (int frame = 1; frame <= 4; frame++) {                    /* Window.run loop */
    CHECK(svo_select_lane(ctx, frame & 1));
    f.frameNumber = frame;                                      /* glUniform1i(5, ++frameNumber) */
    CHECK(svo_render(ctx, &f));                                 /* renderer.dispatchCompute(traceShader, 240, 135, 1) */
    CHECK(svo_read_depth_at(ctx, 320, 180, &crosshair));        /* glGetTexImage(depth) -> crosshair pixel */
    if (frame == 2) {                                           /* placeSDF: renderer.updateSSBO(7, buf, start, end) x 2 */
      uint64_t st[4];
      CHECK(svo_upload_range(ctx, nodes, 0, 64));               /* a range whose bytes did not change costs no transcode */
      CHECK(svo_upload_stats(ctx, st));
      if (st[0] != 0 || st[3] != 0) return 6;
      nodes[mem_offset] = 1;                                    /* a record appended at memOffset that nothing points to yet */
      CHECK(svo_upload_range(ctx, nodes, mem_offset, mem_offset + 7));
      CHECK(svo_upload_stats(ctx, st));
      if (st[3] != 0) return 8;
    }
  }
  CHECK(svo_sync(ctx));
  CHECK(svo_read_color_rgba8(ctx, rgba));
  if (!(rgba[3] == 255) || crosshair < -1.5f) return 7;
  svo_destroy(ctx);
  free(nodes);
  printf("engine loop ok: %llu node bytes, crosshair depth %.6f\n", (unsigned long long)mem_offset, crosshair);
  return 0;
}
