// engine_main.cpp -- TEST INFRASTRUCTURE.  The reference engine's frame loop (Main.preRun / updateEarly / placeSDF,
// Main.java:54-127, 130-289, 338-353) written against include/svo_renderer.hpp, the C++ restatement of Renderer.java,
// with the method names and the call order the Java code uses.  Every frame it draws is compared, byte for byte, with the
// same frame drawn through the bare C ABI on a second context; the error behaviour of the mirror (nothing throws,
// errors queue up for printGLErrors, the reference's own message on an empty range) is exercised as well.
// Exit codes: 0 ok, 77 no CUDA device (the library has no CPU path), anything else = failure at that line.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "svo_renderer.hpp"

#define REQUIRE(cond)                                                    \
  do {                                                                   \
    if (!(cond)) {                                                       \
      std::fprintf(stderr, "line %d: %s\n", __LINE__, #cond);            \
      renderer.printGLErrors();                                          \
      return __LINE__ % 200 + 1;                                         \
    }                                                                    \
  } while (0)

namespace {
const int WINDOW_WIDTH = 640, WINDOW_HEIGHT = 360;  // Constants.java scaled down
const float pos[3] = {1.5f, 1.3f, 2.0f};
const float l1[3] = {-1.6f, -0.9f, -1.0f}, l2[3] = {-1.6f, 0.9f, -1.0f}, r1[3] = {1.6f, -0.9f, -1.0f}, r2[3] = {1.6f, 0.9f, -1.0f};

svo_frame bare_frame(int frameNumber, int renderMode, int maxDepth) {
  svo_frame f;
  std::memset(&f, 0, sizeof f);
  std::memcpy(f.camPos, pos, sizeof pos);
  std::memcpy(f.l1, l1, sizeof l1);
  std::memcpy(f.l2, l2, sizeof l2);
  std::memcpy(f.r1, r1, sizeof r1);
  std::memcpy(f.r2, r2, sizeof r2);
  f.frameNumber = frameNumber;
  f.renderMode = renderMode;
  f.maxDepth = maxDepth;
  f.casts = 2;
  f.coneDepth = 11;
  return f;
}
}  // namespace

int main() {
  svo::Renderer &renderer = svo::Renderer::getInstance();
  int ndev = 0;
  if (svo_device_count(&ndev) != SVO_OK || ndev == 0) {
    // must fail loudly and stay alive: an error in the queue, no exception, no abort
    const bool made = renderer.createImages(0, WINDOW_WIDTH, WINDOW_HEIGHT);
    std::string what;
    const int err = renderer.getError(&what);
    return (!made && err == SVO_ERR_NO_DEVICE) ? 77 : 3;
  }

  // ---- Main.preRun ------------------------------------------------------------------------------------------------
  REQUIRE(renderer.createImages(0, WINDOW_WIDTH, WINDOW_HEIGHT));
  svo::Renderer::Shader *traceShader = renderer.addShader("svotrace", "src/shaders/svotrace.comp");
  svo::Renderer::Shader *beamShader = renderer.addShader("beamShader", "src/shaders/svobeam.comp");
  REQUIRE(traceShader && beamShader && renderer.getShaderByName("svotrace") == traceShader && renderer.getShaderByName("nope") == nullptr);
  const int numGroupsX = (int)std::ceil((double)WINDOW_WIDTH / 8), numGroupsY = (int)std::ceil((double)WINDOW_HEIGHT / 8);

  enum { N = 64, LEVELS = 6 };
  static uint16_t height[N * N];
  static uint8_t mat[N * N];
  REQUIRE(svo_terrain_generate(N, 1, height, mat, 1) == SVO_OK);
  uint64_t memOffset = 0;
  REQUIRE(svo_build_terrain_device(renderer.context(), height, mat, N, 32, &memOffset) == SVO_OK);  // (octree.readBufferFromFile)
  std::vector<uint8_t> nodes(memOffset + 4096, 0);
  renderer.getSSBO(nodes.data(), memOffset);
  renderer.addSSBO(7, nodes.data(), memOffset);
  REQUIRE(renderer.getError() == SVO_OK);

  svo_ctx *bare = nullptr;  // the same calls without the mirror
  REQUIRE(svo_create(&bare, 0, WINDOW_WIDTH, WINDOW_HEIGHT) == SVO_OK);
  REQUIRE(svo_upload(bare, nodes.data(), memOffset) == SVO_OK);

  std::vector<uint8_t> rgba(WINDOW_WIDTH * WINDOW_HEIGHT * 4), rgba_bare(rgba.size());
  std::vector<float> depth(WINDOW_WIDTH * WINDOW_HEIGHT), depth_bare(depth.size());
  renderer.setMaxDepth(LEVELS);
  int frameNumber = 0;
  const int modes[5] = {2, 0, 3, 1, 2};
  for (int it = 0; it < 5; it++) {
    // ---- Main.updateEarly ---------------------------------------------------------------------------------------
    const int renderMode = modes[it];
    const float crosshairDepth = renderer.depthAt(WINDOW_WIDTH / 2, WINDOW_HEIGHT / 2);
    (void)crosshairDepth;
    renderer.useProgram(traceShader);
    renderer.uniform3fv(8, pos);
    renderer.uniform3fv(1, l1);
    renderer.uniform3fv(2, l2);
    renderer.uniform3fv(3, r1);
    renderer.uniform3fv(4, r2);
    frameNumber++;
    renderer.setUniformInteger(5, frameNumber);
    renderer.setUniformInteger(6, renderMode);
    renderer.setUniformInteger(9, (int)memOffset);
    renderer.setUniformInteger(11, 0);
    renderer.dispatchCompute(traceShader, numGroupsX, numGroupsY, 1);
    renderer.getTexImage(0, rgba.data(), rgba.size());
    renderer.getTexImage(1, depth.data(), depth.size() * sizeof(float));
    REQUIRE(renderer.getError() == SVO_OK);

    svo_frame f = bare_frame(frameNumber, renderMode, LEVELS);
    REQUIRE(svo_render(bare, &f) == SVO_OK);
    REQUIRE(svo_read_color_rgba8(bare, rgba_bare.data()) == SVO_OK && svo_read_depth(bare, depth_bare.data()) == SVO_OK);
    REQUIRE(rgba == rgba_bare);
    REQUIRE(std::memcmp(depth.data(), depth_bare.data(), depth.size() * sizeof(float)) == 0);

    if (it == 1) {
      // the beam pre-pass as upstream dispatches it (Main.java:257-266), then the conservative one: same frame bit for bit
      renderer.useProgram(beamShader);
      renderer.setUniformInteger(9, (int)memOffset);
      renderer.uniform3fv(8, pos);
      renderer.setConservativeBeam(true);
      renderer.dispatchCompute(beamShader, numGroupsX / 4, numGroupsY / 4, 1);
      renderer.useProgram(traceShader);
      renderer.setUniformInteger(11, 1);
      renderer.dispatchCompute(traceShader, numGroupsX, numGroupsY, 1);
      std::vector<uint8_t> again(rgba.size());
      renderer.getTexImage(0, again.data(), again.size());
      REQUIRE(renderer.getError() == SVO_OK);
      REQUIRE(again == rgba);
      std::vector<float> beam((WINDOW_WIDTH / 4) * (WINDOW_HEIGHT / 4));
      renderer.getTexImage(2, beam.data(), beam.size() * sizeof(float));
      REQUIRE(renderer.getError() == SVO_OK);
      // ... and the pass as upstream ships it (not a lower bound: the frame changes; it must change the same way on both sides)
      renderer.setConservativeBeam(false);
      renderer.dispatchCompute(beamShader, numGroupsX / 4, numGroupsY / 4, 1);
      renderer.dispatchCompute(traceShader, numGroupsX, numGroupsY, 1);
      renderer.getTexImage(0, again.data(), again.size());
      REQUIRE(renderer.getError() == SVO_OK);
      svo_frame fb = bare_frame(frameNumber, renderMode, LEVELS);
      fb.useBeam = 1;
      REQUIRE(svo_beam(bare, &fb) == SVO_OK && svo_render(bare, &fb) == SVO_OK);
      REQUIRE(svo_read_color_rgba8(bare, rgba_bare.data()) == SVO_OK);
      REQUIRE(again == rgba_bare);
      renderer.setUniformInteger(11, 0);
    }
    if (it == 2) {
      // ---- Main.placeSDF: octree.useSDFBrush edits the host buffer, two ranges go to the device ---------------------
      // (the edit: the first child record of the root's first child block gets another material value)
      const int64_t cp = ((int64_t)nodes[1] << 24) | (nodes[2] << 16) | (nodes[3] << 8) | nodes[4];
      REQUIRE(cp > 0 && (uint64_t)cp < memOffset);
      nodes[cp] = nodes[cp] ? (uint8_t)(nodes[cp] % 3 + 1) : 0;
      renderer.updateSSBO(7, nodes.data(), cp, cp + 7);              // [start0, end0): touched records
      renderer.updateSSBO(7, nodes.data(), (int64_t)memOffset, (int64_t)memOffset);  // [start1, end1): nothing appended -> the reference's message
      REQUIRE(renderer.getError() == SVO_OK);
      REQUIRE(svo_upload(bare, nodes.data(), memOffset) == SVO_OK);
    }
  }

  // ---- error behaviour: nothing throws, the queue fills, printGLErrors drains it --------------------------------------
  renderer.addSSBO(3, nodes.data(), memOffset);                  // wrong binding
  renderer.setUniformInteger(42, 1);                             // no such uniform
  renderer.dispatchCompute(traceShader, numGroupsX, numGroupsY, 2);
  renderer.dispatchCompute(renderer.addShader("chunkgen", "src/shaders/chunkgen.comp"), 1, 1, 1);  // one error for the shader, one for the dispatch
  renderer.updateSSBO(7, nodes.data(), 0, (int64_t)memOffset + (1ll << 40));  // past the buffer: the library refuses, the scene stays
  int queued = 0;
  std::string what;
  while (renderer.getError(&what) != SVO_OK) queued++;
  REQUIRE(queued == 6);
  renderer.dispatchCompute(traceShader, numGroupsX, numGroupsY, 1);  // still alive
  renderer.printGLErrors();
  REQUIRE(renderer.getError() == SVO_OK);

  svo_destroy(bare);
  renderer.destroy();
  std::printf("engine main ok: %d frames through svo::Renderer equal the bare C ABI\n", frameNumber);
  return 0;
}
