// svo_renderer.hpp -- the engine-side face of libsvo_b200, in C++ (header only, over include/svo_b200.h).
//
// The reference's host is Java; its whole contact with the traversal path is the `Renderer` singleton
// (src/engine/Renderer.java) plus a handful of raw GL43C calls in Main (Main.java:62-122 images and shaders,
// :132-146 depth read-back, :257-285 uniforms and dispatch, :338-353 the two updateSSBO ranges of an edit).  No JVM
// exists in the build image, so this header restates that interface in the other compiled language at hand: same
// method names, same argument meaning, same error behaviour -- calls never throw and never abort; a failed call
// queues an error that printGLErrors() drains, the way glGetError is drained at Renderer.java:160-165.  A JNI /
// Panama binding has exactly these methods to fill in.
//
//   Renderer.java                                   here
//   ------------------------------------------------------------------------------------------------
//   getInstance()                                   Renderer::getInstance()   (one per process, like the reference)
//   (Main.java:66-86: images 0/1/2 W x H)           createImages(device, width, height)
//   addShader(name, path)            :43-54         addShader: "svotrace.comp" / "svobeam.comp" select the kernels
//   useProgram(shader)               :114-116       useProgram
//   glUniform3fv / glUniform1i / setUniformInteger  uniform3fv(location, v) / setUniformInteger(location, value):
//                                    :56-58           locations as svotrace.comp:5-18 declares them
//   dispatchCompute(shader, x, y, z) :118-121       dispatchCompute: groups of 8x8 pixels (trace) or of 8x8 beam texels
//   addSSBO(7, data)                 :123-129       addSSBO(7, data, nbytes): the library copies
//   updateSSBO(7, data)              :131-134       updateSSBO(7, data, nbytes)
//   updateSSBO(7, data, start, end)  :136-146       updateSSBO(7, data, start, end): "Update SSBO error: Invalid
//                                                     parameters." on start >= end, like the reference
//   getSSBO(buffer)                  :148-150       getSSBO(buffer, capacity)
//   getShaderByName(name)            :152-158       getShaderByName
//   printGLErrors()                  :160-165       printGLErrors()
//   glGetTexImage(depth) (Main.java:139)            getTexImage(unit, dst, bytes); depthAt(x, y) for the crosshair
//
// Outside the reference's interface, because the shader's #defines became per-frame parameters: setMaxDepth,
// setCasts, setConeDepth (defaults = the shader's: 13 / 2 / 11), setConservativeBeam(true) (the pre-pass that is a
// lower bound) and context() for everything else in svo_b200.h.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <memory>
#include <string>
#include <vector>

#include "svo_b200.h"

namespace svo {

class Renderer {
 public:
  enum class Kind { Trace, Beam, Unknown };
  struct Shader {
    std::string name;
    Kind kind;
  };
  struct Error {
    int code;
    std::string what;
  };

  Renderer() = default;
  Renderer(const Renderer &) = delete;
  Renderer &operator=(const Renderer &) = delete;
  ~Renderer() { destroy(); }

  static Renderer &getInstance() {
    static Renderer instance;
    return instance;
  }

  // Main.preRun's three glTexStorage2D images (RGBA8 W x H, R32F W x H, R32F W/4 x H/4): they live in the library.
  bool createImages(int device, int width, int height) {
    destroy();
    const int rc = svo_create(&ctx_, device, width, height);
    if (rc != SVO_OK) {
      push(rc, svo_last_error(nullptr));
      ctx_ = nullptr;
      return false;
    }
    width_ = width;
    height_ = height;
    return true;
  }
  void destroy() {
    if (ctx_) svo_destroy(ctx_);
    ctx_ = nullptr;
    shaders_.clear();
    current_ = nullptr;
  }

  Shader *addShader(const std::string &name, const std::string &path) {
    const std::string base = path.substr(path.find_last_of('/') == std::string::npos ? 0 : path.find_last_of('/') + 1);
    Kind k = base == "svotrace.comp" ? Kind::Trace : base == "svobeam.comp" ? Kind::Beam : Kind::Unknown;
    if (k == Kind::Unknown) push(SVO_ERR_INVALID, "addShader: " + path + " is not on the traversal path (svotrace.comp, svobeam.comp)");
    shaders_.push_back(std::unique_ptr<Shader>(new Shader{name, k}));
    return shaders_.back().get();
  }
  Shader *getShaderByName(const std::string &name) {
    for (auto &s : shaders_)
      if (s->name == name) return s.get();
    return nullptr;
  }
  void useProgram(Shader *shader) { current_ = shader; }

  // Uniform locations of svotrace.comp:5-18 / svobeam.comp (set at Main.java:257-283).  GL keeps uniforms per program;
  // the two programs of this path are always given the same camera, so one set is kept.
  void uniform3fv(int location, const float *v) {
    float *dst = location == 8 ? frame_.camPos : location == 1 ? frame_.l1 : location == 2 ? frame_.l2 : location == 3 ? frame_.r1
                 : location == 4 ? frame_.r2 : nullptr;
    if (!dst || !v) return push(SVO_ERR_INVALID, "uniform3fv: no vec3 uniform at location " + std::to_string(location));
    std::memcpy(dst, v, 3 * sizeof(float));
  }
  void setUniformInteger(int location, int value) {
    switch (location) {
      case 5: frame_.frameNumber = value; break;
      case 6: frame_.renderMode = value; break;
      case 9: buffer_end_ = value; break;  // bufferEnd feeds nothing in the shader (svotrace.comp:17); kept for getSSBO
      case 11: frame_.useBeam = value; break;
      default: push(SVO_ERR_INVALID, "setUniformInteger: no int uniform at location " + std::to_string(location));
    }
  }
  void setMaxDepth(int d) { frame_.maxDepth = d; }    // #define MAX_DEPTH 13 (svotrace.comp)
  void setCasts(int n) { frame_.casts = n; }          // the bounce loop's trip count (svotrace.comp:443)
  void setConeDepth(int d) { frame_.coneDepth = d; }  // the cone-trace LOD cut (svotrace.comp:275-277)
  void setConservativeBeam(bool on) { conservative_beam_ = on; }

  void dispatchCompute(Shader *shader, int numGroupsX, int numGroupsY, int numGroupsZ) {
    if (!ctx_) return push(SVO_ERR_INVALID, "dispatchCompute before createImages");
    if (!shader || shader->kind == Kind::Unknown) return push(SVO_ERR_INVALID, "dispatchCompute: not a shader of the traversal path");
    if (numGroupsZ != 1 || numGroupsX <= 0 || numGroupsY <= 0) return push(SVO_ERR_INVALID, "dispatchCompute: groups must be (x, y, 1)");
    if (shader->kind == Kind::Beam) {
      // Main.java:265 dispatches numGroups/4 groups of 8x8 beam texels.  As upstream ships it that leaves the last
      // rows of the beam image unwritten (135/4 = 33 groups = 264 of 270 rows); the library always fills the image.
      const int rc = conservative_beam_ ? svo_beam_conservative(ctx_, &frame_) : svo_beam(ctx_, &frame_);
      if (rc != SVO_OK) push(rc, svo_last_error(ctx_));
      beam_is_conservative_ = conservative_beam_;
      return;
    }
    if (numGroupsX * 8 < width_) return push(SVO_ERR_INVALID, "dispatchCompute: the groups must cover the image width");
    svo_frame f = frame_;
    if (f.useBeam && beam_is_conservative_ && (f.renderMode == 0 || f.renderMode == 3)) {
      f.useBeam = 0;  // lower bounds are consumed as a floor on the primary cast, not as upstream's origin shift
      f.flags |= SVO_FRAME_BEAM_FLOOR;
    } else if (beam_is_conservative_) {
      f.useBeam = 0;
    }
    const int rows = numGroupsY * 8 < height_ ? numGroupsY * 8 : height_;
    const int rc = rows == height_ ? svo_render(ctx_, &f) : svo_render_rows(ctx_, &f, 0, rows);
    if (rc != SVO_OK) push(rc, svo_last_error(ctx_));
  }

  void addSSBO(int bindIndex, const uint8_t *data, uint64_t nbytes) {
    if (!bound(bindIndex, "addSSBO")) return;
    const int rc = svo_upload(ctx_, data, nbytes);
    if (rc != SVO_OK) push(rc, svo_last_error(ctx_));
  }
  void updateSSBO(int bindIndex, const uint8_t *data, uint64_t nbytes) { addSSBO(bindIndex, data, nbytes); }
  void updateSSBO(int bindIndex, const uint8_t *data, int64_t start, int64_t end) {
    if (start >= end) {
      std::printf("Update SSBO error: Invalid parameters.\n");
      return;
    }
    if (!bound(bindIndex, "updateSSBO")) return;
    const int rc = svo_upload_range(ctx_, data, (uint64_t)start, (uint64_t)end);
    if (rc != SVO_OK) push(rc, svo_last_error(ctx_));
  }
  void getSSBO(uint8_t *buffer, uint64_t capacity) {
    if (!ctx_) return push(SVO_ERR_INVALID, "getSSBO before createImages");
    const int rc = svo_download(ctx_, buffer, capacity);
    if (rc != SVO_OK) push(rc, svo_last_error(ctx_));
  }

  // glGetTexImage of image unit 0 (RGBA8), 1 (depth, R32F) or 2 (beam, R32F W/4 x H/4); synchronises like GL does.
  void getTexImage(int unit, void *dst, uint64_t bytes) {
    if (!ctx_) return push(SVO_ERR_INVALID, "getTexImage before createImages");
    const int plane = unit == 0 ? SVO_PLANE_COLOR_RGBA8 : unit == 1 ? SVO_PLANE_DEPTH : unit == 2 ? SVO_PLANE_BEAM : -1;
    if (plane < 0) return push(SVO_ERR_INVALID, "getTexImage: image units are 0 (colour), 1 (depth), 2 (beam)");
    const int rc = svo_read_plane(ctx_, plane, dst, bytes);
    if (rc != SVO_OK) push(rc, svo_last_error(ctx_));
  }
  float depthAt(int x, int y) {  // Main.java:144-146 reads the whole depth image for this one pixel
    float d = -1.0f;
    if (!ctx_) {
      push(SVO_ERR_INVALID, "depthAt before createImages");
      return d;
    }
    const int rc = svo_read_depth_at(ctx_, x, y, &d);
    if (rc != SVO_OK) push(rc, svo_last_error(ctx_));
    return d;
  }

  // glGetError: oldest queued error, SVO_OK when none.
  int getError(std::string *what = nullptr) {
    if (errors_.empty()) return SVO_OK;
    const Error e = errors_.front();
    errors_.pop_front();
    if (what) *what = e.what;
    return e.code;
  }
  void printGLErrors() {
    std::string what;
    for (int err; (err = getError(&what)) != SVO_OK;) std::printf("SVO ERR: %d (%s)\n", err, what.c_str());
  }

  svo_ctx *context() { return ctx_; }
  const svo_frame &uniforms() const { return frame_; }
  int width() const { return width_; }
  int height() const { return height_; }

 private:
  bool bound(int bindIndex, const char *who) {
    if (!ctx_) {
      push(SVO_ERR_INVALID, std::string(who) + " before createImages");
      return false;
    }
    if (bindIndex != 7) {  // layout(std430, binding = 7) buffer shaderStorage (svotrace.comp:13)
      push(SVO_ERR_INVALID, std::string(who) + ": the node buffer is binding 7");
      return false;
    }
    return true;
  }
  void push(int code, const std::string &what) {
    if (errors_.size() < 64) errors_.push_back(Error{code, what});
  }
  void push(int code, const char *what) { push(code, std::string(what ? what : "")); }

  static svo_frame default_frame() {
    svo_frame f;
    std::memset(&f, 0, sizeof f);
    f.renderMode = 2;  // Main.java:124
    f.maxDepth = 13;
    f.casts = 2;
    f.coneDepth = 11;
    return f;
  }

  svo_ctx *ctx_ = nullptr;
  int width_ = 0, height_ = 0;
  int buffer_end_ = 0;
  bool conservative_beam_ = false, beam_is_conservative_ = false;
  svo_frame frame_ = default_frame();
  std::vector<std::unique_ptr<Shader>> shaders_;
  Shader *current_ = nullptr;
  std::deque<Error> errors_;
};

}  // namespace svo
