// emu_scene.h -- TEST INFRASTRUCTURE: a node stream plus its transcoded device arrays, held in host memory.
#pragma once
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector>

#include "../../svo_raytracer_b200/csrc/svo_kernels.h"
#include "../../svo_raytracer_b200/csrc/svo_transcode.h"

struct emu_scene {
  std::vector<uint8_t> raw;
  svo::Transcoded t;
};

// what scene_view() of svo_capi.cu builds, over host memory (f == nullptr: every depth counts for the content box)
inline svo::SceneView emu_view_of(const emu_scene *s, const svo::FrameParams *f) {
  svo::SceneView v;
  svo::content_box(s->t.leaf_box, s->t.depth_box, f ? f->maxDepth : -1, f ? f->coneDepth : -1, v.box_lo, v.box_hi);
  v.desc = s->t.desc.data();
  v.refbase = s->t.refbase.data();
  v.raw = s->raw.data();
  v.nbytes = s->raw.size();
  v.ndesc = (uint32_t)s->t.desc.size();
  uint32_t w0 = 0;
  memcpy(&w0, s->raw.data(), std::min<size_t>(4, s->raw.size()));
  v.first_word_zero = w0 == 0u;
  v.zero = 0u;
  v.one = 1u;
  v.two = 2u;
  v.four = 4u;
  v.exp_unit = 1u << 23;
  v.top = nullptr;
  v.ntop = 0;
  return v;
}
