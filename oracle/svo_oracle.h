/*
 * svo_oracle.h -- TEST INFRASTRUCTURE.  CPU restatement ("oracle") of the
 * reference hot path, /root/reference/src/shaders/svotrace.comp, and of the
 * octree builder in /root/reference/src/engine/Octree.java.
 *
 * PINNED TO THE REFERENCE'S OWN CODE: the reference ships no golden vector,
 * known-answer test or level file for this path,
 * but its shader text compiles for the CPU: oracle/build_ref.py builds
 * oracle/_ref/libsvo_ref.so from /root/reference/src/shaders/svotrace.comp and
 * svobeam.comp (g++ through oracle/glsl_shim.h), and tests/test_oracle_ref.py
 * demands that every function below equals it bit for bit (all planes of all
 * render modes on BASELINE configs[0], single casts with every castResult
 * field, ray streams, the iteration cap, degenerate directions, the beam
 * pass).  tests/golden/svo_golden.npz is generated from that library.
 * NOT pinned: the builder (Octree.java needs a JVM; svo_builder.c is a
 * restatement checked by hand-derived known answers), and the two extensions
 * the shipped shader has only as comments (mirror material :500-504,
 * progressive mean :712-719).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product never does.
 */
#ifndef SVO_ORACLE_H
#define SVO_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVO_O_MAX_SCALE 23
#define SVO_O_MAX_RAYCAST_ITERATIONS 1500
#define SVO_O_NO_HIT 0xFFFFFFFFu

/* svotrace.comp:186-197 (castResult) */
typedef struct svo_o_cast_result {
  uint32_t value;
  uint32_t pointer;
  uint32_t iter;
  float t;
  float hitPos[3];
  float scale;
  float debugColor[3];
  float normal[3];
  float voxelPos[3];
  uint32_t depth;
} svo_o_cast_result;

/* Per-frame parameters == the uniforms set at Main.java:269-283.  Field-for-
 * field the same layout as `svo_frame` in include/svo_b200.h (checked by
 * tests/test_abi.py). */
typedef struct svo_o_frame {
  float camPos[3];
  float l1[3], l2[3], r1[3], r2[3];
  int32_t frameNumber;
  int32_t renderMode;
  int32_t useBeam;
  int32_t maxDepth;    /* MAX_DEPTH, reference 13 (svotrace.comp:40) */
  int32_t casts;       /* mode-0 loop count, reference 2 (svotrace.comp:444) */
  int32_t coneDepth;   /* sticky LOD cut, reference 11 (svotrace.comp:275-277) */
  int32_t mirrorValue; /* 0 = as shipped; else voxel value that reflects (svotrace.comp:500-504, commented out upstream) */
  int32_t flags;       /* bit 0: progressive running mean (svotrace.comp:712-719, commented out upstream) */
} svo_o_frame;

typedef struct svo_o_stats {
  uint64_t casts;          /* intersectOctree calls */
  uint64_t iters;          /* loop iterations == child records fetched */
  uint64_t record_bytes;   /* 7 (root) per cast + sum of fetched child-record sizes (7/3/1) */
  uint64_t stale_pops;     /* POPs that read a stack slot not written during the same cast */
  uint64_t capped;         /* casts that hit MAX_RAYCAST_ITERATIONS */
  uint64_t hits;
} svo_o_stats;

/* ray-stream records (new API, SURVEY 8b "ray-stream") */
typedef struct svo_o_ray { float o[3]; float d[3]; } svo_o_ray;
typedef struct svo_o_hit { uint32_t id; float t; uint32_t value; uint32_t iter; } svo_o_hit;

/* One intersectOctree call with a fresh (zeroed) stack.  `res` is in/out
 * (fields not written by the call keep their previous content, as `res`
 * does in the shader).  Returns the shader's bool. */
int svo_oracle_cast(const uint8_t *nodes, uint64_t nbytes, const float o[3], const float d[3],
                    int maxDepth, int coneTrace, int coneDepth, svo_o_cast_result *res,
                    svo_o_stats *stats);

/* n independent casts (coneTrace=false); id = NO_HIT on miss. */
void svo_oracle_cast_rays(const uint8_t *nodes, uint64_t nbytes, const svo_o_ray *rays, uint64_t n,
                          int maxDepth, svo_o_hit *out, int nthreads, svo_o_stats *stats);

/* svotrace.comp main() for rows [y0,y1) of a width x height image.  All
 * output planes are full-size (width*height) and may be NULL.
 *   rgba8     : framebufferImage  (clamp, *255, round-half-up)
 *   depth     : depthbufferImage  (the shader's `depth` out value)
 *   radiance  : finalcolor before quantisation, 4 floats/pixel (a = 1)
 *   hit_id    : primary cast res.pointer, NO_HIT on miss
 *   iter      : primary cast loop iterations (hit or miss)
 *   primary_t : primary cast t_min at exit on hit, 0 on miss
 *   beam      : optional (width/4)x(height/4) beam distances when frame.useBeam
 */
void svo_oracle_render(const uint8_t *nodes, uint64_t nbytes, const svo_o_frame *frame, int width,
                       int height, int y0, int y1, const float *beam, uint8_t *rgba8, float *depth,
                       float *radiance, uint32_t *hit_id, uint32_t *iter, float *primary_t,
                       int nthreads, svo_o_stats *stats);

/* svobeam.comp main(): one un-normalised ray per 4x4 block; out is
 * (width/4)x(height/4) res.t (0 where the shader leaves it undefined). */
void svo_oracle_beam(const uint8_t *nodes, uint64_t nbytes, const svo_o_frame *frame, int width,
                     int height, float *beam_out, int nthreads);

/* deterministic math, exported so tests can compare it with the device copy */
float svo_oracle_sin(float x);
float svo_oracle_cos(float x);
float svo_oracle_acos(float x);
float svo_oracle_exp(float x);
float svo_oracle_rand(float x, float y);

/* ---- builder (Octree.java restatement, brute force over dense voxels) ---- */
/* Builds the node stream for an n^3 terrain (n power of two) from an n x n
 * heightmap (u16, row = z) and material map (u8).  chunk = CHUNK_SIZE
 * generalised (reference 1024).  Returns bytes written, or 0 if `cap` is too
 * small.  counts[4] = surface, non-surface, subdividable, interior. */
uint64_t svo_oracle_build_terrain(const uint16_t *height, const uint8_t *mat, int n, int chunk,
                                  uint8_t *out, uint64_t cap, uint64_t counts[4]);
/* Same, from an explicit dense voxel field of one chunk-sized cube
 * (index x | y<<lg | z<<2lg): exactly one OctreeThread's work over the whole
 * cube, dummy head as root (n <= 256 recommended). */
uint64_t svo_oracle_build_dense(const uint8_t *voxels, int n, uint8_t *out, uint64_t cap,
                                uint64_t counts[4]);

#ifdef __cplusplus
}
#endif
#endif
