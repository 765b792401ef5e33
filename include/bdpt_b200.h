/* bdpt_b200.h — C-ABI of the B200-native bidirectional path tracer (libbdpt_b200.so).
 *
 * This is the drop-in boundary for the reference's BDPT hot path.  The reference
 * puts a per-pixel C++ virtual (`PathTracer::raytrace_pixel`, reference
 * src/pathtracer/pathtracer.h:69, implemented by BidirectionalPathTracer,
 * src/pathtracer/bidirection.cpp:503-542) behind `RaytracedRenderer`
 * (src/pathtracer/raytraced_renderer.cpp).  A GPU wavefront cannot sit behind a
 * per-pixel call, so the boundary is frame-granular: it replaces the role
 * RaytracedRenderer plays behind `OfflineRenderer::{set_scene, set_camera,
 * set_frame_size, render_to_file}` (src/application/renderer.h:14-99).
 *
 * Conventions: plain C, opaque handle, int status (0 = ok, <0 = error, message via
 * bdpt_last_error), no C++ types or exceptions across the boundary, one host
 * thread per handle.  Host pointers unless a name says `_device`.  Images are
 * RGB float, W*H*3, ROW 0 = BOTTOM (reference util/image.h), linear radiance.
 * There is no CPU fallback: every entry point that computes needs a CUDA device
 * and fails with BDPT_ERR_CUDA without one.
 */
#ifndef BDPT_B200_H
#define BDPT_B200_H

#include <stddef.h>
#include <stdint.h>

#include "bdpt_flatscene.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct BdptContext BdptContext;

enum BdptStatus {
  BDPT_OK = 0,
  BDPT_ERR_CUDA = -1,      /* CUDA runtime error (incl. no device) */
  BDPT_ERR_ARG = -2,       /* bad argument / call order            */
  BDPT_ERR_SCENE = -3,     /* scene rejected (unsupported light, bad indices, ...) */
  BDPT_ERR_OVERFLOW = -4   /* internal queue overflow (a bug; never expected)      */
};

enum BdptBuffer {
  BDPT_BUFFER_SUM = 0,   /* sampleBuffer  = eye + light  (bidirection.cpp:457-462,537-538) */
  BDPT_BUFFER_EYE = 1,   /* eyeBuffer     (t >= 2 strategies, bidirection.h:81)            */
  BDPT_BUFFER_LIGHT = 2  /* lightBuffer   (t == 1 splats)                                   */
};

enum BdptRenderFlags {
  BDPT_FLAG_ROULETTE = 1, /* enable the rule the reference ships commented out (bidirection.cpp:87-93) */
  BDPT_FLAG_NO_MIS = 2,   /* test hook: every strategy weighted 1, the reference's "no MIS" switch (bidirection.cpp:124-126) */
  BDPT_FLAG_BALANCE = 4   /* Veach's balance heuristic (exponent 1) instead of the power heuristic with beta = 2 that
                             multiple_importance_sampling_weight hard-codes (bidirection.cpp:121-293) */
};

typedef struct BdptRenderParams {
  int32_t ns_total;      /* -s : samples per pixel of the whole frame (the 1/ns normalisation)   */
  int32_t max_ray_depth; /* -m : PathTracer::max_ray_depth (pathtracer.h:74)                      */
  int32_t first_sample;  /* this call renders sample indices [first_sample, first_sample+n_samples) */
  int32_t n_samples;     /*   of every pixel — how a frame is sharded across GPUs                */
  uint32_t flags;        /* BdptRenderFlags                                                       */
  int32_t only_t;        /* test hook: >= 0 keeps only strategy (i_eye == only_t, i_light == only_s), */
  int32_t only_s;        /*   the author's debug filter at bidirection.cpp:492-493; -1 = all        */
  int32_t reserved;
  uint64_t seed;         /* counter-based RNG key; streams are keyed (seed, pixel, sample index)   */
} BdptRenderParams;

typedef struct BdptStats {
  uint64_t samples;          /* BDPT samples rendered since the last bdpt_clear                     */
  uint64_t rays_closest;     /* closest-hit queries (subpath walks)                                 */
  uint64_t rays_shadow;      /* shadow queries actually traced                                      */
  uint64_t connections;      /* (s,t) strategies evaluated                                          */
  uint64_t kernel_launches;  /* CUDA kernels launched by this library since bdpt_clear              */
  double ms_total;           /* device time of the last bdpt_render (CUDA events on its stream)     */
  double ms_walk;            /* ... of which: subpath walk kernel                                   */
  double ms_connect;         /* ... connection + MIS kernel                                         */
  double ms_shadow;          /* ... shadow-ray kernel                                               */
  double ms_bvh_build;       /* host BVH build of the last upload                                   */
  uint32_t bvh_nodes;        /* wide-BVH node count                                                 */
  uint32_t batch_samples;    /* samples per wavefront batch of the last render                      */
  uint32_t n_batches;        /* batches of the last render                                          */
  uint32_t timed_batches;    /* ... of which the first timed_batches carry the per-kernel times above */
} BdptStats;

/* --- lifetime ----------------------------------------------------------------------------- */
int bdpt_create(int device, BdptContext** out);
void bdpt_destroy(BdptContext* ctx);
/* Last error text of this handle (or of creation when ctx == NULL). Never NULL. */
const char* bdpt_last_error(const BdptContext* ctx);
/* Optional: run on a caller-owned CUDA stream (a cudaStream_t passed as void*), e.g. torch's. */
int bdpt_set_stream(BdptContext* ctx, void* cuda_stream);

/* --- scene / camera / frame : replaces RaytracedRenderer::set_scene (+build_accel),
 *     set_camera, set_frame_size (raytraced_renderer.cpp:105-127,135-148,158-176,350-374) --- */
int bdpt_upload_scene(BdptContext* ctx, const BdptCamera* camera, const BdptMaterial* materials, uint32_t n_materials,
                      const BdptLight* lights, uint32_t n_lights, const BdptPrim* prims, uint32_t n_prims);
int bdpt_set_camera(BdptContext* ctx, const BdptCamera* camera);
/* Lat-long radiance map of the scene's environment light (BDPT_LIGHT_ENV in the light table, the reference's `-e`
 * option: src/application/main.cpp:115-118, EnvironmentLight, src/scene/environment_light.cpp), width x height RGB
 * floats, row 0 = straight up, looked up like EnvironmentLight::sample_dir (:157-165).  The light's `radiance` scales
 * it; width = 0 removes the map (constant environment).  Independent of bdpt_upload_scene; the reference's BDPT aborts
 * on environment lights (:182-208) — see bdpt_flatscene.h for what is defined here instead. */
int bdpt_set_environment_map(BdptContext* ctx, const float* rgb, int width, int height);
/* Applies Camera::set_screen_size semantics (camera.cpp:83-89) and (re)allocates the frame buffers. */
int bdpt_set_frame(BdptContext* ctx, int width, int height);

/* --- render : replaces start_raytracing + the worker threads' raytrace_pixel loop
 *     (raytraced_renderer.cpp:267-328,595-620; bidirection.cpp:503-542) -------------------- */
/* Thin lens: Camera::lensRadius / focalDistance, the reference's `-b` / `-d` (RaytracedRenderer::set_camera,
 * raytraced_renderer.cpp:141-142).  The reference's BDPT renders through a pinhole whatever they are (it calls
 * generate_ray, bidirection.cpp:524), which is lens_radius = 0, the default.  lens_radius > 0 is an
 * extension: eye subpaths start on the lens (Camera::generate_ray_for_thin_lens, camera_lens.cpp:22-43) and t = 1
 * connections go through a sampled point of it; a light subpath that crosses the lens is the t = 0 strategy (splatted to
 * the light image), and every MIS weight gains that alternative.  Behaviour defined by oracle/bdpt_oracle.cpp
 * (splat_lens_hit, mis_weight); lens renders use the generic kernel build (BDPT_TRAIT_LENS). */
int bdpt_set_lens(BdptContext* ctx, double lens_radius, double focal_distance);
/* Cell render, the reference's `-p x y dx dy` (RaytracedRenderer::render_to_file with a cell, raytraced_renderer.cpp:
 * 330-347, tile queue over the cell :301-319): later bdpt_render calls sample only the pixels of [x, x+dx) x [y, y+dy)
 * (row 0 = bottom); their light subpaths still splat anywhere in the frame.  dx = dy = 0 restores the whole frame, and so
 * does bdpt_set_frame with a new size. */
int bdpt_set_cell(BdptContext* ctx, int x, int y, int dx, int dy);
int bdpt_clear(BdptContext* ctx);                               /* BidirectionalPathTracer::clear          */
int bdpt_render(BdptContext* ctx, const BdptRenderParams* p);   /* accumulates; returns when the GPU is done */
int bdpt_read_framebuffer(BdptContext* ctx, int which, float* rgb_out); /* D2H, already divided by ns_total  */
/* Device pointer to the un-normalised fp32 accumulators so that a multi-GPU caller can reduce them in place before
 * reading.  They are ONE allocation of 2*W*H pixels of 4 floats {r, g, b, unused}: the eye image, then the light
 * image.  which = BDPT_BUFFER_EYE / BDPT_BUFFER_LIGHT: that image (W*H*4 floats); BDPT_BUFFER_SUM: the whole
 * allocation (2*W*H*4 floats) — what a single ncclReduce / torch.distributed.reduce sums (bdpt_reduce does that).
 * These are the frame's sums as of the last finished bdpt_render: the kernels splat into per-batch accumulators that a
 * compensated (Kahan) addition folds into them after every batch, so a pixel that receives a splat from every sample of
 * the frame (the image of a point light) does not stop growing at 2^24 splats as a plain fp32 sum would. */
int bdpt_framebuffer_device(BdptContext* ctx, int which, void** device_ptr, size_t* n_floats);
/* toColor (util/image.h:194-209): exposure sqrt(2), gamma 2.2, clamp, RGBA8, row 0 = TOP as PNGs want
 * (RaytracedRenderer::save_image flips, raytraced_renderer.cpp:690-728). */
int bdpt_tonemap_rgba8(BdptContext* ctx, int which, uint8_t* rgba_out);
int bdpt_get_stats(BdptContext* ctx, BdptStats* out);

/* --- multi-GPU: the frame is sharded by sample index (scene replicated, bdpt_render's first_sample / n_samples) and
 *     ends with ONE sum-reduce of the accumulators to the root over NCCL (NVLink / NVSwitch).  The reference has no
 *     counterpart (single process; its shared sampleBuffer sits behind one mutex, bidirection.cpp:544-551).
 *     NCCL is bound at run time (libnccl.so.2); these calls fail with BDPT_ERR_CUDA where it is absent. ------------ */
#define BDPT_COMM_ID_BYTES 128
/* One process per GPU: rank 0 makes an id, ships the 128 bytes to the others by any means (torchrun's store, MPI, a
 * file), then every rank joins.  The reduce runs on the context's stream, ordered after the renders before it. */
int bdpt_comm_unique_id(void* id_out /* BDPT_COMM_ID_BYTES */);
int bdpt_comm_init_rank(BdptContext* ctx, int nranks, int rank, const void* id_bytes);
int bdpt_reduce(BdptContext* ctx, int root);
/* One process driving n contexts on n devices (pathtracer -g N): ranks are positions in ctxs. */
int bdpt_comm_init_all(BdptContext** ctxs, int n);
int bdpt_reduce_all(BdptContext** ctxs, int n, int root);
int bdpt_comm_destroy(BdptContext* ctx);

/* --- test hooks (parity against the oracle) ------------------------------------------------ */
/* Host-only (needs no GPU): builds the 8-wide BVH on the host — the binned-SAH build that BDPT_BVH=host uploads,
 * or, with BDPT_BVH_HOST_ALGO=ploc, the host restatement of the device build's algorithm — and returns it for
 * inspection — 20 words per node (layout: bidirectional-pathtracing_b200/csrc/bvh8.h), the leaf order
 * (prim_order_out[k] = primitive index of leaf slot k, n_prims entries) and the tree depth.  Pass
 * nodes_out = NULL to query n_nodes first.  Stands where BVHAccel::construct_bvh does (bvh.cpp:51-129). */
int bdpt_host_build_bvh(const BdptPrim* prims, uint32_t n_prims, uint32_t* nodes_out, uint32_t nodes_cap,
                        uint32_t* n_nodes, uint32_t* prim_order_out, uint32_t* max_depth);
/* The tree bdpt_upload_scene built for the current scene — on the device (Morton order + PLOC + the same 8-wide collapse,
 * bidirectional-pathtracing_b200/csrc/bvh_gpu.cuh), or, for scenes of at most 1,024 primitives, by the host restatement
 * of that build (the same tree; BDPT_BVH=device forces the device) — copied back for inspection: same layout as above;
 * *rounds = PLOC rounds the device build took (0 = the tree came from the host).  Pass nodes_out = NULL to query
 * n_nodes first. */
int bdpt_device_bvh(BdptContext* ctx, uint32_t* nodes_out, uint32_t nodes_cap, uint32_t* n_nodes, uint32_t* prim_order_out,
                    uint32_t* max_depth, uint32_t* rounds);
/* rays: n x {o[3], d[3], tmin, tmax} fp64 -> hits: n x {prim id (-1 miss), t, n[3], material} fp64.
 * Mirrors BVHAccel::intersect (bvh.h:104-107). */
int bdpt_intersect_batch(BdptContext* ctx, const double* rays, size_t n, double* hits);
/* Any-hit form used for shadow rays: occluded[i] = 1 if anything lies in [tmin, tmax]. */
int bdpt_occluded_batch(BdptContext* ctx, const double* rays, size_t n, int32_t* occluded);
/* in: n x {material, wo[3], wi[3]} -> out: n x {f[3], sample_pdf(0,wi), is_delta, emission[3]} (BSDF::f etc.) */
int bdpt_bsdf_eval_batch(BdptContext* ctx, const double* in, size_t n, double* out);
/* in: n x {material, wo[3], u1, u2, u_coin} -> out: n x {wi[3], pdf, f[3]} (BSDF::sample_f with explicit uniforms) */
int bdpt_bsdf_sample_batch(BdptContext* ctx, const double* in, size_t n, double* out);
/* in: n x {light, p[3], wi[3]} -> out: n x {contain_point, Le[3], point_pdf, dir_pdf} (SceneLight::sample_pdf) */
int bdpt_light_eval_batch(BdptContext* ctx, const double* in, size_t n, double* out);
/* in: n x {x, y, p[3]} -> out: n x {o[3], d[3], We, dir_pdf, raster x, raster y, dist}
 * (Camera::generate_ray / sample_ray_pdf) */
int bdpt_camera_eval_batch(BdptContext* ctx, const double* in, size_t n, double* out);
/* One BDPT sample, fully recorded.  out (capacity cap doubles) receives
 *   [n_eye, n_light, n_eye x 13 vertex records, n_light x 13 vertex records,
 *    then per evaluated strategy {t, s, contrib[3], w, raster x, raster y}]
 * with the vertex record {pos[3], n[3], alpha[3], p, material, prim, dir_pdf}; entries the CUDA path does
 * not keep (p, prim) are NaN.  Returns the number of doubles the full record needs via *n_out. */
int bdpt_trace_sample(BdptContext* ctx, const BdptRenderParams* p, int x, int y, int sample, double* out, size_t cap,
                      size_t* n_out);

#ifdef __cplusplus
}
#endif
#endif /* BDPT_B200_H */
