// kernels.cuh — the wavefront BDPT kernels.
//
// One batch of B samples goes through three kernels:
//   k_walk_pt  persistent warps, one subpath per lane as a state machine (B eye + B light subpaths): random walk
//              with closest-hit traversal and BSDF sampling; writes 64-byte path-vertex records to HBM
//              (reference prepare_bidirectional_subpath + sample_light_ray, bidirection.cpp:20-118)
//   k_connect  one block per 128 samples: the block lists the strategies (t, s) its samples actually have, sorted by
//              kind, and its threads evaluate them one strategy per thread and round: unweighted contribution, MIS
//              weight, and a shadow ray {segment, pixel, weighted contribution} appended to a queue with
//              warp-aggregated atomics (reference estimate_bidirection_radiance + multiple_importance_sampling_weight,
//              bidirection.cpp:121-469; replaces the double loop at :490-497)
//   k_shadow_pt persistent warps pull queue entries, any-hit traversal, splat the contribution if unoccluded
//              (the reference's shadow bvh->intersect, :419-433, and update_pixel, :544-551)
//
// MIS in O(1) per strategy.  The reference walks both ratio chains for every strategy (O(s+t)
// each, O(m^3) per sample).  Every factor of those chains except the one at the connection is a
// function of one subpath alone, so the walk caches per vertex j
//     D_j = pdf_{j-1}(-> j) * G(j-1, j)        (the chain's denominator at j)
//     S_j = sum_{i<=j} flag_i * prod_{l=i}^{j-1} (R_l)^2,  R_l = pdf_{l+1}(-> l) * G(l+1, l) / D_l
// and a strategy's weight is 1 / (1 + r_t^2 S_t + r_s^2 S^L_s) with r the connection ratio.
// Same numbers as the reference's loops (bidirection.cpp:133-221,224-289), including its quirks:
// wo = 0 in sample_pdf (Q5), the original light vertex in the s = 1 light chain (Q4), the
// plane-only contain_point (Q6), light_path[0].alpha = 1 for s = 0 (Q8), no t = 0 (Q9).
// With BDPT_FLAG_BALANCE the exponent 2 becomes 1 (Veach's balance heuristic; the reference ships the
// power heuristic only).
#pragma once
#include "common.cuh"
#include "rng.cuh"
#include "shading.cuh"
#include "traverse.cuh"

namespace bdpt {

struct Job {
  DScene sc;
  DCamera cam;
  // vertex records [2*(m+1)][B] x 64 bytes (four float4 per record); counts [2][B] = length | emissive-end << 8
  float4* v0;
  int* counts;
  // shadow queue: 32-byte ray records {o, tmax | dir, pixel} (q0, stride 2) and the contributions (q2), in one
  // segment of seg_cap entries per k_connect block (qcount[b] of them used): a block appends with a SHARED-memory
  // counter, so no global atomic sits between a connection and its queue slot
  // When a block is done it publishes its entries as chunks of <= 256 {first entry, count}: the unit of work of
  // k_shadow_pt's warps.
  float4 *q0, *q2;
  uint2* chunks;
  unsigned int seg_cap, chunk_cap;
  // counters: [0] chunks published  [1] shadow cursor (chunks)  [3] walk cursor (cleared per batch)  [4] queue overflow (per render)
  unsigned int* qctr;
  // accumulators (un-normalised), W*H float4 each: RGB + one unused lane, so that a splat is ONE 16-byte reduction
  float4 *acc_eye, *acc_light;
  unsigned long long* stats;  // {closest rays, shadow rays, connections}
  double* dbg;                // trace of one sample (bdpt_trace_sample) or null
  unsigned long long seed;
  unsigned long long g0;      // first job-sample index of this batch
  unsigned int k0, p0;        // g0 = k0 * npix_pad + p0 (split on the host: no 64-bit division on the device)
  unsigned long long g_end;   // one past the last job-sample index of the job
  int B;                      // samples in this batch
  int m;                      // max_ray_depth
  int first_sample;           // sample index of job-sample 0
  int tiles_x, npix_pad;      // 8x4 pixel tiles per row; padded pixel count (of the cell)
  int cx0, cy0, cx1, cy1;     // the pixels that are sampled: the frame, or the cell of a cell render
  FastDiv div_tiles_x, div_npix_pad;
  int roulette;
  int ntop;                   // BVH nodes staged in shared memory by the persistent kernels
  int no_mis;                 // weight 1 for every strategy: the reference's own debug switch (bidirection.cpp:124-126)
  int balance;                // MIS exponent 1 instead of 2
  int only_t, only_s;
  int point_lights;           // the scene has a point light: light-image splats collide (shadow_splat)
  int single_x, single_y;     // >= 0: the job is this one pixel (trace mode)
};

// job-sample index -> (pixel, sample number).  Consecutive indices walk an 8x4 pixel tile, so a
// warp's 32 primary rays are neighbours.
__device__ inline bool decode_sample(const Job& J, unsigned long long g, int* x, int* y, int* sample) {
  if (g >= J.g_end) return false;
  if (J.single_x >= 0) { *x = J.single_x; *y = J.single_y; *sample = J.first_sample + (int)g; return true; }
  unsigned int q = J.p0 + (unsigned int)(g - J.g0);   // < npix_pad + B < 2^32
  const unsigned int qd = fast_div(q, J.div_npix_pad);
  unsigned int k = J.k0 + qd;
  int p = (int)(q - qd * (unsigned int)J.npix_pad);
  int tile = p >> 5, w = p & 31;
  const int ty = (int)fast_div((unsigned int)tile, J.div_tiles_x), tx = tile - ty * J.tiles_x;
  *x = J.cx0 + tx * 8 + (w & 7);
  *y = J.cy0 + ty * 4 + (w >> 3);
  *sample = J.first_sample + (int)k;
  return *x < J.cx1 && *y < J.cy1;
}

// float4 index of the first quarter of the 64-byte record of vertex j of subpath (kind, i)
#define VSLOT(J, kind, j, i) (4 * ((size_t)((kind) * ((J).m + 1) + ((j)-1)) * (size_t)(J).B + (size_t)(i)))

// 32-byte global accesses (LDG.256 / STG.256, sm_100): half of a vertex record per instruction.  The vertex records and
// the shadow queue are written once and read once, gigabytes per batch, next to a scene (nodes, primitives) that every
// ray re-reads from L1 / L2: their accesses carry cache qualifiers, one macro per site so that each can be measured
// ("cache qualifiers").  Defaults = the measured optimum on B200:
//   vertex stores (k_walk*)      L1::no_allocate + L2::evict_first   walk -3.4 % (the records no longer push the scene out of L2)
//   vertex loads  (k_connect)    L1::no_allocate                     connect -1.8 % (with L2::evict_first: +4.4 %, the kernel's own L2 prefetch is lost)
//   queue stores  (k_connect)    L1::no_allocate + L2::evict_first   connect -0.9 %
//   queue loads   (k_shadow_pt)  L1::no_allocate
#ifndef BDPT_SHADOW_PF_NEXT
#define BDPT_SHADOW_PF_NEXT 32
#endif
#ifndef BDPT_HINT_VST
#define BDPT_HINT_VST ".L1::no_allocate.L2::evict_first"
#endif
#ifndef BDPT_HINT_VLD
#define BDPT_HINT_VLD ".L1::no_allocate"
#endif
#ifndef BDPT_HINT_QST
#define BDPT_HINT_QST ".L1::no_allocate.L2::evict_first"
#endif
#ifndef BDPT_HINT_QLD
#define BDPT_HINT_QLD ".L1::no_allocate"
#endif
#define BDPT_DEF_LD256(name, hint)                                                                                          \
  __device__ inline void name(const float4* p, float4* a, float4* b) {                                                      \
    asm volatile("ld.global" hint ".v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                                                \
                 : "=f"(a->x), "=f"(a->y), "=f"(a->z), "=f"(a->w), "=f"(b->x), "=f"(b->y), "=f"(b->z), "=f"(b->w) : "l"(p)); \
  }
#define BDPT_DEF_ST256(name, hint)                                                                                          \
  __device__ inline void name(float4* p, float4 a, float4 b) {                                                              \
    asm volatile("st.global" hint ".v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"                                                \
                 :: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");      \
  }
BDPT_DEF_LD256(ld256_vtx, BDPT_HINT_VLD)
BDPT_DEF_LD256(ld256_q, BDPT_HINT_QLD)
BDPT_DEF_ST256(st256_vtx, BDPT_HINT_VST)
BDPT_DEF_ST256(st256_q, BDPT_HINT_QST)
// the 64-byte record of a vertex: {pos, material | n, D | alpha, S | wo, q}
__device__ inline void store_vertex(const Job& J, size_t slot, float4 a, float4 b, float4 c, float4 e) {
  st256_vtx(J.v0 + slot, a, b);
  st256_vtx(J.v0 + slot + 2, c, e);
}

// update_pixel (bidirection.cpp:544-551) without the lock: one 16-byte reduction per splat (REDG.E.ADD.F32x4, sm_90+)
// into the RGBA-padded accumulator instead of three scalar ones — one L2 atomic transaction per ray.
__device__ inline void splat_rgb(float4* px, float r, float g, float b) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(px), "f"(r), "f"(g), "f"(b), "f"(0.f) : "memory");
}

__device__ inline float clampS(float s) { return fminf(s, 1e30f); }
// the MIS exponent: ratio^2 (power heuristic, what the reference ships) or ratio (balance heuristic)
__device__ inline float mis_pow(const Job& J, float r) { return J.balance ? fabsf(r) : r * r; }

// ---------------------------------------------------------------------------------------- walk
// Per-lane subpath state.
struct WalkState {
  Rng rng;
  f3 prev_n;
  f3 thr;   // alpha_{j-1} * f_{j-1} / pdf_{j-1}: alpha_j is this times the cosine at vertex j-1 (bidirection.cpp:66) — one
            // vector instead of alpha, f and pdf: four registers less across the traversal loop
  float pdfeval_prev, D_prev, S_prev;
  // subpath index i = kind * B + si (kind 0 = eye, 1 = light) | bit 30: the previous vertex is a delta surface | bit 31: the
  // last stored vertex is an emitter — and j, the vertex the current ray looks for (the subpath holds j - 1 so far): two
  // registers where kind, si, j, nv and two flags took six
  unsigned idx;
  int j;
  __device__ inline int i() const { return (int)(idx & 0x3fffffffu); }
  __device__ inline bool delta_prev() const { return (idx & 0x40000000u) != 0u; }
  __device__ inline bool emis() const { return (idx & 0x80000000u) != 0u; }
};

// New subpath: pinhole ray (raytrace_pixel, bidirection.cpp:517-524) or light emission
// (sample_light_ray, :105-118).  Returns false if the index maps to no pixel.
template <int TR = 0>
__device__ inline bool walk_begin(const Job& J, unsigned i, WalkState& W, Trav& T) {
  const bool lens = (TR & BDPT_TRAIT_LENS) && J.cam.lens_radius > 0.f;
  const int kind = i >= (unsigned)J.B ? 1 : 0;
  const int si = (int)i - kind * J.B;
  int x, y, sample;
  if (!decode_sample(J, J.g0 + (unsigned long long)si, &x, &y, &sample)) { J.counts[i] = 0; return false; }
  W.rng.begin(J.seed, (unsigned long long)y * (unsigned long long)J.cam.W + (unsigned long long)x,
              (unsigned long long)sample, (uint32_t)kind);
  W.idx = i; W.j = 2;
  f3 o, d;
  if (kind == 0) {
    float sx = W.rng.uniform(), sy = W.rng.uniform();
    if (lens) {   // thin lens (extension): two more draws, the eye subpath starts on the lens
      const float u1 = W.rng.uniform(), u2 = W.rng.uniform();
      camera_generate_ray_lens(J.cam, ((float)x + sx) / (float)J.cam.W, ((float)y + sy) / (float)J.cam.H, u1, u2, &o, &d);
    } else
    camera_generate_ray(J.cam, ((float)x + sx) / (float)J.cam.W, ((float)y + sy) / (float)J.cam.H, &o, &d);
    W.prev_n = d; W.thr = mk3(1.f);
    W.pdfeval_prev = 1.f; W.D_prev = 1.f; W.S_prev = 0.f;
    if (lens) {   // thin lens: the lens point can also be reached by a light subpath (t = 0), density 1 / lens area
      W.S_prev = 1.f; W.D_prev = 1.0f / (BDPT_PI * J.cam.lens_radius * J.cam.lens_radius);
    }
    T.begin(o, d, J.cam.nclip, J.cam.fclip);
  } else {
    int li = W.rng.pick(J.sc.n_lights);
    Light L = load_light(J.sc, li);
    float pA, pdir; f3 nl;
    f3 Le = light_sample_Le(J.sc, L, W.rng, &o, &d, &pA, &pdir, &nl);
    pA /= (float)J.sc.n_lights;
    W.prev_n = nl;
    const f3 alpha1 = Le * (1.0f / pA);
    W.thr = alpha1 * (1.0f / pdir);
    W.pdfeval_prev = pdir; W.D_prev = pA; W.S_prev = 1.f;
    size_t s1 = VSLOT(J, 1, 1, si);
    store_vertex(J, s1, make_float4(o.x, o.y, o.z, __int_as_float(-1)),
                 make_float4(nl.x, nl.y, nl.z, pA),
                 make_float4(alpha1.x, alpha1.y, alpha1.z, 1.f),
                 make_float4(pdir, 0.f, 0.f, 1.f));
    T.begin(o, d, BDPT_EPS_F, INFINITY);
  }
  return true;
}

// Thin lens: the camera vertex of an eye subpath is not stored; its lens point is re-drawn from the eye stream
// (walk_begin: two raster draws, then the two lens draws).
__device__ inline f3 lens_vertex(const Job& J, unsigned long long pix, unsigned long long sample) {
  Rng re;
  re.begin(J.seed, pix, sample, 0u);
  re.uniform(); re.uniform();
  const float u1 = re.uniform(), u2 = re.uniform();
  const f3 lc = lens_point(J.cam, u1, u2);
  return J.cam.pos + lc.x * J.cam.c2w_x + lc.y * J.cam.c2w_y;
}

// Thin lens, strategy t = 0 (extension; oracle/bdpt_oracle.cpp splat_lens_hit): the segment o + t d of a LIGHT subpath
// crosses the lens before it reaches a surface.  C = alpha at the lens * We / (lens area * cos at the lens), weighted
// against t' = 1 (the camera samples this lens point and connects) and the rest of the light-side chain; it goes to the
// light image like a t = 1 splat.  The lens is no occluder: the walk goes on to its surface hit.
__device__ inline void lens_strategy(const Job& J, const WalkState& W, f3 o, f3 d, float tmax) {
  const DCamera& c = J.cam;
  const int s = W.j - 1;
  if (J.only_t >= 0 && (J.only_t != 0 || J.only_s != s)) return;
  const f3 oc = o - c.pos;
  const f3 ocl = mk3(dot(c.w2c_r0, oc), dot(c.w2c_r1, oc), dot(c.w2c_r2, oc));
  const f3 dcl = mk3(dot(c.w2c_r0, d), dot(c.w2c_r1, d), dot(c.w2c_r2, d));
  if (!(dcl.z > 0.f)) return;                       // must arrive from the front
  const float tl = -ocl.z / dcl.z;
  if (!(tl > BDPT_EPS_F && tl < tmax)) return;
  const float lx = ocl.x + dcl.x * tl, ly = ocl.y + dcl.y * tl;
  if (lx * lx + ly * ly >= c.lens_radius * c.lens_radius) return;
  f3 eye, wi; float dist, cam_dir_pdf; int ex, ey;
  const float We = camera_sample_ray_pdf_lens(c, o, mk3(lx, ly, 0.f), &eye, &wi, &dist, &cam_dir_pdf, &ex, &ey);
  if (!(ex >= 0 && ey >= 0 && ex < c.W && ey < c.H)) return;
  const float A = BDPT_PI * c.lens_radius * c.lens_radius;
  const float cos_prev = fabsf(dot(W.prev_n, d));
  const float cos_u = dcl.z / norm(dcl);
  const f3 contrib = W.thr * cos_prev * (We / (A * cos_u));
  if (!(norm(contrib) > BDPT_EPS_F)) return;
  float w = 1.f;
  if (!J.no_mis) {
    const float g0 = cos_prev / (dist * dist);
    const float ratio1 = (1.0f / A) / (W.pdfeval_prev * g0);          // p_1 / p_0
    const float r_s = cam_dir_pdf * g0 / W.D_prev;
    const double winv = 1.0 + (double)mis_pow(J, ratio1) * ((double)((s >= 2 && W.delta_prev()) ? 0.f : 1.f) + (double)mis_pow(J, r_s) * (double)W.S_prev);
    w = (float)(1.0 / winv);
    if (!(w >= 0.f)) w = 0.f;
  }
  splat_rgb(J.acc_light + (size_t)(ey * c.W + ex), contrib.x * w, contrib.y * w, contrib.z * w);
}

// The traversal of the current ray finished: store the vertex (if any) and sample the next ray
// (body of the while loop of prepare_bidirectional_subpath, bidirection.cpp:44-101).
// Returns true if the subpath continues (T holds the next ray), false if it ended.
template <int TR = 0>
__device__ inline bool walk_shade(const Job& J, WalkState& W, Trav& T) {
  const DScene& sc = J.sc;
  const f3 o = T.o, d = T.d;
  const int kind = W.i() >= J.B ? 1 : 0, si = W.i() - kind * J.B;
  if ((TR & BDPT_TRAIT_LENS) && kind == 1 && J.cam.lens_radius > 0.f) lens_strategy(J, W, o, d, T.hit.leaf < 0 ? INFINITY : T.hit.t);
  if (T.hit.leaf < 0) {
    W.j--;   // (no vertex for this ray, unless the environment gives one below)
    if (sc.env_light < 0 || kind != 0) return false;   // (a LIGHT subpath's vertex on an emitter carries nothing: f = 0)
    // The ray leaves the scene: it meets the environment's sphere (shading.cuh, "environment light"), whose surface
    // is an emitter — the vertex is stored like any other emissive end of a subpath, and the walk ends there.
    const Light L = load_light(sc, sc.env_light);
    const f3 oc = o - L.position;
    const float b = dot(oc, d), cc = norm2(oc) - L.dim_x.x * L.dim_x.x;
    if (cc >= 0.f) return false;
    const float te = -b + sqrtf(b * b - cc);
    const f3 pe = o + d * te, ne = normalized(L.position - pe);
    const float cos_prev = fabsf(dot(W.prev_n, d));
    const float ge = cos_prev * fabsf(dot(ne, d)) / (te * te);
    const f3 alpha = W.thr * cos_prev;
    W.j++;
    const float D_e = (kind == 0 && W.j == 2) ? 1.f : W.pdfeval_prev * ge;
    float S_e = W.delta_prev() ? 0.f : 1.f;
    if (W.S_prev > 0.f) S_e += mis_pow(J, cosine_pdf(-dot(ne, d)) * ge / W.D_prev) * W.S_prev;
    store_vertex(J, VSLOT(J, kind, W.j, si), make_float4(pe.x, pe.y, pe.z, __int_as_float(sc.n_materials)),
                 make_float4(ne.x, ne.y, ne.z, D_e), make_float4(alpha.x, alpha.y, alpha.z, clampS(S_e)),
                 make_float4(-d.x, -d.y, -d.z, 1.f));
    W.idx |= 0x80000000u;
    return false;
  }
  f3 n, hit_p; int prim, mat;
  finish_hit<TR>(sc, T.hit, o, d, &n, &prim, &mat, &hit_p);
  const float t = T.hit.t;
  const Frame fr = make_frame(n);
  const f3 w_out = fr.to_local(-d);
  const Mat M = load_mat<TR>(sc, mat);
  f3 wi; float pdf;
  const f3 f = bsdf_sample_f(M, w_out, &wi, &pdf, W.rng);
  const f3 wi_world = normalized(fr.to_world(wi));

  // bidirection.cpp:64-69
  const float cos_prev = fabsf(dot(W.prev_n, d));
  const float g = cos_prev * fabsf(dot(n, d)) / (t * t);
  const f3 alpha = W.thr * cos_prev;

  // roulette (off in the reference: bidirection.cpp:80; the disabled rule is :87-93)
  float p_keep = 1.f;
  if (J.roulette && W.j > 2) p_keep = fminf(1.f, norm(f) / pdf);

  // MIS cache
  const bool delta_j = type_is_delta(M.type);
  const float D_j = (kind == 0 && W.j == 2) ? 1.f : W.pdfeval_prev * g;
  const float D_rev = bsdf_pdf_mis(M, wi, w_out) * p_keep * g;  // density of (j -> j-1) as the chains use it
  float S_j = (delta_j || W.delta_prev()) ? 0.f : 1.f;
  if (W.S_prev > 0.f) S_j += mis_pow(J, D_rev / W.D_prev) * W.S_prev;
  S_j = clampS(S_j);

  size_t sl = VSLOT(J, kind, W.j, si);
  store_vertex(J, sl, make_float4(hit_p.x, hit_p.y, hit_p.z, __int_as_float(mat)),
               make_float4(n.x, n.y, n.z, D_j),
               make_float4(alpha.x, alpha.y, alpha.z, S_j),
               make_float4(-d.x, -d.y, -d.z, p_keep));
  W.idx = (W.idx & 0x7fffffffu) | (M.type == BDPT_MAT_EMISSION ? 0x80000000u : 0u);
  if (W.j >= J.m + 1) return false;
  if (J.roulette && !(W.rng.uniform() < p_keep)) return false;
  if (is_zero(f)) return false;  // throughput is exactly zero from here on: later vertices add nothing

  W.pdfeval_prev = bsdf_pdf_mis(M, w_out, wi) * p_keep;
  W.thr = alpha * f * (1.0f / (pdf * p_keep)); W.prev_n = n;
  W.D_prev = D_j; W.S_prev = S_j;
  W.idx = (W.idx & 0xbfffffffu) | (delta_j ? 0x40000000u : 0u);
  W.j++;
  T.begin(hit_p, wi_world, BDPT_EPS_F, INFINITY);
  return true;
}

// what k_connect reads per subpath: its length and whether its LAST vertex is an emitter.  (An EmissionBSDF
// returns f = 0, bsdf.cpp:106-110, so the walk ends there: no other vertex of a subpath can be one.)
__device__ inline void walk_end(const Job& J, const WalkState& W) {
  J.counts[W.i()] = W.j | (W.emis() ? 0x100 : 0);
}

// One subpath per lane, start to end.  Used for scenes whose BVH has at most 8 nodes (nothing diverges there).
// (compiled for 7 CTAs/SM, 72 registers: 5.59 ms per 2^24 samples on CBspheres; 6 / 8: 5.67; the 95 registers the compiler
// takes without a bound — 5 CTAs —: 5.94)
#ifndef BDPT_KWALK_MINB
#define BDPT_KWALK_MINB 7
#endif
template <int TR = 0>
__global__ void __launch_bounds__(128, BDPT_KWALK_MINB) k_walk(const __grid_constant__ Job J) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;   // whole warps are in or out: 2B is a multiple of 32
  if (i >= 2u * (unsigned)J.B) return;
  WalkState W; Trav T;
  uint2 stack[BDPT_STACK];
  unsigned nrays = 0;
  if (walk_begin<TR>(J, i, W, T)) {
    for (;;) {
      while (!T.template step<false, TR>(J.sc, stack)) {}
      nrays++;
      if (!walk_shade<TR>(J, W, T)) break;
    }
    walk_end(J, W);
  }
  const unsigned tot = __reduce_add_sync(0xffffffffu, nrays);   // one atomic per warp for the ray counter
  if ((threadIdx.x & 31) == 0 && tot) atomicAdd(J.stats + 0, (unsigned long long)tot);
}

// ------------------------------------------------------------------------------------- connect
struct Vtx {
  f3 pos, n, alpha, wo;
  float D, S, q;
  int mat;
};

__device__ inline Vtx load_vtx(const Job& J, int kind, int j, int si) {
  size_t s = VSLOT(J, kind, j, si);
  float4 a, b, c, e;
  ld256_vtx(J.v0 + s, &a, &b); ld256_vtx(J.v0 + s + 2, &c, &e);
  Vtx v;
  v.pos = mk3(a.x, a.y, a.z); v.mat = __float_as_int(a.w);
  v.n = mk3(b.x, b.y, b.z); v.D = b.w;
  v.alpha = mk3(c.x, c.y, c.z); v.S = c.w;
  v.wo = mk3(e.x, e.y, e.z); v.q = e.w;
  return v;
}

#define BDPT_CONNECT_NT 128   // samples (= threads) per block of k_connect

// The strategies of a sample are the pairs (t, s), t = 1..nE eye vertices, s = 0..nL light vertices
// (bidirection.cpp:490-497).  The reference — and the first version of this kernel, one thread per (sample, t)
// looping over s — visits all of them in a fixed order; subpaths have different lengths, a third of the strategies
// need a fresh light sample (s = 1), a few need the camera (t = 1), s = 0 only matters where the eye subpath ENDS on an
// emitter, and the warp ran at 18 of 32 lanes.  Here a block takes 128 consecutive samples, counts the strategies each
// has in five kinds
//     0: (t = 1, s = 1)    1: (t >= 2, s = 1)    2: (t = 1, s >= 2)    3: (t >= 2, s >= 2)    4: (t = nE, s = 0)
// prefix-sums the counts, and then hands strategy q of the block's list (sorted by kind, then sample) to thread q mod
// 128: every lane has a strategy, a warp's strategies are of one kind (but for the warps that straddle a boundary)
// and of the same or neighbouring samples (their vertex loads coalesce), there is no loop-carried state, and
// each vertex record is fetched from DRAM by exactly one block.
template <bool TRACE, int MINB, int TR = 0>
__global__ void __launch_bounds__(BDPT_CONNECT_NT, MINB) k_connect(const __grid_constant__ Job J) {
  constexpr int NT = BDPT_CONNECT_NT, NW = NT / 32;
  // (Splitting kind 1 into (t = 2, s = 1) — rays from 128 neighbouring first hits to the light, as coherent as shadow rays
  // get, then next to each other in the queue — and (t >= 3, s = 1) was measured: shadow kernel -0.2 %, this kernel +1.2 %.)
  constexpr int NK = 5;
  __shared__ unsigned s_pre[NK][NT];    // exclusive prefix of the per-sample strategy counts, per kind
  __shared__ unsigned s_wsum[NK][NW];
  __shared__ unsigned s_pix[NT];
  __shared__ int s_smp[NT];
  __shared__ unsigned s_len[NT];       // nE | nL << 8
  __shared__ unsigned s_qn;            // entries of this block's queue segment
  const DScene& sc = J.sc;
  const bool lens_on = (TR & BDPT_TRAIT_LENS) && J.cam.lens_radius > 0.f;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int si0 = blockIdx.x * NT;
  const f3 zero = mk3(0.f);

  if (threadIdx.x == 0) s_qn = 0u;
  // ---- one thread per sample: subpath lengths, pixel, strategy counts
  unsigned cnt[NK] = {};
  unsigned nconn = 0;
  {
    const int si = si0 + tid;
    int nE = 0, nL = 0, emis = 0;
    if (si < J.B) {
      const int ce = J.counts[si];
      nE = ce & 0xff; emis = (ce >> 8) & 1;
      nL = J.counts[J.B + si] & 0xff;
    }
    int x = 0, y = 0, sample = 0;
    if (nE > 0) decode_sample(J, J.g0 + (unsigned long long)si, &x, &y, &sample);
    else nL = 0;   // padding lane of the frame's last tile row / column
    s_pix[tid] = (unsigned)(y * J.cam.W + x); s_smp[tid] = sample; s_len[tid] = (unsigned)nE | ((unsigned)nL << 8);
    if (nE > 0) {
      // every record of this sample is read by the rounds below: request them all into L2 now (connect -2.6 %)
      for (int jj = 2; jj <= nE; jj++) { const float4* pp = J.v0 + VSLOT(J, 0, jj, si); asm volatile("prefetch.global.L2 [%0];" :: "l"(pp)); }
      for (int jj = 1; jj <= nL; jj++) { const float4* pp = J.v0 + VSLOT(J, 1, jj, si); asm volatile("prefetch.global.L2 [%0];" :: "l"(pp)); }
      const unsigned e1 = (unsigned)(nE - 1), l1 = nL > 0 ? (unsigned)(nL - 1) : 0u;
      cnt[0] = nL > 0 ? 1u : 0u; cnt[1] = nL > 0 ? e1 : 0u; cnt[2] = l1; cnt[3] = e1 * l1;
      cnt[4] = TRACE ? (unsigned)nE : (unsigned)emis;   // (the trace records the null s = 0 strategies as well)
      if (J.only_t < 0) nconn = (unsigned)(nE * (nL + 1));
    }
  }
  unsigned tot[NK];
#pragma unroll
  for (int k = 0; k < NK; k++) {
    unsigned v = cnt[k];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned n = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += n; }
    if (lane == 31) s_wsum[k][wid] = v;
    cnt[k] = v - cnt[k];   // exclusive within the warp
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NK; k++) {
    unsigned off = 0, all = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) { const unsigned ws = s_wsum[k][w]; if (w < wid) off += ws; all += ws; }
    s_pre[k][tid] = off + cnt[k];
    tot[k] = all;
  }
  __syncthreads();
  unsigned bnd[NK + 1];   // first strategy of every kind in the block's list
  bnd[0] = 0u;
#pragma unroll
  for (int k = 0; k < NK; k++) bnd[k + 1] = bnd[k] + tot[k];
  const unsigned T = bnd[NK];

  // ---- one strategy per thread and round
#pragma unroll 1
  for (unsigned base = 0; base < T; base += NT) {
    const unsigned q = base + (unsigned)tid;
    bool valid = q < T;
    int cls = 0;
    unsigned r = q;
#pragma unroll
    for (int k = 1; k < NK; k++) if (q >= bnd[k]) { cls = k; r = q - bnd[k]; }
    const unsigned* pre = s_pre[cls];
    int j = 0;   // the last sample whose prefix is <= r (samples without a strategy of this kind share a prefix)
#pragma unroll
    for (int step = NT / 2; step > 0; step >>= 1) if (pre[j + step] <= r) j += step;
    const unsigned k = r - pre[j];
    const unsigned len = s_len[j];
    const int nE = (int)(len & 0xffu), nL = (int)(len >> 8);
    int t = 1, s = 1;
    if (cls == 1) t = 2 + (int)k;
    else if (cls == 2) s = 2 + (int)k;
    else if (cls == 3) {
      const int d = nL - 1;   // k < (nE-1) * d <= 64 * 64: the float quotient is exact to well under 1/(2d)
      const int tt = (int)__fdividef((float)k + 0.5f, (float)d);
      t = 2 + tt; s = 2 + ((int)k - tt * d);
    } else if (cls == 4) { s = 0; t = TRACE ? 1 + (int)k : nE; }
    if (J.only_t >= 0) { if (t != J.only_t || s != J.only_s) valid = false; else if (valid) nconn++; }
    const int sj = si0 + j;
    const unsigned pix_own = s_pix[j];

    f3 contrib = zero, ill = zero;
    float w = 0.f;
    int eye_x = -1, eye_y = -1;
    Vtx E, Lv; Mat ME, ML;
    f3 ve_pos = zero, ve_n = zero, cdir = zero;
    float woEz = 0.f, woLz = 0.f, cEz = 0.f, cLz = 0.f, hxyE = 0.f, hxyL = 0.f, g = 0.f, dist = 0.f;
    float ls_dir_pdf = 0.f, cam_dir_pdf = 0.f;
    bool gate = false, in_frame = true, push = false;
    unsigned int target = pix_own;

    // The records this strategy needs are requested TOGETHER, before anything waits for one of them: the eye vertex
    // (t >= 2), the light vertex (s >= 2) or — for s = 1 — the light subpath's origin, which the MIS chain reads
    // (SURVEY Q4); then both materials.  (One dependent round trip after the other was a fifth of this kernel's
    // stall samples: vertex -> its material -> the other vertex -> its material.)
    const bool needE = valid && t >= 2, needL = valid && s >= 1;
    float4 ea = make_float4(0.f, 0.f, 0.f, 0.f), eb = ea, ec = ea, ee = ea, la = ea, lb = ea, lc = ea, le = ea;
    if (needE) { const float4* pe = J.v0 + VSLOT(J, 0, t, sj); ld256_vtx(pe, &ea, &eb); ld256_vtx(pe + 2, &ec, &ee); }
    if (needL) { const float4* pl = J.v0 + VSLOT(J, 1, s >= 2 ? s : 1, sj); ld256_vtx(pl, &la, &lb); if (s >= 2) ld256_vtx(pl + 2, &lc, &le); }
    E.mat = -1; E.D = 1.f; E.S = 0.f; E.q = 1.f; ME.type = -1; ML.type = -1;
    if (needE) {
      E.pos = mk3(ea.x, ea.y, ea.z); E.mat = __float_as_int(ea.w); E.n = mk3(eb.x, eb.y, eb.z); E.D = eb.w;
      E.alpha = mk3(ec.x, ec.y, ec.z); E.S = ec.w; E.wo = mk3(ee.x, ee.y, ee.z); E.q = ee.w;
      ME = load_mat<TR>(sc, E.mat);
    } else {
      E.pos = J.cam.pos; E.alpha = mk3(1.f); E.n = zero; E.wo = zero;
    }
    if (valid && s >= 2) {
      Lv.pos = mk3(la.x, la.y, la.z); Lv.mat = __float_as_int(la.w); Lv.n = mk3(lb.x, lb.y, lb.z); Lv.D = lb.w;
      Lv.alpha = mk3(lc.x, lc.y, lc.z); Lv.S = lc.w; Lv.wo = mk3(le.x, le.y, le.z); Lv.q = le.w;
      ML = load_mat<TR>(sc, Lv.mat);
    }

    if (valid) {
      if (t >= 2) woEz = dot(E.n, E.wo);   // local z of the direction back along the eye subpath
      if (s == 0) {
        // c_{0,t}: the eye subpath hit an emitter (bidirection.cpp:307-328)
        f3 c = (t >= 2 && ME.type == BDPT_MAT_EMISSION) ? ME.radiance : zero;
        int lj = -1; Light L;
        if (norm(c) > BDPT_EPS_F) {
          for (int i = 0; i < sc.n_lights; i++) {
            L = load_light(sc, i);
            if (light_contain_point(L, E.pos)) { lj = i; break; }
          }
          if (lj < 0) c = zero;
          else { float pp, dp; c = light_sample_pdf(sc, L, E.pos, -E.wo, &pp, &dp); }
        }
        contrib = E.alpha * c;
        if (norm(contrib) > BDPT_EPS_F && lj >= 0) {
          // MIS, eye chain only (bidirection.cpp:159-176,181-188)
          const float r_t = (1.0f / L.area) / E.D;
          float tail = 0.f;
          bool delta_prev = false;
          if (t >= 3) {
            Vtx P = load_vtx(J, 0, t - 1, sj);
            delta_prev = type_is_delta(mat_type<TR>(sc, P.mat));
            if (P.S > 0.f) {
              float pp, dp;
              light_sample_pdf(sc, L, E.pos, -E.wo, &pp, &dp);  // dir pdf of leaving the light toward e_{t-1}
              f3 wv = P.pos - E.pos; float d2 = dot(wv, wv);
              f3 wd = wv * rsqrtf(d2);
              float gp = fabsf(dot(E.n, wd) * dot(wd, P.n)) / d2;
              tail = mis_pow(J, dp * gp / P.D) * P.S;
            }
          }
          if (t == 2 && lens_on) {
            // thin lens: the emitter could have emitted straight onto the lens point this eye subpath started from
            // (t' = 0)
            const f3 wv = lens_vertex(J, (unsigned long long)pix_own, (unsigned long long)s_smp[j]) - E.pos;
            const float d2 = dot(wv, wv);
            float pp, dp;
            light_sample_pdf(sc, L, E.pos, -E.wo, &pp, &dp);
            const float gp = fabsf(dot(E.n, wv * rsqrtf(d2))) / d2;     // (the camera vertex's cosine is 1 by convention)
            tail = mis_pow(J, dp * gp * (BDPT_PI * J.cam.lens_radius * J.cam.lens_radius));
          }
          double winv = 1.0 + (double)mis_pow(J, r_t) * ((double)(delta_prev ? 0.f : 1.f) + (double)tail);
          w = J.no_mis ? 1.f : (float)(1.0 / winv);
          if (!(w >= 0.f)) w = 0.f;
          ill = contrib * w;
          splat_rgb(J.acc_eye + pix_own, ill.x, ill.y, ill.z);
        }
      } else {
        // ---- light-side endpoint vl
        if (s == 1) {
          // fresh light point for this t (bidirection.cpp:332-358)
          Rng rng;
          rng.begin(J.seed, (unsigned long long)pix_own, (unsigned long long)s_smp[j], 2u + (uint32_t)t);
          int li = rng.pick(sc.n_lights);
          Light L = load_light(sc, li);
          float pA;
          f3 rad = light_sample_Le_point(sc, L, E.pos, rng, &Lv.pos, &pA, &ls_dir_pdf, &Lv.n);
          pA /= (float)sc.n_lights;
          Lv.alpha = rad * (1.0f / pA);
          Lv.mat = -1; Lv.q = 1.f; Lv.D = pA; Lv.S = 1.f; Lv.wo = zero;
        } else {
          woLz = dot(Lv.n, Lv.wo);
        }
        // ---- eye-side endpoint ve (t == 1: the pinhole, bidirection.cpp:360-383)
        ve_pos = E.pos; ve_n = E.n;
        f3 eye_alpha = E.alpha;
        if (t == 1) {
          f3 wi; float dd;
          float We;
          if (lens_on) {   // thin lens (extension): a fresh point of the lens, stream 64 + s
            Rng r2;
            r2.begin(J.seed, (unsigned long long)pix_own, (unsigned long long)s_smp[j], 64u + (uint32_t)s);
            const float u1 = r2.uniform(), u2 = r2.uniform();
            We = camera_sample_ray_pdf_lens(J.cam, Lv.pos, lens_point(J.cam, u1, u2), &ve_pos, &wi, &dd, &cam_dir_pdf, &eye_x, &eye_y);
          } else
          We = camera_sample_ray_pdf(J.cam, Lv.pos, &wi, &dd, &cam_dir_pdf, &eye_x, &eye_y);
          ve_n = -wi;
          eye_alpha = mk3(We);
          // a connection that leaves the frame never reaches the queue
          in_frame = eye_x >= 0 && eye_y >= 0 && eye_x < J.cam.W && eye_y < J.cam.H;
        }
        const f3 cv = Lv.pos - ve_pos;
        const float dist2 = dot(cv, cv);
        dist = sqrtf(dist2);
        cdir = cv * (1.0f / dist);
        // local z's of the connection at both ends (every BSDF here is isotropic about its normal: no tangent frames)
        cEz = dot(ve_n, cdir);               // t == 1: ve_n is the connection direction -> 1
        cLz = -dot(Lv.n, cdir);
        if (!(TR & BDPT_TRAIT_NOMF) && !(TR & BDPT_TRAIT_LAMBERT)) {
          if (ME.type == BDPT_MAT_MICROFACET) hxyE = half_xy2_world(E.wo, cdir, E.n);
          if (ML.type == BDPT_MAT_MICROFACET) hxyL = half_xy2_world(Lv.wo, -cdir, Lv.n);
        }
        // bsdf->f at both ends (bidirection.cpp:385-416)
        f3 f_eye = mk3(1.f), f_light = mk3(1.f);
        if (t >= 2) f_eye = bsdf_f_z(ME, woEz, cEz, hxyE);
        if (s >= 2) f_light = bsdf_f_z(ML, cLz, woLz, hxyL);
        g = fabsf(cLz * cEz) / dist2;
        contrib = eye_alpha * Lv.alpha * (f_eye * f_light) * g;
        gate = norm(contrib) > BDPT_EPS_F;   // bidirection.cpp:445
      }
    }
    if (gate && (TRACE || in_frame)) {   // (the trace hook records the weight of off-frame connections too)
      // ---- MIS weight (bidirection.cpp:121-293 in closed form, see header)
      double winv = 1.0;
      if (t >= 2) {
        float P_le = s == 1 ? ls_dir_pdf : bsdf_pdf_mis_z(ML, woLz, cLz, hxyL) * Lv.q;
        float r_t = P_le * g / E.D;
        float S_t = E.S;
        if (ME.type == BDPT_MAT_MICROFACET && (t >= 3 || lens_on)) {
          // wo-dependent density: the chain factor next to the connection cannot come from the walk's cache
          Vtx P;
          if (t >= 3) P = load_vtx(J, 0, t - 1, sj);
          else {   // thin lens: the camera vertex (walk_begin: density 1 / lens area, S = 1, cosine 1 by convention)
            P.pos = lens_vertex(J, (unsigned long long)pix_own, (unsigned long long)s_smp[j]);
            P.n = E.wo; P.mat = -1; P.S = 1.f;
            P.D = 1.0f / (BDPT_PI * J.cam.lens_radius * J.cam.lens_radius);
          }
          f3 wv = P.pos - E.pos; float d2 = dot(wv, wv);
          float gp = fabsf(woEz * dot(wv * rsqrtf(d2), P.n)) / d2;
          float R = bsdf_pdf_mis_z(ME, cEz, woEz, hxyE) * E.q * gp / P.D;
          S_t = clampS(((t >= 3 && type_is_delta(mat_type<TR>(sc, P.mat))) ? 0.f : 1.f) + (P.S > 0.f ? mis_pow(J, R) * P.S : 0.f));
        }
        winv += (double)mis_pow(J, r_t) * (double)S_t;
      }
      if (s >= 2) {
        float P_el = t == 1 ? cam_dir_pdf : bsdf_pdf_mis_z(ME, woEz, cEz, hxyE) * E.q;
        float r_s = P_el * g / Lv.D;
        float S_s = Lv.S;
        if (ML.type == BDPT_MAT_MICROFACET) {
          Vtx P = load_vtx(J, 1, s - 1, sj);   // s - 1 == 1: the light origin record holds D = p_A, S = 1
          f3 wv = P.pos - Lv.pos; float d2 = dot(wv, wv);
          float gp = fabsf(woLz * dot(wv * rsqrtf(d2), P.n)) / d2;
          float R = bsdf_pdf_mis_z(ML, cLz, woLz, hxyL) * Lv.q * gp / P.D;
          bool dprev = s - 1 >= 2 && type_is_delta(mat_type<TR>(sc, P.mat));
          S_s = clampS((dprev ? 0.f : 1.f) + (P.S > 0.f ? mis_pow(J, R) * P.S : 0.f));
        }
        winv += (double)mis_pow(J, r_s) * (double)S_s;
      } else {
        // s == 1: the chain is evaluated at the ORIGINAL light-subpath origin: light_path[1] = {la, lb}
        f3 wv = mk3(la.x, la.y, la.z) - ve_pos;
        float d2 = dot(wv, wv);
        f3 wd = wv * rsqrtf(d2);
        const float cosA = dot(ve_n, wd);
        float hx = 0.f;
        if (!(TR & BDPT_TRAIT_NOMF) && !(TR & BDPT_TRAIT_LAMBERT) && ME.type == BDPT_MAT_MICROFACET) hx = half_xy2_world(E.wo, wd, E.n);
        const float P_el = t == 1 ? cam_dir_pdf : bsdf_pdf_mis_z(ME, woEz, cosA, hx) * E.q;
        float gp = fabsf(cosA * dot(wd, mk3(lb.x, lb.y, lb.z))) / d2;
        float r_s = P_el * gp / lb.w;
        winv += (double)mis_pow(J, r_s);
      }
      if (t == 1 && lens_on) {
        // thin lens: the light vertex could also have scattered onto this very lens point (t' = 0)
        const float P_le = s == 1 ? ls_dir_pdf : bsdf_pdf_mis_z(ML, woLz, cLz, hxyL) * Lv.q;
        winv += (double)mis_pow(J, P_le * g * (BDPT_PI * J.cam.lens_radius * J.cam.lens_radius));
      }
      w = J.no_mis ? 1.f : (float)(1.0 / winv);
      if (!(w >= 0.f)) w = 0.f;  // NaN guard
      ill = contrib * w;
      push = w > 0.f && in_frame;
      if (t == 1 && in_frame) target = (unsigned int)(eye_y * J.cam.W + eye_x) | 0x80000000u;
    }
    if (TRACE) {
      if (valid && s > 0) {
        // trace mode resolves visibility inline so that the record says "occluded" (NaN) like the oracle's
        Hit h = traverse<true, TR>(sc, ve_pos, cdir, BDPT_EPS_F, dist - BDPT_EPS_F);
        if (h.leaf >= 0) { contrib = mk3(NAN); w = NAN; push = false; eye_x = eye_y = -1; }
      }
      if (valid && J.dbg) {
        double* rr = J.dbg + 8 * ((size_t)(t - 1) * (size_t)(J.m + 2) + (size_t)s);
        rr[0] = t; rr[1] = s; rr[2] = contrib.x; rr[3] = contrib.y; rr[4] = contrib.z; rr[5] = w; rr[6] = eye_x; rr[7] = eye_y;
      }
    }
    // ---- append {ray, pixel, weighted contribution} to this block's segment of the shadow queue: one shared-memory
    // atomic per warp (every lane of the warp is here — the round loop is uniform — so the votes are full-mask)
    const unsigned a_mask = __ballot_sync(0xffffffffu, push);
    if (a_mask != 0u) {
      const int leader = __ffs(a_mask) - 1;
      unsigned ticket = 0u;
      if (lane == leader) ticket = atomicAdd(&s_qn, (unsigned int)__popc(a_mask));
      const unsigned slot0 = __shfl_sync(0xffffffffu, ticket, leader);
      if (push) {
        const unsigned k_in = slot0 + __popc(a_mask & ((1u << lane) - 1u));
        if (k_in >= J.seg_cap) J.qctr[4] = 1u;
        else {
          const size_t idx = (size_t)blockIdx.x * J.seg_cap + k_in;
          // (the pixel travels with the contribution: the splat then needs one 16-byte load, not two)
          st256_q(J.q0 + 2 * idx, make_float4(ve_pos.x, ve_pos.y, ve_pos.z, dist - BDPT_EPS_F),
                make_float4(cdir.x, cdir.y, cdir.z, 0.f));
          J.q2[idx] = make_float4(ill.x, ill.y, ill.z, __uint_as_float(target));
        }
      }
    }
  }
  __syncthreads();
  {
    // publish the segment as chunks of 256 entries: one global atomic per block
    const unsigned n = min(s_qn, J.seg_cap), nchunks = (n + 255u) >> 8;
    __shared__ unsigned s_c0;
    if (threadIdx.x == 0 && nchunks) s_c0 = atomicAdd(J.qctr + 0, nchunks);
    __syncthreads();
    for (unsigned c = threadIdx.x; c < nchunks; c += NT) {
      const unsigned slot = s_c0 + c;
      if (slot < J.chunk_cap) J.chunks[slot] = make_uint2(blockIdx.x * J.seg_cap + (c << 8), min(256u, n - (c << 8)));
      else J.qctr[4] = 1u;
    }
  }
  const unsigned tc = __reduce_add_sync(0xffffffffu, nconn);
  if (lane == 0 && tc) atomicAdd(J.stats + 2, (unsigned long long)tc);
}

// -------------------------------------------------------------------------------------- shadow
// Adds unoccluded connections to their pixels.  Called by the lanes of `mask` together (every lane of it holds
// queue entry idx).  In scenes with a point light (J.point_lights) the t=1, s=1 connection of EVERY sample is the same
// ray and lands on the same pixel of the light image; 38 M same-address atomics per batch took 57 of CBempty's
// 60 ms.  There, lanes of the set that splat the same light-image pixel are summed first and their leader issues the
// reduction.
__device__ inline void shadow_splat(const Job& J, unsigned mask, unsigned idx) {
  const float4 c = J.q2[idx];
  const unsigned int tg = __float_as_uint(c.w);
  float sx = c.x, sy = c.y, sz = c.z;
  if (J.point_lights) {
    const unsigned li = __ballot_sync(mask, (tg & 0x80000000u) != 0u);
    if (tg & 0x80000000u) {
      const unsigned peers = __match_any_sync(li, tg);
      if (peers & (peers - 1u)) {   // more than this lane
        sx = sy = sz = 0.f;
        for (unsigned rem = peers; rem; rem &= rem - 1u) {
          const int src = __ffs(rem) - 1;
          sx += __shfl_sync(peers, c.x, src); sy += __shfl_sync(peers, c.y, src); sz += __shfl_sync(peers, c.z, src);
        }
        if ((int)(threadIdx.x & 31) != __ffs(peers) - 1) return;
      }
    }
  }
  splat_rgb(((tg & 0x80000000u) ? J.acc_light : J.acc_eye) + (size_t)(tg & 0x7fffffffu), sx, sy, sz);
}

// ------------------------------------------------------------------- persistent (scheduled) kernels
// Giving every lane one ray (or one subpath) and waiting for the slowest lane of the warp ran at 6.6 (walk) and 9.0
// (shadow) active lanes per instruction (profiles/r01_*).  The persistent forms keep a lane busy: warps claim work in
// chunks with one atomic per chunk, hand indices to idle lanes with ballot + popc prefix, and run the resumable
// traversal (Trav::step) as a per-lane state machine; lanes refill the moment their ray ends.
struct WarpFetch {   // warp-uniform
  unsigned next, end;
  bool exhausted;
};

#define BDPT_NONE 0xffffffffu

// All 32 lanes call this.  Lanes with want == true receive distinct work indices (or BDPT_NONE).
__device__ inline unsigned warp_fetch(unsigned* cursor, unsigned total, unsigned chunk, bool want, WarpFetch& wf) {
  const unsigned mask = __ballot_sync(0xffffffffu, want);
  if (mask == 0u || wf.exhausted) return BDPT_NONE;
  const int lane = threadIdx.x & 31;
  if (wf.next >= wf.end) {
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(cursor, chunk);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= total) { wf.exhausted = true; return BDPT_NONE; }
    wf.next = base;
    wf.end = min(base + chunk, total);
  }
  const unsigned my = wf.next + __popc(mask & ((1u << lane) - 1u));
  const bool got = want && my < wf.end;
  wf.next = min(wf.next + (unsigned)__popc(mask), wf.end);
  return got ? my : BDPT_NONE;
}

// The shadow queue is a list of chunks of <= 256 entries (Job::chunks, published by k_connect's blocks): a warp takes a
// chunk with one atomic and hands its entries to idle lanes as above.
__device__ inline unsigned warp_fetch_chunk(const Job& J, unsigned nchunks, bool want, WarpFetch& wf) {
  const unsigned mask = __ballot_sync(0xffffffffu, want);
  if (mask == 0u || wf.exhausted) return BDPT_NONE;
  const int lane = threadIdx.x & 31;
  if (wf.next >= wf.end) {
    unsigned c = 0;
    if (lane == 0) c = atomicAdd(J.qctr + 1, 1u);
    c = __shfl_sync(0xffffffffu, c, 0);
    if (c >= nchunks) { wf.exhausted = true; return BDPT_NONE; }
    const uint2 d = __ldg(J.chunks + c);
    wf.next = d.x;
    wf.end = d.x + d.y;
    // the chunk's entries are handed out over the next rounds: request them into L2 (256 x 32 B of rays, 256 x 16 B of
    // contributions; shadow kernel -1.8 %)
    {
      const char* r0 = (const char*)(J.q0 + 2 * (size_t)d.x); const unsigned nb0 = d.y * 32u;
      for (unsigned off = (unsigned)lane * 128u; off < nb0; off += 32u * 128u) asm volatile("prefetch.global.L2 [%0];" :: "l"(r0 + off));
      const char* r2 = (const char*)(J.q2 + (size_t)d.x); const unsigned nb2 = d.y * 16u;
      for (unsigned off = (unsigned)lane * 128u; off < nb2; off += 32u * 128u) asm volatile("prefetch.global.L2 [%0];" :: "l"(r2 + off));
    }
  }
  const unsigned my = wf.next + __popc(mask & ((1u << lane) - 1u));
  const bool got = want && my < wf.end;
  wf.next = min(wf.next + (unsigned)__popc(mask), wf.end);
#if BDPT_SHADOW_PF_NEXT
  // the entries the NEXT hand-out will give away: into L1 now (BDPT_SHADOW_PF_NEXT of them, 32 B each)
  if (lane < BDPT_SHADOW_PF_NEXT && wf.next + lane < wf.end) asm volatile("prefetch.global.L1 [%0];" :: "l"(J.q0 + 2 * (size_t)(wf.next + lane)));
#endif
  return got ? my : BDPT_NONE;
}

// Copies the first `ntop` BVH nodes (the upper levels: nodes are stored breadth first) into shared memory.
__device__ inline void stage_top_nodes(const Job& J, uint4* top) {
  for (int i = threadIdx.x; i < J.ntop * 5; i += blockDim.x) top[i] = J.sc.nodes[(i / 5) * BDPT_NODE_Q + i % 5];
  __syncthreads();
}

// FETCH: idle lanes a warp tolerates before it fetches; SPLAT: parked rays it collects before it splats.  Round 1
// (6.2 ms kernel): FETCH 4: 6.53, 8: 6.31, 12: 6.22, 16: 6.24 ms per 2^24 samples; SPLAT 8: 6.62, 16: 6.53, 24: 6.70.
// Re-measured on the 4.75 ms kernel (PLOC tree, one-load splats, L1-prefetched hand-outs), CBbunny4: FETCH 8: 4.734,
// 12: 4.755, 16: 4.93; SPLAT 12: 4.732, 16: 4.755, 22: 4.81.
#ifndef BDPT_SHADOW_FT
#define BDPT_SHADOW_FT 8
#endif
#ifndef BDPT_SHADOW_SP
#define BDPT_SHADOW_SP 12
#endif
template <int MINB, int TR = 0, int FETCH = BDPT_SHADOW_FT, int SPLAT = BDPT_SHADOW_SP>
__global__ void __launch_bounds__(128, MINB) k_shadow_pt(const Job J) {
  extern __shared__ uint4 top[];
  stage_top_nodes(J, top);
  const unsigned nchunks = min(J.qctr[0], J.chunk_cap);
  Trav T;
  uint2 stack[BDPT_STACK];
  bool active = false;
  unsigned idx = 0, traced = 0;
  // Rays end a few lanes at a time (ncu: the splat ran at 4-6 lanes).  A lane parks its unoccluded ray here and goes
  // on with the next one; the warp splats when half of its lanes have one parked.
  unsigned pend = BDPT_NONE;
  WarpFetch wf; wf.next = 0; wf.end = 0; wf.exhausted = false;
#pragma unroll 1
  for (;;) {
    const unsigned idle = __ballot_sync(0xffffffffu, !active);
    const unsigned parked = __ballot_sync(0xffffffffu, pend != BDPT_NONE);
    if (__popc(parked) >= SPLAT) {
      if (pend != BDPT_NONE) { shadow_splat(J, parked, pend); pend = BDPT_NONE; }
    }
    if (!wf.exhausted && (__popc(idle) >= FETCH || idle == 0xffffffffu)) {
      unsigned my = warp_fetch_chunk(J, nchunks, !active, wf);
      if (my != BDPT_NONE) {
        float4 a, b; ld256_q(J.q0 + 2 * (size_t)my, &a, &b);
        T.begin(mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), BDPT_EPS_F, a.w);
        idx = my; active = true;
      }
    }
    if (__ballot_sync(0xffffffffu, active) == 0u) {
      if (wf.exhausted) break;
      continue;
    }
    if (active && T.step<true, TR>(J.sc, stack, top, (uint32_t)J.ntop)) {
      traced++;
      active = false;
      if (T.hit.leaf < 0) {
        if (pend != BDPT_NONE) shadow_splat(J, 1u << (threadIdx.x & 31), pend);   // the slot is taken: this lane cannot wait
        pend = idx;
      }
    }
  }
  const unsigned left = __ballot_sync(0xffffffffu, pend != BDPT_NONE);
  if (pend != BDPT_NONE) shadow_splat(J, left, pend);
  unsigned int tot = __reduce_add_sync(0xffffffffu, traced);
  if ((threadIdx.x & 31) == 0 && tot) atomicAdd(J.stats + 1, (unsigned long long)tot);
}

// phase: 0 = empty (wants a subpath), 1 = traversing, 2 = traversal finished (wants shading)
template <int SHADE_THRESHOLD, int FETCH_THRESHOLD, int MINB = 4, int TR = 0>
__global__ void __launch_bounds__(128, MINB) k_walk_pt(const __grid_constant__ Job J) {
  extern __shared__ uint4 top[];
  stage_top_nodes(J, top);
  Trav T;
  uint2 stack[BDPT_STACK];
  WalkState W;
  int phase = 0;
  unsigned nrays = 0;
  WarpFetch wf; wf.next = 0; wf.end = 0; wf.exhausted = false;
  const unsigned total = 2u * (unsigned)J.B;
#pragma unroll 1
  for (;;) {
    const unsigned trav = __ballot_sync(0xffffffffu, phase == 1);
    const unsigned shade = __ballot_sync(0xffffffffu, phase == 2);
    // shading waits until enough lanes want it (or nothing else can run) ...
    if (shade != 0u && (trav == 0u || __popc(shade) >= SHADE_THRESHOLD)) {
      if (phase == 2) {
        if (walk_shade<TR>(J, W, T)) phase = 1;
        else { walk_end(J, W); phase = 0; }
      }
    }
    // ... and so does starting new subpaths
    const unsigned empty = __ballot_sync(0xffffffffu, phase == 0);
    if (empty != 0u && !wf.exhausted && (__popc(empty) >= FETCH_THRESHOLD || __ballot_sync(0xffffffffu, phase == 1) == 0u)) {
      unsigned my = warp_fetch(J.qctr + 3, total, 128u, phase == 0, wf);
      if (my != BDPT_NONE && walk_begin<TR>(J, my, W, T)) phase = 1;
    }
    if (__ballot_sync(0xffffffffu, phase != 0) == 0u && wf.exhausted) break;
    if (phase == 1) {
      // (requesting the winner's normals and fp64 geometry into L1 here, for the shading that follows a few rounds later —
      // three prefetch.global.L1 per ray — made the walk 37 % slower)
      if (T.step<false, TR>(J.sc, stack, top, (uint32_t)J.ntop)) { phase = 2; nrays++; }
    }
  }
  unsigned int tot = __reduce_add_sync(0xffffffffu, nrays);
  if ((threadIdx.x & 31) == 0 && tot) atomicAdd(J.stats + 0, (unsigned long long)tot);
}

// ------------------------------------------------------------------------------ test-hook kernels
__global__ void k_intersect_batch(DScene sc, const float* rays, size_t n, float* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* r = rays + 8 * i;
  f3 o = mk3(r[0], r[1], r[2]), d = mk3(r[3], r[4], r[5]);
  Hit h = traverse<false>(sc, o, d, r[6], r[7]);
  float* q = out + 6 * i;
  if (h.leaf < 0) { q[0] = -1.f; q[1] = INFINITY; q[2] = q[3] = q[4] = 0.f; q[5] = -1.f; return; }
  f3 nn, hp; int prim, mat;
  finish_hit(sc, h, o, d, &nn, &prim, &mat, &hp);
  q[0] = __int_as_float(prim); q[1] = h.t; q[2] = nn.x; q[3] = nn.y; q[4] = nn.z; q[5] = __int_as_float(mat);
}

__global__ void k_occluded_batch(DScene sc, const float* rays, size_t n, int* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* r = rays + 8 * i;
  Hit h = traverse<true>(sc, mk3(r[0], r[1], r[2]), mk3(r[3], r[4], r[5]), r[6], r[7]);
  out[i] = h.leaf >= 0 ? 1 : 0;
}

__global__ void k_bsdf_eval(DScene sc, const float* in, size_t n, float* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* q = in + 7 * i; float* o = out + 8 * i;
  Mat M = load_mat(sc, (int)q[0]);
  f3 wo = mk3(q[1], q[2], q[3]), wi = mk3(q[4], q[5], q[6]);
  f3 f = bsdf_f(M, wo, wi);
  o[0] = f.x; o[1] = f.y; o[2] = f.z;
  o[3] = M.type == BDPT_MAT_MICROFACET ? NAN : bsdf_pdf(M, mk3(0.f), wi);
  o[4] = type_is_delta(M.type) ? 1.f : 0.f;
  f3 e = M.type == BDPT_MAT_EMISSION ? M.radiance : mk3(0.f);
  o[5] = e.x; o[6] = e.y; o[7] = e.z;
}

// explicit-uniform sampler: a tiny Rng stand-in that replays q[4..6] = {u1, u2, u_coin}
struct ScriptRng {
  const float* u; int pos;
  __device__ float uniform() { return u[pos++]; }
};

__global__ void k_bsdf_sample(DScene sc, const float* in, size_t n, float* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* q = in + 7 * i; float* o = out + 7 * i;
  Mat M = load_mat(sc, (int)q[0]);
  ScriptRng rng; rng.u = q + 4; rng.pos = M.type == BDPT_MAT_GLASS ? 2 : 0;  // glass draws only the coin
  f3 wi; float pdf = 0.f;
  f3 f = bsdf_sample_f(M, mk3(q[1], q[2], q[3]), &wi, &pdf, rng);
  o[0] = wi.x; o[1] = wi.y; o[2] = wi.z; o[3] = pdf; o[4] = f.x; o[5] = f.y; o[6] = f.z;
}

__global__ void k_light_eval(DScene sc, const float* in, size_t n, float* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* q = in + 7 * i; float* o = out + 6 * i;
  Light L = load_light(sc, (int)q[0]);
  f3 p = mk3(q[1], q[2], q[3]), wi = mk3(q[4], q[5], q[6]);
  float pp = 0.f, dp = 0.f;
  o[0] = light_contain_point(L, p) ? 1.f : 0.f;
  f3 le = light_sample_pdf(sc, L, p, wi, &pp, &dp);
  o[1] = le.x; o[2] = le.y; o[3] = le.z; o[4] = pp; o[5] = dp;
}

__global__ void k_camera_eval(DCamera cam, const float* in, size_t n, float* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* q = in + 5 * i; float* o = out + 11 * i;
  f3 ro, rd;
  camera_generate_ray(cam, q[0], q[1], &ro, &rd);
  o[0] = ro.x; o[1] = ro.y; o[2] = ro.z; o[3] = rd.x; o[4] = rd.y; o[5] = rd.z;
  f3 wi; float dist, dp; int px, py;
  float We = camera_sample_ray_pdf(cam, mk3(q[2], q[3], q[4]), &wi, &dist, &dp, &px, &py);
  o[6] = We; o[7] = dp; o[8] = (float)px; o[9] = (float)py; o[10] = dist;
}

// toColor, reference util/image.h:194-209: exposure sqrt(2), gamma 2.2, clamp, truncate to u8.
// Output row 0 = top (RaytracedRenderer::save_image flips, raytraced_renderer.cpp:690-728).
__device__ inline f3 resolve_pixel(const float4* eye, const float4* light, int which, float inv_ns, size_t p) {
  const float4 e = eye[p], l = light[p];
  const f3 v = which == 0 ? mk3(e.x + l.x, e.y + l.y, e.z + l.z) : which == 1 ? mk3(e.x, e.y, e.z) : mk3(l.x, l.y, l.z);
  return v * inv_ns;
}

__global__ void k_tonemap(const float4* eye, const float4* light, int which, float inv_ns, int W, int H, uint8_t* rgba) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= W * H) return;
  int x = p % W, y = p / W;
  const f3 px = resolve_pixel(eye, light, which, inv_ns, (size_t)p);
  const float v[3] = {px.x, px.y, px.z};
  uint8_t* o = rgba + 4 * ((size_t)(H - 1 - y) * W + x);
  for (int c = 0; c < 3; c++) {
    float g = powf(fmaxf(v[c], 0.f) * 1.41421356f, 1.0f / 2.2f);
    g = fminf(fmaxf(g, 0.f), 1.f);
    o[c] = (uint8_t)(g * 255.f);
  }
  o[3] = 255;
}

// End of a batch: the batch's splats (`batch`, the target of the kernels' reductions) are added to the frame's sums with
// Kahan's compensated addition and the batch accumulators are zeroed.  An fp32 sum of N equal splats stops growing at
// 2^24 of them; the one pixel that images a point light receives a splat from EVERY sample of the frame (8e8 at
// 1024x768x1024 spp).  Inside a batch (<= 2^24 samples where a point light exists: api.cu) the sums stay exact multiples of the splat; across batches the
// compensation term keeps what the fp32 frame sum cannot hold, so its error stays at one rounding for any frame.
__global__ void k_fold(float4* __restrict__ acc, float4* __restrict__ comp, float4* __restrict__ batch, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 b = batch[i];
  if (b.x == 0.f && b.y == 0.f && b.z == 0.f) return;   // pixel untouched by this batch (NaN compares false: folded)
  float4 a = acc[i], c = comp[i];
  float y, t;
  y = b.x - c.x; t = a.x + y; c.x = (t - a.x) - y; a.x = t;
  y = b.y - c.y; t = a.y + y; c.y = (t - a.y) - y; a.y = t;
  y = b.z - c.z; t = a.z + y; c.z = (t - a.z) - y; a.z = t;
  acc[i] = a; comp[i] = c;
  batch[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// RGBA-padded accumulators -> the W*H*3 image the caller reads, divided by ns
__global__ void k_resolve(const float4* eye, const float4* light, int which, float inv_ns, size_t npix, float* out) {
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const f3 v = resolve_pixel(eye, light, which, inv_ns, p);
  out[3 * p] = v.x; out[3 * p + 1] = v.y; out[3 * p + 2] = v.z;
}

}  // namespace bdpt
