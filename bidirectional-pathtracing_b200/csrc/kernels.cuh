// kernels.cuh — the wavefront BDPT kernels.
//
// One batch of B samples goes through three kernels:
//   k_walk     one thread per subpath (B eye + B light): random walk with closest-hit traversal
//              and BSDF sampling; writes path vertices to SoA planes in HBM
//              (reference prepare_bidirectional_subpath + sample_light_ray, bidirection.cpp:20-118)
//   k_connect  one thread per (sample, t): loops the light-subpath length s, evaluates the
//              unweighted contribution and the MIS weight of strategy (s,t), and appends a shadow
//              ray {segment, pixel, weighted contribution} to a queue with warp-aggregated atomics
//              (reference estimate_bidirection_radiance + multiple_importance_sampling_weight,
//              bidirection.cpp:121-469)
//   k_shadow   persistent warps pull 32 queue entries at a time, any-hit traversal, splat the
//              contribution if unoccluded (the reference's shadow bvh->intersect, :419-433, and
//              update_pixel, :544-551)
//
// MIS in O(1) per strategy.  The reference walks both ratio chains for every strategy (O(s+t)
// each, O(m^3) per sample).  Every factor of those chains except the one at the connection is a
// function of one subpath alone, so k_walk caches per vertex j
//     D_j = pdf_{j-1}(-> j) * G(j-1, j)        (the chain's denominator at j)
//     S_j = sum_{i<=j} flag_i * prod_{l=i}^{j-1} (R_l)^2,  R_l = pdf_{l+1}(-> l) * G(l+1, l) / D_l
// and a strategy's weight is 1 / (1 + r_t^2 S_t + r_s^2 S^L_s) with r the connection ratio.
// Same numbers as the reference's loops (bidirection.cpp:133-221,224-289), including its quirks:
// wo = 0 in sample_pdf (Q5), the original light vertex in the s = 1 light chain (Q4), the
// plane-only contain_point (Q6), light_path[0].alpha = 1 for s = 0 (Q8), no t = 0 (Q9).
#pragma once
#include "common.cuh"
#include "rng.cuh"
#include "shading.cuh"
#include "traverse.cuh"

namespace bdpt {

struct Job {
  DScene sc;
  DCamera cam;
  // vertex records [2*(m+1)][B] x 64 bytes (v0..v3: the four float4 of a record, interleaved); counts [2][B]
  float4 *v0, *v1, *v2, *v3;
  int* counts;
  // shadow queue, SoA float4 x 3, and its counters {n_entries, cursor, overflow, -}
  float4 *q0, *q1, *q2;
  unsigned int* qctr;
  unsigned int qcap;
  // accumulators (un-normalised), W*H*3 floats each
  float *acc_eye, *acc_light;
  unsigned long long* stats;  // {closest rays, shadow rays, connections}
  double* dbg;                // trace of one sample (bdpt_trace_sample) or null
  unsigned long long seed;
  unsigned long long g0;      // first job-sample index of this batch
  unsigned int k0, p0;        // g0 = k0 * npix_pad + p0 (split on the host: no 64-bit division on the device)
  unsigned long long g_end;   // one past the last job-sample index of the job
  int B;                      // samples in this batch
  int m;                      // max_ray_depth
  int first_sample;           // sample index of job-sample 0
  int tiles_x, npix_pad;      // 8x4 pixel tiles per row; padded pixel count
  FastDiv div_tiles_x, div_npix_pad;
  int roulette;
  int ntop;                   // BVH nodes staged in shared memory by the persistent kernels
  int no_mis;                 // weight 1 for every strategy: the reference's own debug switch (bidirection.cpp:124-126)
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
  *x = tx * 8 + (w & 7);
  *y = ty * 4 + (w >> 3);
  *sample = J.first_sample + (int)k;
  return *x < J.cam.W && *y < J.cam.H;
}

// float4 index of the first quarter of the 64-byte record of vertex j of subpath (kind, i); v1..v3 = v0 + 1..3
#define VSLOT(J, kind, j, i) (4 * ((size_t)((kind) * ((J).m + 1) + ((j)-1)) * (size_t)(J).B + (size_t)(i)))

// 32-byte global accesses (LDG.256 / STG.256, sm_100): half of a vertex record per instruction
__device__ inline void ld256(const float4* p, float4* a, float4* b) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a->x), "=f"(a->y), "=f"(a->z), "=f"(a->w), "=f"(b->x), "=f"(b->y), "=f"(b->z), "=f"(b->w) : "l"(p));
}
__device__ inline void st256(float4* p, float4 a, float4 b) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}
// the 64-byte record of a vertex: {pos, material | n, D | alpha, S | wo, q}
__device__ inline void store_vertex(const Job& J, size_t slot, float4 a, float4 b, float4 c, float4 e) {
  st256(J.v0 + slot, a, b);
  st256(J.v0 + slot + 2, c, e);
}

__device__ inline float clampS(float s) { return fminf(s, 1e30f); }

// ---------------------------------------------------------------------------------------- walk
template <int TR = 0>
__global__ void __launch_bounds__(128) k_walk(const Job J) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * J.B) return;
  const int kind = i >= J.B ? 1 : 0;  // 0 eye, 1 light (warp-uniform: B is a multiple of 32)
  const int si = i - kind * J.B;
  int x, y, sample;
  if (!decode_sample(J, J.g0 + (unsigned long long)si, &x, &y, &sample)) { J.counts[i] = 0; return; }
  const DScene& sc = J.sc;
  Rng rng;
  rng.begin(J.seed, (unsigned long long)y * (unsigned long long)J.cam.W + (unsigned long long)x,
            (unsigned long long)sample, (uint32_t)kind);

  f3 o, d, prev_n, alpha, prev_f = mk3(1.f);
  float tmin, tmax, prev_pdf, pdfeval_prev, D_prev, S_prev;
  bool delta_prev = false;
  int nv = 1;
  if (kind == 0) {
    // raytrace_pixel bidirection.cpp:517-524 and v1 of the eye subpath (:29-37)
    float sx = rng.uniform(), sy = rng.uniform();
    camera_generate_ray(J.cam, ((float)x + sx) / (float)J.cam.W, ((float)y + sy) / (float)J.cam.H, &o, &d);
    tmin = J.cam.nclip; tmax = J.cam.fclip;
    prev_n = d; alpha = mk3(1.f); prev_pdf = 1.f;
    pdfeval_prev = 1.f; D_prev = 1.f; S_prev = 0.f;
  } else {
    // sample_light_ray bidirection.cpp:105-118 and v1 of the light subpath
    int li = rng.pick(sc.n_lights);
    Light L = load_light(sc, li);
    float pA, pdir; f3 nl;
    f3 Le = light_sample_Le(L, rng, &o, &d, &pA, &pdir, &nl);
    pA /= (float)sc.n_lights;
    tmin = BDPT_EPS_F; tmax = INFINITY;
    prev_n = nl; alpha = Le * (1.0f / pA); prev_pdf = pdir;
    pdfeval_prev = pdir; D_prev = pA; S_prev = 1.f;
    size_t s1 = VSLOT(J, 1, 1, si);
    store_vertex(J, s1, make_float4(o.x, o.y, o.z, __int_as_float(-1)),
                 make_float4(nl.x, nl.y, nl.z, pA),
                 make_float4(alpha.x, alpha.y, alpha.z, 1.f),
                 make_float4(pdir, 0.f, 0.f, 1.f));
  }

  unsigned int nrays = 0;
  for (int j = 2; j <= J.m + 1; j++) {
    Hit h = traverse<false, TR>(sc, o, d, tmin, tmax);
    nrays++;
    if (h.leaf < 0) break;
    f3 n, hit_p; int prim, mat;
    finish_hit<TR>(sc, h, o, d, &n, &prim, &mat, &hit_p);
    const Frame fr = make_frame(n);
    const f3 w_out = fr.to_local(-d);
    const Mat M = load_mat<TR>(sc, mat);
    f3 wi; float pdf;
    const f3 f = bsdf_sample_f(M, w_out, &wi, &pdf, rng);
    const f3 wi_world = normalized(fr.to_world(wi));

    // bidirection.cpp:64-69
    const float cos_prev = fabsf(dot(prev_n, d));
    const float g = cos_prev * fabsf(dot(n, d)) / (h.t * h.t);
    alpha = alpha * prev_f * (cos_prev / prev_pdf);

    // roulette (off in the reference: bidirection.cpp:80; the disabled rule is :87-93)
    float p_keep = 1.f;
    if (J.roulette && j > 2) p_keep = fminf(1.f, norm(f) / pdf);

    // MIS cache
    const bool delta_j = type_is_delta(M.type);
    const float D_j = (kind == 0 && j == 2) ? 1.f : pdfeval_prev * g;
    const f3 zero = mk3(0.f);
    const float D_rev = bsdf_pdf_mis(M, wi, w_out) * p_keep * g;  // density of (j -> j-1) as the chains use it
    float S_j = (delta_j || delta_prev) ? 0.f : 1.f;
    if (S_prev > 0.f) { float R = D_rev / D_prev; S_j += R * R * S_prev; }
    S_j = clampS(S_j);

    size_t sl = VSLOT(J, kind, j, si);
    store_vertex(J, sl, make_float4(hit_p.x, hit_p.y, hit_p.z, __int_as_float(mat)),
                 make_float4(n.x, n.y, n.z, D_j),
                 make_float4(alpha.x, alpha.y, alpha.z, S_j),
                 make_float4(-d.x, -d.y, -d.z, p_keep));
    nv = j;
    if (j >= J.m + 1) break;
    if (J.roulette && !(rng.uniform() < p_keep)) break;
    if (is_zero(f)) break;  // throughput is exactly zero from here on: later vertices add nothing

    pdfeval_prev = bsdf_pdf_mis(M, w_out, wi) * p_keep;
    prev_f = f; prev_n = n; prev_pdf = pdf * p_keep;
    D_prev = D_j; S_prev = S_j; delta_prev = delta_j;
    o = hit_p; d = wi_world; tmin = BDPT_EPS_F; tmax = INFINITY;
  }
  J.counts[i] = nv;
  // one atomic per warp for the ray counter
  unsigned int tot = __reduce_add_sync(__activemask(), nrays);
  if ((threadIdx.x & 31) == (__ffs(__activemask()) - 1)) atomicAdd(J.stats + 0, (unsigned long long)tot);
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
  ld256(J.v0 + s, &a, &b); ld256(J.v0 + s + 2, &c, &e);
  Vtx v;
  v.pos = mk3(a.x, a.y, a.z); v.mat = __float_as_int(a.w);
  v.n = mk3(b.x, b.y, b.z); v.D = b.w;
  v.alpha = mk3(c.x, c.y, c.z); v.S = c.w;
  v.wo = mk3(e.x, e.y, e.z); v.q = e.w;
  return v;
}

__device__ inline void push_shadow(const Job& J, bool push, f3 o, f3 dir, float tmax, unsigned int target, f3 ill) {
  const unsigned int active = __activemask();
  const unsigned int mask = __ballot_sync(active, push);
  if (mask == 0u) return;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(mask) - 1;
  unsigned int base = 0;
  if (lane == leader) base = atomicAdd(J.qctr + 0, (unsigned int)__popc(mask));
  base = __shfl_sync(active, base, leader);
  if (!push) return;
  unsigned int idx = base + __popc(mask & ((1u << lane) - 1u));
  if (idx >= J.qcap) { J.qctr[2] = 1u; return; }
  // the ray {o, tmax | dir, pixel} is one 32-byte record (q0 / q1 interleaved), its contribution a separate plane
  st256(J.q0 + 2 * (size_t)idx, make_float4(o.x, o.y, o.z, tmax), make_float4(dir.x, dir.y, dir.z, __uint_as_float(target)));
  J.q2[idx] = make_float4(ill.x, ill.y, ill.z, 0.f);
}

struct VRaw { float4 a, b, c, e; };
__device__ inline VRaw load_raw(const Job& J, int kind, int j, int si) {
  size_t s = VSLOT(J, kind, j, si);
  VRaw r; ld256(J.v0 + s, &r.a, &r.b); ld256(J.v0 + s + 2, &r.c, &r.e);
  return r;
}
__device__ inline Vtx unpack(const VRaw& r) {
  Vtx v;
  v.pos = mk3(r.a.x, r.a.y, r.a.z); v.mat = __float_as_int(r.a.w);
  v.n = mk3(r.b.x, r.b.y, r.b.z); v.D = r.b.w;
  v.alpha = mk3(r.c.x, r.c.y, r.c.z); v.S = r.c.w;
  v.wo = mk3(r.e.x, r.e.y, r.e.z); v.q = r.e.w;
  return v;
}

// Where k_connect's threads find the vertices of their sample.  (A variant that staged a block's vertices
// in shared memory first — S samples x (m+1) threads per block — measured 2.3x SLOWER on B200: blocks then
// live as long as their slowest t, and the occupancy lost to that outweighs the latency saved.)
struct GlobalVerts {
  const Job& J; int si;
  __device__ inline Vtx get(int kind, int j) const { return load_vtx(J, kind, j, si); }
  __device__ inline void light_origin(float4* a, float4* b) const { size_t s1 = VSLOT(J, 1, 1, si); ld256(J.v0 + s1, a, b); }
};
// All strategies (s = 0 .. nL) of one (sample, t): estimate_bidirection_radiance + the closed-form MIS weight.
template <bool TRACE, int TR, class VERTS>
__device__ inline void connect_strategies(const Job& J, const int si, const int t, const int nL, const VERTS& V) {
  int x, y, sample;
  decode_sample(J, J.g0 + (unsigned long long)si, &x, &y, &sample);
  const DScene& sc = J.sc;
  const f3 zero = mk3(0.f);

  Vtx E; Mat ME;
  E.mat = -1; E.D = 1.f; E.S = 0.f; E.q = 1.f;
  if (t >= 2) {
    E = V.get(0, t);
    ME = load_mat<TR>(sc, E.mat);
  } else {
    E.pos = J.cam.pos; E.alpha = mk3(1.f); E.n = zero; E.wo = zero;
    ME.type = -1;
  }
  // local z of the direction back along the eye subpath (every BSDF here needs only z's and one dot product)
  const float woEz = t >= 2 ? dot(E.n, E.wo) : 0.f;
  const unsigned int pix_own = (unsigned int)(y * J.cam.W + x);
  unsigned int nconn = 0;

  for (int s = 0; s <= nL; s++) {
    if (J.only_t >= 0 && (t != J.only_t || s != J.only_s)) continue;
    f3 contrib = zero, ill = zero;
    float w = 0.f;
    int eye_x = -1, eye_y = -1;
    nconn++;
    if (s == 0) {
      // c_{0,t}: the eye subpath hit an emitter (bidirection.cpp:307-328)
      f3 c = (t >= 2 && ME.type == BDPT_MAT_EMISSION) ? ME.radiance : zero;
      int lj = -1; Light L;
      if (norm(c) > BDPT_EPS_F) {
        for (int j = 0; j < sc.n_lights; j++) {
          L = load_light(sc, j);
          if (light_contain_point(L, E.pos)) { lj = j; break; }
        }
        if (lj < 0) c = zero;
        else { float pp, dp; c = light_sample_pdf(L, E.pos, -E.wo, &pp, &dp); }
      }
      contrib = E.alpha * c;
      if (norm(contrib) > BDPT_EPS_F && lj >= 0) {
        // MIS, eye chain only (bidirection.cpp:159-176,181-188)
        const float r_t = (1.0f / L.area) / E.D;
        float tail = 0.f;
        bool delta_prev = false;
        if (t >= 3) {
          Vtx P = V.get(0, t - 1);
          delta_prev = type_is_delta(mat_type<TR>(sc, P.mat));
          if (P.S > 0.f) {
            float pp, dp;
            light_sample_pdf(L, E.pos, -E.wo, &pp, &dp);  // dir pdf of leaving the light toward e_{t-1}
            f3 wv = P.pos - E.pos; float d2 = dot(wv, wv);
            f3 wd = wv * rsqrtf(d2);
            float gp = fabsf(dot(E.n, wd) * dot(wd, P.n)) / d2;
            float Rp = dp * gp / P.D;
            tail = Rp * Rp * P.S;
          }
        }
        double winv = 1.0 + (double)r_t * (double)r_t * ((double)(delta_prev ? 0.f : 1.f) + (double)tail);
        w = J.no_mis ? 1.f : (float)(1.0 / winv);
        if (!(w >= 0.f)) w = 0.f;
        ill = contrib * w;
        float* px = J.acc_eye + 3 * (size_t)pix_own;
        atomicAdd(px + 0, ill.x); atomicAdd(px + 1, ill.y); atomicAdd(px + 2, ill.z);
      }
    } else {
      // ---- light-side endpoint vl
      Vtx Lv; Mat ML;
      float ls_dir_pdf = 0.f;
      if (s == 1) {
        // fresh light point for this t (bidirection.cpp:332-358)
        Rng rng;
        rng.begin(J.seed, (unsigned long long)pix_own, (unsigned long long)sample, 2u + (uint32_t)t);
        int li = rng.pick(sc.n_lights);
        Light L = load_light(sc, li);
        float pA;
        f3 rad = light_sample_Le_point(L, E.pos, rng, &Lv.pos, &pA, &ls_dir_pdf, &Lv.n);
        pA /= (float)sc.n_lights;
        Lv.alpha = rad * (1.0f / pA);
        Lv.mat = -1; Lv.q = 1.f; Lv.D = pA; Lv.S = 1.f; Lv.wo = zero;
        ML.type = -1;
      } else {
        Lv = V.get(1, s);
        ML = load_mat<TR>(sc, Lv.mat);
      }
      // ---- eye-side endpoint ve (t == 1: the pinhole, bidirection.cpp:360-383)
      f3 ve_pos = E.pos, ve_n = E.n, eye_alpha = E.alpha;
      float cam_dir_pdf = 0.f;
      if (t == 1) {
        f3 wi; float dist;
        float We = camera_sample_ray_pdf(J.cam, Lv.pos, &wi, &dist, &cam_dir_pdf, &eye_x, &eye_y);
        ve_n = -wi;
        eye_alpha = mk3(We);
      }
      f3 cv = Lv.pos - ve_pos;
      const float dist2 = dot(cv, cv);
      const float dist = sqrtf(dist2);
      const f3 cdir = cv * (1.0f / dist);
      // local z's of the connection at both ends, and the dot products with the directions back along the subpaths
      const float cEz = dot(ve_n, cdir);               // t == 1: ve_n = cdir direction -> 1
      const float cLz = -dot(Lv.n, cdir);
      const float woLz = s >= 2 ? dot(Lv.n, Lv.wo) : 0.f;
      const float dotE = t >= 2 ? dot(E.wo, cdir) : 0.f;
      const float dotL = s >= 2 ? -dot(Lv.wo, cdir) : 0.f;
      // bsdf->f at both ends (bidirection.cpp:385-416)
      f3 f_eye = mk3(1.f), f_light = mk3(1.f);
      if (t >= 2) f_eye = bsdf_f_z(ME, woEz, cEz, dotE);
      if (s >= 2) f_light = bsdf_f_z(ML, cLz, woLz, dotL);
      const float g = fabsf(cLz * cEz) / dist2;
      contrib = eye_alpha * Lv.alpha * (f_eye * f_light) * g;
      bool push = false;
      unsigned int target = pix_own;
      // (for t = 1 the raster position is already known: a connection that leaves the frame never reaches the queue)
      const bool in_frame = t != 1 || (eye_x >= 0 && eye_y >= 0 && eye_x < J.cam.W && eye_y < J.cam.H);
      const bool gate = norm(contrib) > BDPT_EPS_F;
      const bool reserve = gate && in_frame;
      // The queue slot is reserved here, before the weight is known: the atomic's round trip to L2 (ncu put 23 % of
      // this kernel's stall samples on it) then overlaps the weight computation.  A connection that passes the gate
      // but ends up not pushed (weight 0 / NaN, or off-frame for t = 1: rare) fills its slot with a null ray.
      unsigned int r_active = 0u, r_mask = 0u, r_ticket = 0u;
      if (!TRACE) {
        r_active = __activemask();
        r_mask = __ballot_sync(r_active, reserve);
        if (r_mask != 0u && (int)(threadIdx.x & 31) == __ffs(r_mask) - 1) r_ticket = atomicAdd(J.qctr + 0, (unsigned int)__popc(r_mask));
      }
      if (gate && (TRACE || in_frame)) {   // (the trace hook records the weight of off-frame connections too)
        // ---- MIS weight (bidirection.cpp:121-293 in closed form, see header)
        double winv = 1.0;
        if (t >= 2) {
          float P_le = s == 1 ? ls_dir_pdf : bsdf_pdf_mis_z(ML, woLz, cLz, dotL) * Lv.q;
          float r_t = P_le * g / E.D;
          float S_t = E.S;
          if (ME.type == BDPT_MAT_MICROFACET && t >= 3) {
            // wo-dependent density: the chain factor next to the connection cannot come from the walk's cache
            Vtx P = V.get(0, t - 1);
            f3 wv = P.pos - E.pos; float d2 = dot(wv, wv);
            float gp = fabsf(woEz * dot(wv * rsqrtf(d2), P.n)) / d2;
            float R = bsdf_pdf_mis_z(ME, cEz, woEz, dotE) * E.q * gp / P.D;
            S_t = clampS((type_is_delta(mat_type<TR>(sc, P.mat)) ? 0.f : 1.f) + (P.S > 0.f ? R * R * P.S : 0.f));
          }
          winv += (double)r_t * (double)r_t * (double)S_t;
        }
        if (s >= 2) {
          float P_el = t == 1 ? cam_dir_pdf : bsdf_pdf_mis_z(ME, woEz, cEz, dotE) * E.q;
          float r_s = P_el * g / Lv.D;
          float S_s = Lv.S;
          if (ML.type == BDPT_MAT_MICROFACET) {
            Vtx P = V.get(1, s - 1);   // s - 1 == 1: the light origin record holds D = p_A, S = 1
            f3 wv = P.pos - Lv.pos; float d2 = dot(wv, wv);
            float gp = fabsf(woLz * dot(wv * rsqrtf(d2), P.n)) / d2;
            float R = bsdf_pdf_mis_z(ML, cLz, woLz, dotL) * Lv.q * gp / P.D;
            bool dprev = s - 1 >= 2 && type_is_delta(mat_type<TR>(sc, P.mat));
            S_s = clampS((dprev ? 0.f : 1.f) + (P.S > 0.f ? R * R * P.S : 0.f));
          }
          winv += (double)r_s * (double)r_s * (double)S_s;
        } else {
          // s == 1: the chain is evaluated at the ORIGINAL light-subpath origin
          float4 a, b;   // light_path[1]
          V.light_origin(&a, &b);
          f3 wv = mk3(a.x, a.y, a.z) - ve_pos;
          float d2 = dot(wv, wv);
          f3 wd = wv * rsqrtf(d2);
          const float cosA = dot(ve_n, wd);
          const float P_el = t == 1 ? cam_dir_pdf : bsdf_pdf_mis_z(ME, woEz, cosA, dot(E.wo, wd)) * E.q;
          float gp = fabsf(cosA * dot(wd, mk3(b.x, b.y, b.z))) / d2;
          float r_s = P_el * gp / b.w;
          winv += (double)r_s * (double)r_s;
        }
        w = J.no_mis ? 1.f : (float)(1.0 / winv);
        if (!(w >= 0.f)) w = 0.f;  // NaN guard
        ill = contrib * w;
        push = w > 0.f;
        if (t == 1) {
          if (eye_x >= 0 && eye_y >= 0 && eye_x < J.cam.W && eye_y < J.cam.H)
            target = (unsigned int)(eye_y * J.cam.W + eye_x) | 0x80000000u;
          else push = false;
        }
      }
      if (TRACE) {
        // trace mode resolves visibility inline so that the record says "occluded" (NaN) like the oracle's
        Hit h = traverse<true, TR>(sc, ve_pos, cdir, BDPT_EPS_F, dist - BDPT_EPS_F);
        if (h.leaf >= 0) { contrib = mk3(NAN); w = NAN; push = false; eye_x = eye_y = -1; }
      }
      if (TRACE) {
        push_shadow(J, push, ve_pos, cdir, dist - BDPT_EPS_F, target, ill);
      } else if (r_mask != 0u) {
        const unsigned int lane = threadIdx.x & 31;
        const unsigned int base = __shfl_sync(r_active, r_ticket, __ffs(r_mask) - 1);
        if (reserve) {
          const unsigned int idx = base + __popc(r_mask & ((1u << lane) - 1u));
          if (idx >= J.qcap) J.qctr[2] = 1u;
          else {
            const float tmax = push ? dist - BDPT_EPS_F : -1.f;     // null ray: nothing to hit, adds zero
            const f3 add = push ? ill : zero;
            st256(J.q0 + 2 * (size_t)idx, make_float4(ve_pos.x, ve_pos.y, ve_pos.z, tmax),
                  make_float4(cdir.x, cdir.y, cdir.z, __uint_as_float(push ? target : pix_own)));
            J.q2[idx] = make_float4(add.x, add.y, add.z, 0.f);
          }
        }
      }
    }
    if (TRACE && J.dbg) {
      double* r = J.dbg + 8 * ((size_t)(t - 1) * (size_t)(J.m + 2) + (size_t)s);
      r[0] = t; r[1] = s; r[2] = contrib.x; r[3] = contrib.y; r[4] = contrib.z; r[5] = w; r[6] = eye_x; r[7] = eye_y;
    }
  }
  unsigned int tot = __reduce_add_sync(__activemask(), nconn);
  if ((threadIdx.x & 31) == (__ffs(__activemask()) - 1)) atomicAdd(J.stats + 2, (unsigned long long)tot);
}

// One thread per (sample, t), vertices read from HBM/L2 as needed.  Grid: 1-D, the (m+1) values of t of
// one group of 128 samples are consecutive blocks, so the light vertices they share meet in L2.
template <bool TRACE, int MINB, int TR = 0>
__global__ void __launch_bounds__(128, MINB) k_connect(const Job J) {
  const int t = (int)(blockIdx.x % (unsigned)(J.m + 1)) + 1;
  const int si = (int)(blockIdx.x / (unsigned)(J.m + 1)) * blockDim.x + threadIdx.x;
  if (si >= J.B) return;
  const int nE = J.counts[si];
  const int nL = J.counts[J.B + si];
  if (nE == 0 || t > nE) return;  // invalid lane or the eye subpath is shorter than t
  GlobalVerts V{J, si};
  connect_strategies<TRACE, TR>(J, si, t, nL, V);
}

// -------------------------------------------------------------------------------------- shadow
// Adds unoccluded connections to their pixels.  Called by a converged set of lanes (`mask`, every lane of it holds
// queue entry idx).  In scenes with a point light (J.point_lights) the t=1, s=1 connection of EVERY sample is the same
// ray and lands on the same pixel of the light image; 38 M same-address atomics per batch took 57 of CBempty's
// 60 ms.  There, lanes of the set that splat the same light-image pixel are summed first and their leader issues the
// three atomics.
__device__ inline void shadow_splat(const Job& J, unsigned mask, unsigned idx) {
  const float4 c = J.q2[idx];
  const unsigned int tg = __float_as_uint(J.q1[2 * (size_t)idx].w);
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
  float* px = ((tg & 0x80000000u) ? J.acc_light : J.acc_eye) + 3 * (size_t)(tg & 0x7fffffffu);
  atomicAdd(px + 0, sx); atomicAdd(px + 1, sy); atomicAdd(px + 2, sz);
}

template <int TR = 0>
__global__ void __launch_bounds__(128) k_shadow(const Job J) {
  const unsigned int n = min(J.qctr[0], J.qcap);
  const int lane = threadIdx.x & 31;
  unsigned int traced = 0;
  for (;;) {
    unsigned int base = 0;
    if (lane == 0) base = atomicAdd(J.qctr + 1, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= n) break;
    unsigned int idx = base + lane;
    if (idx < n) {
      float4 a, b; ld256(J.q0 + 2 * (size_t)idx, &a, &b);
      const bool real = a.w >= 0.f;   // tmax < 0: a reserved slot whose connection turned out to carry nothing
      Hit h; h.leaf = 0;
      if (real) { h = traverse<true, TR>(J.sc, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), BDPT_EPS_F, a.w); traced++; }
      const unsigned clear = __ballot_sync(__activemask(), h.leaf < 0);
      if (h.leaf < 0) shadow_splat(J, clear, idx);
    }
  }
  unsigned int tot = __reduce_add_sync(0xffffffffu, traced);
  if (lane == 0 && tot) atomicAdd(J.stats + 1, (unsigned long long)tot);
}

// ------------------------------------------------------------------- persistent (scheduled) kernels
// The two traversal kernels above give every lane one ray (or one subpath) and wait for the
// slowest lane of the warp: ncu shows 6.6 (k_walk) and 9.0 (k_shadow) active lanes per instruction
// (profiles/r01_*).  The persistent forms keep a lane busy: warps claim work in chunks with one
// atomic per chunk, hand indices to idle lanes with ballot + popc prefix, and run the resumable
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

// Copies the first `ntop` BVH nodes (the upper levels: nodes are stored breadth first) into shared memory.
__device__ inline void stage_top_nodes(const Job& J, uint4* top) {
  for (int i = threadIdx.x; i < J.ntop * 5; i += blockDim.x) top[i] = J.sc.nodes[i];
  __syncthreads();
}

// FETCH: idle lanes a warp tolerates before it fetches (measured 4: 6.53, 8: 6.31, 12: 6.22, 16: 6.24 ms per 2^24
// samples); SPLAT: parked rays it collects before it splats (8: 6.62, 16: 6.53, 24: 6.70 at FETCH = 4).
template <int MINB, int TR = 0, int FETCH = 12, int SPLAT = 16>
__global__ void __launch_bounds__(128, MINB) k_shadow_pt(const Job J) {
  extern __shared__ uint4 top[];
  stage_top_nodes(J, top);
  const unsigned n = min(J.qctr[0], J.qcap);
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
      unsigned my = warp_fetch(J.qctr + 1, n, 256u, !active, wf);
      if (my != BDPT_NONE) {
        float4 a, b; ld256(J.q0 + 2 * (size_t)my, &a, &b);
        if (a.w >= 0.f) {   // tmax < 0: a reserved slot whose connection turned out to carry nothing
          T.begin(mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), BDPT_EPS_F, a.w);
          idx = my; active = true;
        }
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

// Per-lane subpath state of the persistent walk.
struct WalkState {
  Rng rng;
  f3 prev_n, alpha, prev_f;
  float prev_pdf, pdfeval_prev, D_prev, S_prev;
  int kind, si, j, nv;
  bool delta_prev;
};

// New subpath: pinhole ray (raytrace_pixel, bidirection.cpp:517-524) or light emission
// (sample_light_ray, :105-118).  Returns false if the index maps to no pixel.
__device__ inline bool walk_begin(const Job& J, unsigned i, WalkState& W, Trav& T) {
  W.kind = i >= (unsigned)J.B ? 1 : 0;
  W.si = (int)i - W.kind * J.B;
  int x, y, sample;
  if (!decode_sample(J, J.g0 + (unsigned long long)W.si, &x, &y, &sample)) { J.counts[i] = 0; return false; }
  W.rng.begin(J.seed, (unsigned long long)y * (unsigned long long)J.cam.W + (unsigned long long)x,
              (unsigned long long)sample, (uint32_t)W.kind);
  W.prev_f = mk3(1.f); W.delta_prev = false; W.nv = 1; W.j = 2;
  f3 o, d;
  if (W.kind == 0) {
    float sx = W.rng.uniform(), sy = W.rng.uniform();
    camera_generate_ray(J.cam, ((float)x + sx) / (float)J.cam.W, ((float)y + sy) / (float)J.cam.H, &o, &d);
    W.prev_n = d; W.alpha = mk3(1.f); W.prev_pdf = 1.f;
    W.pdfeval_prev = 1.f; W.D_prev = 1.f; W.S_prev = 0.f;
    T.begin(o, d, J.cam.nclip, J.cam.fclip);
  } else {
    int li = W.rng.pick(J.sc.n_lights);
    Light L = load_light(J.sc, li);
    float pA, pdir; f3 nl;
    f3 Le = light_sample_Le(L, W.rng, &o, &d, &pA, &pdir, &nl);
    pA /= (float)J.sc.n_lights;
    W.prev_n = nl; W.alpha = Le * (1.0f / pA); W.prev_pdf = pdir;
    W.pdfeval_prev = pdir; W.D_prev = pA; W.S_prev = 1.f;
    size_t s1 = VSLOT(J, 1, 1, W.si);
    store_vertex(J, s1, make_float4(o.x, o.y, o.z, __int_as_float(-1)),
                 make_float4(nl.x, nl.y, nl.z, pA),
                 make_float4(W.alpha.x, W.alpha.y, W.alpha.z, 1.f),
                 make_float4(pdir, 0.f, 0.f, 1.f));
    T.begin(o, d, BDPT_EPS_F, INFINITY);
  }
  return true;
}

// The traversal of the current ray finished: store the vertex (if any) and sample the next ray
// (body of the while loop of prepare_bidirectional_subpath, bidirection.cpp:44-101).
// Returns true if the subpath continues (T holds the next ray), false if it ended.
template <int TR = 0>
__device__ inline bool walk_shade(const Job& J, WalkState& W, Trav& T) {
  const DScene& sc = J.sc;
  if (T.hit.leaf < 0) return false;
  const f3 o = T.o, d = T.d;
  f3 n, hit_p; int prim, mat;
  finish_hit<TR>(sc, T.hit, o, d, &n, &prim, &mat, &hit_p);
  const float t = T.hit.t;
  const Frame fr = make_frame(n);
  const f3 w_out = fr.to_local(-d);
  const Mat M = load_mat<TR>(sc, mat);
  f3 wi; float pdf;
  const f3 f = bsdf_sample_f(M, w_out, &wi, &pdf, W.rng);
  const f3 wi_world = normalized(fr.to_world(wi));

  const float cos_prev = fabsf(dot(W.prev_n, d));
  const float g = cos_prev * fabsf(dot(n, d)) / (t * t);
  W.alpha = W.alpha * W.prev_f * (cos_prev / W.prev_pdf);

  float p_keep = 1.f;
  if (J.roulette && W.j > 2) p_keep = fminf(1.f, norm(f) / pdf);

  const bool delta_j = type_is_delta(M.type);
  const float D_j = (W.kind == 0 && W.j == 2) ? 1.f : W.pdfeval_prev * g;
  const f3 zero = mk3(0.f);
  const float D_rev = bsdf_pdf_mis(M, wi, w_out) * p_keep * g;
  float S_j = (delta_j || W.delta_prev) ? 0.f : 1.f;
  if (W.S_prev > 0.f) { float R = D_rev / W.D_prev; S_j += R * R * W.S_prev; }
  S_j = clampS(S_j);

  size_t sl = VSLOT(J, W.kind, W.j, W.si);
  store_vertex(J, sl, make_float4(hit_p.x, hit_p.y, hit_p.z, __int_as_float(mat)),
               make_float4(n.x, n.y, n.z, D_j),
               make_float4(W.alpha.x, W.alpha.y, W.alpha.z, S_j),
               make_float4(-d.x, -d.y, -d.z, p_keep));
  W.nv = W.j;
  if (W.j >= J.m + 1) return false;
  if (J.roulette && !(W.rng.uniform() < p_keep)) return false;
  if (is_zero(f)) return false;

  W.pdfeval_prev = bsdf_pdf_mis(M, w_out, wi) * p_keep;
  W.prev_f = f; W.prev_n = n; W.prev_pdf = pdf * p_keep;
  W.D_prev = D_j; W.S_prev = S_j; W.delta_prev = delta_j;
  W.j++;
  T.begin(hit_p, wi_world, BDPT_EPS_F, INFINITY);
  return true;
}

// phase: 0 = empty (wants a subpath), 1 = traversing, 2 = traversal finished (wants shading)
template <int SHADE_THRESHOLD, int FETCH_THRESHOLD, int MINB = 4, int TR = 0>
__global__ void __launch_bounds__(128, MINB) k_walk_pt(const Job J) {
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
        else { J.counts[W.kind * J.B + W.si] = W.nv; phase = 0; }
      }
    }
    // ... and so does starting new subpaths
    const unsigned empty = __ballot_sync(0xffffffffu, phase == 0);
    if (empty != 0u && !wf.exhausted && (__popc(empty) >= FETCH_THRESHOLD || __ballot_sync(0xffffffffu, phase == 1) == 0u)) {
      unsigned my = warp_fetch(J.qctr + 3, total, 128u, phase == 0, wf);
      if (my != BDPT_NONE && walk_begin(J, my, W, T)) phase = 1;
    }
    if (__ballot_sync(0xffffffffu, phase != 0) == 0u && wf.exhausted) break;
    if (phase == 1) {
      if (T.step<false, TR>(J.sc, stack, top, (uint32_t)J.ntop)) { phase = 2; nrays++; }
    }
  }
  unsigned int tot = __reduce_add_sync(0xffffffffu, nrays);
  if ((threadIdx.x & 31) == 0 && tot) atomicAdd(J.stats + 0, (unsigned long long)tot);
}

// ------------------------------------------------------------------- wavefront walk (per bounce)
// The walk split at the ray boundary: k_wf_gen writes every subpath's first ray; then, per bounce,
// k_wf_trace (persistent, dynamic fetch, pure closest-hit traversal: few registers, small code)
// fills a hit record per live subpath and k_wf_shade (one thread per live subpath, full warps)
// stores the vertex, samples the BSDF and compacts the survivors into the next bounce's queue
// with ballot + prefix.  Subpath state lives in SoA planes indexed by subpath id:
//   r0 = {o, tmin}  r1 = {d, tmax}  h = {t, leaf, b1, b2}
//   s0 = rng {state, inc}  s1 = {alpha, prev_pdf}  s2 = {prev_f, pdfeval_prev}
//   s3 = {prev_n, D_prev}  s4 = {S_prev, j | delta_prev << 8, nv, -}
struct WfBuf {
  float4 *r0, *r1, *h, *s0, *s1, *s2, *s3, *s4;
  unsigned int *queue_a, *queue_b;   // subpath ids alive at this / the next bounce
  unsigned int* ctr;                 // this bounce: [0] n_in  [1] trace cursor; [4] = n_in of the next bounce
};

__device__ inline void wf_store_state(const WfBuf& F, unsigned i, const WalkState& W) {
  ulonglong2 r; r.x = W.rng.state; r.y = W.rng.inc;
  *reinterpret_cast<ulonglong2*>(F.s0 + i) = r;
  F.s1[i] = make_float4(W.alpha.x, W.alpha.y, W.alpha.z, W.prev_pdf);
  F.s2[i] = make_float4(W.prev_f.x, W.prev_f.y, W.prev_f.z, W.pdfeval_prev);
  F.s3[i] = make_float4(W.prev_n.x, W.prev_n.y, W.prev_n.z, W.D_prev);
  F.s4[i] = make_float4(W.S_prev, __int_as_float(W.j | (W.delta_prev ? 256 : 0)), __int_as_float(W.nv), 0.f);
}

__device__ inline void wf_load_state(const WfBuf& F, unsigned i, int B, WalkState& W) {
  ulonglong2 r = *reinterpret_cast<const ulonglong2*>(F.s0 + i);
  W.rng.state = r.x; W.rng.inc = r.y;
  float4 a = F.s1[i], b = F.s2[i], c = F.s3[i], d = F.s4[i];
  W.alpha = mk3(a.x, a.y, a.z); W.prev_pdf = a.w;
  W.prev_f = mk3(b.x, b.y, b.z); W.pdfeval_prev = b.w;
  W.prev_n = mk3(c.x, c.y, c.z); W.D_prev = c.w;
  W.S_prev = d.x;
  int jj = __float_as_int(d.y);
  W.j = jj & 255; W.delta_prev = (jj & 256) != 0; W.nv = __float_as_int(d.z);
  W.kind = i >= (unsigned)B ? 1 : 0;
  W.si = (int)i - W.kind * B;
}

__global__ void __launch_bounds__(128) k_wf_gen(const Job J, const WfBuf F) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2u * (unsigned)J.B) return;
  WalkState W; Trav T;
  bool ok = walk_begin(J, i, W, T);
  // first bounce: the queue is the identity; dead (padding) lanes get a ray that cannot hit
  F.r0[i] = make_float4(T.o.x, T.o.y, T.o.z, ok ? T.tmin : 1.f);
  F.r1[i] = make_float4(T.d.x, T.d.y, T.d.z, ok ? T.hit.t : 0.f);
  if (ok) wf_store_state(F, i, W);
  else F.s4[i] = make_float4(0.f, __int_as_float(0), __int_as_float(0), 0.f);
}

// n_in = F.ctr[0] (or `fixed_n` for the first bounce, where the queue is the identity)
__global__ void __launch_bounds__(128) k_wf_trace(const Job J, const WfBuf F, const unsigned fixed_n) {
  const unsigned n = fixed_n ? fixed_n : F.ctr[0];
  Trav T;
  uint2 stack[BDPT_STACK];
  bool active = false;
  unsigned id = 0, traced = 0;
  WarpFetch wf; wf.next = 0; wf.end = 0; wf.exhausted = false;
#pragma unroll 1
  for (;;) {
    const unsigned idle = __ballot_sync(0xffffffffu, !active);
    if (!wf.exhausted && (__popc(idle) >= 4 || idle == 0xffffffffu)) {
      unsigned my = warp_fetch(F.ctr + 1, n, 256u, !active, wf);
      if (my != BDPT_NONE) {
        id = fixed_n ? my : F.queue_a[my];
        float4 a = F.r0[id], b = F.r1[id];
        T.begin(mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), a.w, b.w);
        active = true;
      }
    }
    if (__ballot_sync(0xffffffffu, active) == 0u) {
      if (wf.exhausted) break;
      continue;
    }
    if (active && T.step<false>(J.sc, stack)) {
      traced++;
      active = false;
      F.h[id] = make_float4(T.hit.t, __int_as_float(T.hit.leaf), T.hit.b1, T.hit.b2);
    }
  }
  unsigned int tot = __reduce_add_sync(0xffffffffu, traced);
  if ((threadIdx.x & 31) == 0 && tot) atomicAdd(J.stats + 0, (unsigned long long)tot);
}

__global__ void __launch_bounds__(128) k_wf_shade(const Job J, const WfBuf F, const unsigned fixed_n) {
  const unsigned n = fixed_n ? fixed_n : F.ctr[0];
  const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
  bool cont = false;
  unsigned id = 0;
  if (q < n) {
    id = fixed_n ? q : F.queue_a[q];
    WalkState W;
    wf_load_state(F, id, J.B, W);
    if (W.j != 0) {   // j == 0 marks a padding lane of the first bounce
      Trav T;
      float4 a = F.r0[id], b = F.r1[id], h = F.h[id];
      T.o = mk3(a.x, a.y, a.z); T.d = mk3(b.x, b.y, b.z);
      T.hit.t = h.x; T.hit.leaf = __float_as_int(h.y); T.hit.b1 = h.z; T.hit.b2 = h.w;
      cont = walk_shade(J, W, T);
      if (cont) {
        F.r0[id] = make_float4(T.o.x, T.o.y, T.o.z, T.tmin);
        F.r1[id] = make_float4(T.d.x, T.d.y, T.d.z, T.hit.t);
        wf_store_state(F, id, W);
      } else {
        J.counts[W.kind * J.B + W.si] = W.nv;
      }
    }
  }
  // compact the survivors into the next bounce's queue
  const unsigned mask = __ballot_sync(0xffffffffu, cont);
  if (mask) {
    const int lane = threadIdx.x & 31;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(F.ctr + 4, (unsigned)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (cont) F.queue_b[base + __popc(mask & ((1u << lane) - 1u))] = id;
  }
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
  f3 le = light_sample_pdf(L, p, wi, &pp, &dp);
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
__global__ void k_tonemap(const float* eye, const float* light, int which, float inv_ns, int W, int H, uint8_t* rgba) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= W * H) return;
  int x = p % W, y = p / W;
  float v[3];
  for (int c = 0; c < 3; c++) {
    float e = eye[3 * p + c], l = light[3 * p + c];
    v[c] = (which == 0 ? e + l : which == 1 ? e : l) * inv_ns;
  }
  uint8_t* o = rgba + 4 * ((size_t)(H - 1 - y) * W + x);
  for (int c = 0; c < 3; c++) {
    float g = powf(fmaxf(v[c], 0.f) * 1.41421356f, 1.0f / 2.2f);
    g = fminf(fmaxf(g, 0.f), 1.f);
    o[c] = (uint8_t)(g * 255.f);
  }
  o[3] = 255;
}

__global__ void k_resolve(const float* eye, const float* light, int which, float inv_ns, size_t n, float* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float e = eye[i], l = light[i];
  out[i] = (which == 0 ? e + l : which == 1 ? e : l) * inv_ns;
}

}  // namespace bdpt
