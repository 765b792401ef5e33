// common.cuh — fp32 vector helpers, constants and the device-side scene tables.
// Written for sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bdpt {

// reference CGL/include/CGL/misc.h:11-15
#define BDPT_PI 3.14159265358979323f
#define BDPT_INV_PI 0.318309886183790672f
#define BDPT_EPS_F 0.00001f

struct f3 {
  float x, y, z;
};

__host__ __device__ inline f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__host__ __device__ inline f3 mk3(float c) { return mk3(c, c, c); }
__host__ __device__ inline f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ inline f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ inline f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
__host__ __device__ inline f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__host__ __device__ inline f3 operator*(f3 a, float c) { return mk3(a.x * c, a.y * c, a.z * c); }
__host__ __device__ inline f3 operator*(float c, f3 a) { return mk3(a.x * c, a.y * c, a.z * c); }
__host__ __device__ inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ inline f3 cross(f3 u, f3 v) {
  return mk3(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
}
__device__ inline float norm2(f3 a) { return dot(a, a); }
__device__ inline float norm(f3 a) { return sqrtf(dot(a, a)); }
// rsqrtf (MUFU.RSQ, <= 2 ulp): the IEEE sqrt + divide form was ~10 % of k_connect's issued instructions
__device__ inline f3 normalized(f3 a) { return a * rsqrtf(dot(a, a)); }
__device__ inline bool is_zero(f3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }

// Orthonormal frame around a normal; same axis choice as the reference's make_coord_space
// (src/pathtracer/bsdf.cpp:21-41) so that a local direction maps to the same world direction.
struct Frame {
  f3 x, y, z;
  __device__ inline f3 to_world(f3 v) const { return v.x * x + v.y * y + v.z * z; }
  __device__ inline f3 to_local(f3 v) const { return mk3(dot(x, v), dot(y, v), dot(z, v)); }
};

__device__ inline Frame make_frame(f3 n) {
  f3 h = n;
  float ax = fabsf(h.x), ay = fabsf(h.y), az = fabsf(h.z);
  if (ax <= ay && ax <= az) h.x = 1.0f;
  else if (ay <= ax && ay <= az) h.y = 1.0f;
  else h.z = 1.0f;
  Frame f;
  f.z = normalized(n);
  f.y = normalized(cross(h, f.z));
  f.x = normalized(cross(f.z, f.y));
  return f;
}

// ------------------------------------------------------------------ device scene tables
// Material record: 5 x float4 (80 B), read with 16-byte loads.
//   q0 = {reflectance.rgb, as_float(type)}
//   q1 = {transmittance.rgb, ior}
//   q2 = {radiance.rgb, alpha}
//   q3 = {eta.rgb, 0}
//   q4 = {k.rgb, 0}
struct DMaterial {
  float4 q[5];
};

// Light record: 5 x float4.
//   q0 = {radiance.rgb, as_float(type)}
//   q1 = {position.xyz, area}
//   q2 = {direction.xyz (unit, as stored by the loader), 0}
//   q3 = {dim_x.xyz, 0}
//   q4 = {dim_y.xyz, 0}
// environment (BDPT_LIGHT_ENV): q1 = {centre of the far sphere, its area 4 pi R^2}, q3 = {R, scene bounding radius, 0, 0}
struct DLight {
  float4 q[5];
};

struct DCamera {
  f3 pos;
  f3 c2w_x, c2w_y, c2w_z;  // columns of c2w
  f3 w2c_r0, w2c_r1, w2c_r2;  // rows of w2c (= inverse of c2w)
  float tan_h, tan_v;  // tan(hFov/2), tan(vFov/2)   (camera.cpp:200-201)
  float nclip, fclip;
  int W, H;
  // thin lens (camera.h lensRadius / focalDistance, -b / -d).  The reference's BDPT never reads them:
  // lens_radius > 0 switches on the extension of shading.cuh ("thin lens")
  float lens_radius, focal_dist;
};

struct alignas(16) Node8 {  // 80 bytes, see bvh8.h
  uint4 q[5];
};

// Scene traits the hot kernels are specialised on (decided once per scene at upload, api.cu).  The walk kernel is
// bound by instruction fetch (71 KB of SASS when every material and both primitive kinds are compiled in); most
// scenes of the reference are triangle meshes with Lambertian surfaces, and the kernels built for those traits are
// half the size.  A trait is a promise about the scene's contents, never a change of behaviour.
#define BDPT_TRAIT_TRIS 1      // every primitive is a triangle
#define BDPT_TRAIT_LAMBERT 2   // every material is DiffuseBSDF or EmissionBSDF
#define BDPT_TRAIT_NOMF 4      // no MicrofacetBSDF (implied by LAMBERT; used when LAMBERT does not hold)
// ... and one that ADDS code: the thin-lens extension (bdpt_set_lens) lives only in the builds that carry this bit —
// the generic build plus the lens — so that the pinhole renders the reference ships pay nothing for it (inlined into
// every build it cost the walk kernel 2 %).
#define BDPT_TRAIT_LENS 8

#ifndef BDPT_NODE_Q
#define BDPT_NODE_Q 6   // uint4 per node in the traversal array (the builders emit 5; the sixth is padding): shadow kernel -2.6 %
#endif
#ifndef BDPT_PRIM_Q
#define BDPT_PRIM_Q 3   // float4 per leaf primitive record (4 = padded to 64 bytes, first two rows in one 32-byte load: measured
                        // slower, walk +2 %, shadow +3.9 % — the records' footprint grows by a third)
#endif

struct DScene {
  const uint4* nodes;       // Node8 (5 x uint4) at a stride of BDPT_NODE_Q uint4: 96 bytes, three 32-byte loads
  const float4* prims;      // leaf-ordered primitive records, BDPT_PRIM_Q float4 each (layouts: bvh_gpu.cuh, make_leaf_record)
  const float4* shading;    // leaf-ordered shading normals n1,n2,n3, 3 x float4 each; .w: primitive id, material, -
  const double* prims64;    // leaf-ordered fp64 geometry, 10 doubles each: triangle p1 p2 p3 (+pad) / sphere c, r
  const float4* materials;  // DMaterial as 5 x float4
  const float4* lights;     // DLight as 5 x float4
  const float4* envmap;     // lat-long radiance map of the environment light (row 0 = up) or null: constant
  int env_w, env_h;
  int env_light;            // index of the environment light, -1 = none; its sphere carries material n_materials
  int n_lights;
  int n_materials;
  int n_nodes;
  int n_prims;
  // 0x47000000 (the float 2^15).  Lives here so that the traversal's byte decode reads it from the kernel's
  // constant bank: PRMT takes one immediate, and that slot goes to the byte selector (traverse.cuh, qbyte).
  uint32_t qbias_bits;
};

// Division of a 32-bit unsigned by a divisor fixed for the whole launch (Granlund & Montgomery 1994, the branch-free
// form): q = (t + ((n - t) >> s1)) >> s2 with t = umulhi(m, n); exact for every n < 2^32.  The hardware has no integer
// divide, so `n / d` is ~25 instructions; this is 4.
struct FastDiv {
  uint32_t d, m, s1, s2;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f; f.d = d;
  uint32_t L = 0; while (L < 32 && ((uint64_t)1 << L) < d) L++;
  f.m = (uint32_t)((((uint64_t)1 << 32) * (((uint64_t)1 << L) - d)) / d + 1);
  f.s1 = L < 1 ? L : 1; f.s2 = L > 1 ? L - 1 : 0;
  return f;
}
__host__ __device__ inline uint32_t fast_div(uint32_t n, const FastDiv& f) {
#ifdef __CUDA_ARCH__
  const uint32_t t = __umulhi(f.m, n);
#else
  const uint32_t t = (uint32_t)(((uint64_t)f.m * n) >> 32);
#endif
  return (t + ((n - t) >> f.s1)) >> f.s2;
}

// 1/x by the SFU alone (MUFU.RCP, <= 1 ulp on normal inputs; denormals flush): __fdividef(1, x) adds a range fix-up
__device__ inline float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// 32-byte read-only loads (LDG.E.256.CONSTANT, sm_100): one L1 wavefront per lane where two 16-byte loads take two — the
// traversal kernels run the L1 data pipe at 62-77 % of its wavefront peak (profiles/README.md, r02d)
// (BDPT_HINT_SCENE: an L2 qualifier for the scene's records, measured in DESIGN.md §4)
#ifndef BDPT_HINT_SCENE
#define BDPT_HINT_SCENE ""
#endif
__device__ inline void ld256nc(const uint4* p, uint4* a, uint4* b) {
  asm volatile("ld.global.nc" BDPT_HINT_SCENE ".v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a->x), "=r"(a->y), "=r"(a->z), "=r"(a->w), "=r"(b->x), "=r"(b->y), "=r"(b->z), "=r"(b->w) : "l"(p));
}
__device__ inline void ld256nc(const float4* p, float4* a, float4* b) {
  asm volatile("ld.global.nc" BDPT_HINT_SCENE ".v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a->x), "=f"(a->y), "=f"(a->z), "=f"(a->w), "=f"(b->x), "=f"(b->y), "=f"(b->z), "=f"(b->w) : "l"(p));
}
__device__ inline float4 ldg4(const float4* p) { return __ldg(p); }
__device__ inline uint4 ldg4(const uint4* p) { return __ldg(p); }

}  // namespace bdpt
