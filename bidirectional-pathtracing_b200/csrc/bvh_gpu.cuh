// bvh_gpu.cuh — the wide BVH built ON THE DEVICE, in place of the reference's host-side
// BVHAccel::construct_bvh (src/scene/bvh.cpp:51-129, called from RaytracedRenderer::build_accel,
// raytraced_renderer.cpp:350-374).  Steps (bvh_ploc.h has the algorithm, bvh8_wide.h the collapse):
//   k_bvh_bounds      fp32 bounds of every primitive as the kernels test it + bounds of the centroids
//   k_bvh_morton      63-bit Morton keys;  cub::DeviceRadixSort sorts (key, primitive)
//   k_bvh_leaves      leaf nodes in Morton order = the first cluster array
//   k_ploc_nearest / k_ploc_merge / k_ploc_scan / k_ploc_scatter   one PLOC round (order-preserving compaction)
//   k_bvh_collapse_dp bottom-up SAH-optimal cut into 8-wide nodes (second thread to arrive at a node does it)
//   k_bvh_emit_level  one level of wide nodes per launch: slots, quantisation, leaf order, next level's items
//   k_bvh_records     leaf-ordered primitive / shading-normal / fp64 geometry records from the raw scene
// Everything runs on the context's stream; the host reads back two counters per few rounds / per level.
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include "../../include/bdpt_flatscene.h"
#include "bvh_ploc.h"
#include "common.cuh"

namespace bdpt {

struct DevTree {           // binary tree as flat arrays; nodes [0, N) are the leaves in Morton order
  float4 *lo, *hi;
  int *left, *right, *parent, *count;
  uint32_t* leaf_prim;
};

struct DevTreeView {       // tree view for bvh8_wide.h
  DevTree t;
  __device__ void box(int n, float* lo, float* hi) const {
    const float4 a = t.lo[n], b = t.hi[n];
    lo[0] = a.x; lo[1] = a.y; lo[2] = a.z; hi[0] = b.x; hi[1] = b.y; hi[2] = b.z;
  }
  __device__ int left(int n) const { return t.left[n]; }
  __device__ int right(int n) const { return t.right[n]; }
  __device__ int count(int n) const { return t.count[n]; }
  __device__ void gather(int n, uint32_t* out) const {   // at most 3 primitives: a subtree of at most 5 nodes
    int st[4], sp = 0, k = 0;
    st[sp++] = n;
    while (sp > 0) {
      const int c = st[--sp];
      if (t.left[c] < 0) out[k++] = t.leaf_prim[c];
      else { st[sp++] = t.right[c]; st[sp++] = t.left[c]; }
    }
  }
};

// order-preserving float <-> uint for atomicMin / atomicMax
__device__ inline unsigned bvh_fenc(float f) { unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ inline float bvh_fdec(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// fp32 bounds of what the kernels actually test: the fp32-narrowed triangle, a slightly padded sphere (api.cu,
// make_build_prims does the same on the host).  cb: {min xyz, max xyz} of the centroids, encoded (init: ~0 / 0).
__global__ void k_bvh_bounds(const BdptPrim* prims, int n, float4* blo, float4* bhi, unsigned* cb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  if (i < n) {
    const double* d = prims[i].data;
    if (prims[i].type == BDPT_PRIM_TRIANGLE) {
      for (int a = 0; a < 3; a++) {
        const float v0 = (float)d[a], v1 = (float)d[3 + a], v2 = (float)d[6 + a];
        lo[a] = fminf(v0, fminf(v1, v2)); hi[a] = fmaxf(v0, fmaxf(v1, v2));
      }
    } else {
      const double r = fabs(d[3]) * (1.0 + 1e-6) + 1e-7;
      for (int a = 0; a < 3; a++) { lo[a] = __double2float_rd(d[a] - r); hi[a] = __double2float_ru(d[a] + r); }
    }
    blo[i] = make_float4(lo[0], lo[1], lo[2], 0.f); bhi[i] = make_float4(hi[0], hi[1], hi[2], 0.f);
  }
  // centroid bounds: warp reduce, one atomic per warp and component
  for (int a = 0; a < 3; a++) {
    float cmin = i < n ? 0.5f * (lo[a] + hi[a]) : FLT_MAX, cmax = i < n ? 0.5f * (lo[a] + hi[a]) : -FLT_MAX;
    for (int o = 16; o > 0; o >>= 1) { cmin = fminf(cmin, __shfl_xor_sync(0xffffffffu, cmin, o)); cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(cb + a, bvh_fenc(cmin)); atomicMax(cb + 3 + a, bvh_fenc(cmax)); }
  }
}

__global__ void k_bvh_morton(const float4* blo, const float4* bhi, int n, const unsigned* cb, unsigned long long* keys, uint32_t* vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 a = blo[i], b = bhi[i];
  const float c[3] = {0.5f * (a.x + b.x), 0.5f * (a.y + b.y), 0.5f * (a.z + b.z)};
  const float clo[3] = {bvh_fdec(cb[0]), bvh_fdec(cb[1]), bvh_fdec(cb[2])}, chi[3] = {bvh_fdec(cb[3]), bvh_fdec(cb[4]), bvh_fdec(cb[5])};
  keys[i] = ploc_morton(c, clo, chi);
  vals[i] = (uint32_t)i;
}

// leaf k = the primitive with the k-th smallest key; the first cluster array is the identity
__global__ void k_bvh_leaves(const float4* blo, const float4* bhi, const uint32_t* sorted, int n, DevTree T, int* cluster, float4* clo, float4* chi) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t p = sorted[k];
  const float4 a = blo[p], b = bhi[p];
  T.lo[k] = a; T.hi[k] = b; T.left[k] = -1; T.right[k] = -1; T.parent[k] = -1; T.count[k] = 1; T.leaf_prim[k] = p;
  cluster[k] = k; clo[k] = a; chi[k] = b;
}

// ---- one PLOC round.  ctr: [0] clusters now  [1] clusters after this round  [2] next free node id  [3] rounds of the tail
// kernel  [4..6] the emission's counters  [8] clusters at the start of the round (for its scatter)
__global__ void k_ploc_nearest(const unsigned* ctr, const float4* clo, const float4* chi, int* nn, int radius) {
  const int n = (int)ctr[0];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || n < 2) return;
  nn[i] = ploc_nearest(i, n, (const float*)clo, (const float*)chi, radius);
}

#define BDPT_PLOC_BLOCK 256
// mutual nearest neighbours merge (the lower position keeps the new cluster, the higher one leaves the array);
// out_id / out boxes: what each position holds after the round, keep: whether it stays; bcount: survivors per block
__global__ void __launch_bounds__(BDPT_PLOC_BLOCK) k_ploc_merge(unsigned* ctr, const int* nn, const int* cluster, const float4* clo, const float4* chi,
                                                                 DevTree T, int* out_id, float4* olo, float4* ohi, unsigned char* keep, unsigned* bcount) {
  const int n = (int)ctr[0];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool stay = false;
  if (i < n && n >= 2) {
    const int j = nn[i];
    const bool mutual = nn[j] == i;
    stay = !(mutual && i > j);
    if (stay) {
      int id = cluster[i];
      float4 lo = clo[i], hi = chi[i];
      if (mutual) {
        const int cj = cluster[j];
        const float4 l2 = clo[j], h2 = chi[j];
        lo = make_float4(fminf(lo.x, l2.x), fminf(lo.y, l2.y), fminf(lo.z, l2.z), 0.f);
        hi = make_float4(fmaxf(hi.x, h2.x), fmaxf(hi.y, h2.y), fmaxf(hi.z, h2.z), 0.f);
        const int ci = id;
        id = (int)atomicAdd(ctr + 2, 1u);
        T.lo[id] = lo; T.hi[id] = hi; T.left[id] = ci; T.right[id] = cj; T.parent[id] = -1;
        T.count[id] = T.count[ci] + T.count[cj];
        T.parent[ci] = id; T.parent[cj] = id;
      }
      out_id[i] = id; olo[i] = lo; ohi[i] = hi;
    }
    keep[i] = stay ? 1 : 0;
  }
  const int c = __syncthreads_count(stay);
  if (threadIdx.x == 0 && blockIdx.x * blockDim.x < n) bcount[blockIdx.x] = (unsigned)c;
}

// exclusive scan of the per-block survivor counts (one block; a few thousand entries at most per pass)
__global__ void __launch_bounds__(1024) k_ploc_scan(unsigned* ctr, const unsigned* bcount, unsigned* boff) {
  __shared__ unsigned s_w[32];
  __shared__ unsigned s_carry;
  const int n = (int)ctr[0];
  const int nb = (n + BDPT_PLOC_BLOCK - 1) / BDPT_PLOC_BLOCK;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0u;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    const int b = base + threadIdx.x;
    const unsigned v = b < nb ? bcount[b] : 0u;
    unsigned x = v;
    for (int d = 1; d < 32; d <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
    if (lane == 31) s_w[wid] = x;
    __syncthreads();
    if (wid == 0) {
      unsigned w = s_w[lane];
      for (int d = 1; d < 32; d <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += y; }
      s_w[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    const unsigned woff = wid ? s_w[wid - 1] : 0u;
    if (b < nb) boff[b] = s_carry + woff + x - v;
    __syncthreads();
    if (threadIdx.x == 0) s_carry += s_w[31];
    __syncthreads();
  }
  // the round's output count becomes the next round's input; the scatter that follows reads this round's from ctr[8]
  if (threadIdx.x == 0) { const unsigned nn = n >= 2 ? s_carry : (unsigned)n; ctr[8] = (unsigned)n; ctr[1] = nn; ctr[0] = nn; }
}

__global__ void __launch_bounds__(BDPT_PLOC_BLOCK) k_ploc_scatter(const unsigned* ctr, const unsigned char* keep, const unsigned* boff, const int* out_id,
                                                                   const float4* olo, const float4* ohi, int* cluster, float4* clo, float4* chi) {
  __shared__ unsigned s_w[BDPT_PLOC_BLOCK / 32];
  const int n = (int)ctr[8];   // (this round's input count: k_ploc_scan has already moved ctr[0] on)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < 2) return;   // (whole grid: nothing to do, the array is final)
  const bool stay = i < n && keep[i];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned m = __ballot_sync(0xffffffffu, stay);
  if (lane == 0) s_w[wid] = (unsigned)__popc(m);
  __syncthreads();
  unsigned off = 0;
  for (int w = 0; w < wid; w++) off += s_w[w];
  if (stay) {
    const unsigned pos = boff[blockIdx.x] + off + (unsigned)__popc(m & ((1u << lane) - 1u));
    cluster[pos] = out_id[i]; clo[pos] = olo[i]; chi[pos] = ohi[i];
  }
}

// The last rounds in one launch.  PLOC's cluster count falls fast at first and slowly at the end (CBbunny4: 114 k -> 1 k in
// ~28 of its 56 rounds; a scene of 252 primitives alone takes 24): once at most 1,024 clusters are left, one block keeps the
// cluster array in shared memory and runs the remaining rounds — the same nearest / mutual-merge / order-preserving
// compaction steps as the kernels above, so the same tree — instead of five tiny launches per round.
// ctr[3] receives the number of rounds it ran.
__global__ void __launch_bounds__(1024) k_ploc_tail(unsigned* ctr, int* cluster, float4* clo, float4* chi, DevTree T, int radius) {
  __shared__ float4 s_lo[1024], s_hi[1024];
  __shared__ int s_id[1024], s_nn[1024];
  __shared__ unsigned s_w[32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int n = (int)ctr[0];
  if (tid < n) { s_id[tid] = cluster[tid]; s_lo[tid] = clo[tid]; s_hi[tid] = chi[tid]; }
  unsigned rounds = 0;
  __syncthreads();
  while (n >= 2) {
    if (tid < n) s_nn[tid] = ploc_nearest(tid, n, (const float*)s_lo, (const float*)s_hi, radius);
    __syncthreads();
    bool stay = false;
    int id = 0;
    float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
    if (tid < n) {
      const int j = s_nn[tid];
      const bool mutual = s_nn[j] == tid;
      stay = !(mutual && tid > j);
      if (stay) {
        id = s_id[tid]; lo = s_lo[tid]; hi = s_hi[tid];
        if (mutual) {
          const int cj = s_id[j];
          const float4 l2 = s_lo[j], h2 = s_hi[j];
          lo = make_float4(fminf(lo.x, l2.x), fminf(lo.y, l2.y), fminf(lo.z, l2.z), 0.f);
          hi = make_float4(fmaxf(hi.x, h2.x), fmaxf(hi.y, h2.y), fmaxf(hi.z, h2.z), 0.f);
          const int ci = id;
          id = (int)atomicAdd(ctr + 2, 1u);
          T.lo[id] = lo; T.hi[id] = hi; T.left[id] = ci; T.right[id] = cj; T.parent[id] = -1;
          T.count[id] = T.count[ci] + T.count[cj];
          T.parent[ci] = id; T.parent[cj] = id;
        }
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, stay);
    if (lane == 0) s_w[wid] = (unsigned)__popc(m);
    __syncthreads();   // (also: every read of the round's arrays is done)
    if (wid == 0) {
      unsigned w = s_w[lane];
      for (int d = 1; d < 32; d <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += y; }
      s_w[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    if (stay) {
      const unsigned pos = (wid ? s_w[wid - 1] : 0u) + (unsigned)__popc(m & ((1u << lane) - 1u));
      s_id[pos] = id; s_lo[pos] = lo; s_hi[pos] = hi;
    }
    n = (int)s_w[31];
    rounds++;
    __syncthreads();
  }
  if (tid < n) { cluster[tid] = s_id[tid]; clo[tid] = s_lo[tid]; chi[tid] = s_hi[tid]; }
  if (tid == 0) { ctr[0] = (unsigned)n; ctr[1] = (unsigned)n; ctr[3] = rounds; }
}

// ---- bottom-up collapse DP: one thread per leaf walks up; the second thread to arrive at a node computes it
__global__ void k_bvh_collapse_dp(DevTree T, int n_leaves, WideParams P, float* C, uint8_t* split, uint8_t* is_leaf, unsigned* visit) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_leaves) return;
  DevTreeView V{T};
  wide_dp_node(V, k, P, C, split, is_leaf);
  int node = k;
  for (;;) {
    const int p = T.parent[node];
    if (p < 0) break;
    __threadfence();
    if (atomicAdd(visit + p, 1u) == 0u) break;   // the sibling subtree is not done yet: its thread will do p
    __threadfence();
    // the children's rows were written by other threads, possibly on other SMs: read them past L1
    float cl[8], cr[8], lo[3], hi[3];
    const float4* pl = reinterpret_cast<const float4*>(C + 8 * (size_t)T.left[p]);
    const float4* pr = reinterpret_cast<const float4*>(C + 8 * (size_t)T.right[p]);
    const float4 l0 = __ldcg(pl), l1 = __ldcg(pl + 1), r0 = __ldcg(pr), r1 = __ldcg(pr + 1);
    cl[0] = l0.x; cl[1] = l0.y; cl[2] = l0.z; cl[3] = l0.w; cl[4] = l1.x; cl[5] = l1.y; cl[6] = l1.z; cl[7] = l1.w;
    cr[0] = r0.x; cr[1] = r0.y; cr[2] = r0.z; cr[3] = r0.w; cr[4] = r1.x; cr[5] = r1.y; cr[6] = r1.z; cr[7] = r1.w;
    V.box(p, lo, hi);
    wide_dp_combine(wide_half_area(lo, hi), T.count[p], cl, cr, P, C + 8 * (size_t)p, split + 9 * (size_t)p, is_leaf + p);
    node = p;
  }
}

// level 0 of the emission: wide node 0 stands for the root = the one cluster left
__global__ void k_bvh_root(const int* cluster, int2* items, unsigned* ectr) {
  items[0] = make_int2(0, cluster[0]);
  ectr[0] = 1u; ectr[1] = 0u; ectr[2] = 0u;
}

// ---- one level of wide nodes.  items: {wide node index, binary node}; ectr: [0] wide nodes so far  [1] leaf
// primitives placed so far  [2] items of the next level
__global__ void k_bvh_emit_level(DevTree T, const uint8_t* split, const uint8_t* is_leaf, const int2* items, int n_items,
                                 uint32_t* nodes, uint32_t* prim_order, int2* next_items, unsigned* ectr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_items) return;
  DevTreeView V{T};
  const int2 it = items[i];
  WidePlan plan;
  wide_plan_node(V, split, is_leaf, it.y, &plan);
  const uint32_t child_base = plan.n_internal ? atomicAdd(ectr + 0, (unsigned)plan.n_internal) : 0u;
  const uint32_t prim_base = plan.n_leaf_prims ? atomicAdd(ectr + 1, (unsigned)plan.n_leaf_prims) : 0u;
  int next_bn[8];
  uint32_t w[20];
  wide_write_node(V, is_leaf, it.y, plan, child_base, prim_base, w, prim_order, next_bn);
  uint4* dst = reinterpret_cast<uint4*>(nodes + 20 * (size_t)it.x);
  for (int q = 0; q < 5; q++) dst[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
  if (plan.n_internal) {
    const unsigned at = atomicAdd(ectr + 2, (unsigned)plan.n_internal);
    for (int k = 0; k < plan.n_internal; k++) next_items[at + k] = make_int2((int)child_base + k, next_bn[k]);
  }
}

// ---- leaf-ordered records the traversal and shading read (one definition for the device build and the host one)
//   prims   BDPT_PRIM_Q float4 per leaf slot, in the order {w row, u row, v row}.  Triangle: the three rows of the affine map world -> (u, v, w) with
//           p = v0 + u e1 + v e2 + w n, n = e1 x e2 (the "unit triangle" form of Woop et al.; computed in fp64): a ray test
//           is 5 three-term dot products and one reciprocal instead of Moller-Trumbore's two cross products and four.
//           A degenerate triangle gets zero rows (never hit, as in the reference: triangle.cpp:57-95 divides by 0 and
//           every comparison with the NaN fails).  Sphere: {0, 0, 0, NaN} (the NaN is the type tag), {centre, r^2}, -, -.
//   shading 3 x float4: the vertex normals; .w of the first two: primitive id in build_accel order, material
//           (| sign bit for a sphere)
//   prims64 10 doubles: the fp64 geometry finish_hit re-intersects the winner with
__host__ __device__ inline void make_leaf_record(const BdptPrim& p, uint32_t id, float4* hp, float4* hs, double* g) {
  const double* d = p.data;
  for (int a = 0; a < 9; a++) g[a] = d[a];
  g[9] = 0.0;
  union { int i; float f; } idb, matb;
  idb.i = (int)id;
  if (p.type == BDPT_PRIM_TRIANGLE) {
    matb.i = p.material;
    const double e1[3] = {d[3] - d[0], d[4] - d[1], d[5] - d[2]}, e2[3] = {d[6] - d[0], d[7] - d[1], d[8] - d[2]};
    const double n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    const double det = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
    double R[3][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
    if (det > 0.0 && det == det && 1.0 / det < 1e300) {
      const double inv = 1.0 / det;
      R[0][0] = (e2[1] * n[2] - e2[2] * n[1]) * inv; R[0][1] = (e2[2] * n[0] - e2[0] * n[2]) * inv; R[0][2] = (e2[0] * n[1] - e2[1] * n[0]) * inv;
      R[1][0] = (n[1] * e1[2] - n[2] * e1[1]) * inv; R[1][1] = (n[2] * e1[0] - n[0] * e1[2]) * inv; R[1][2] = (n[0] * e1[1] - n[1] * e1[0]) * inv;
      R[2][0] = n[0] * inv; R[2][1] = n[1] * inv; R[2][2] = n[2] * inv;
      for (int r = 0; r < 3; r++) R[r][3] = -(R[r][0] * d[0] + R[r][1] * d[1] + R[r][2] * d[2]);
    }
    for (int r = 0; r < 3; r++) hp[(r + 1) % 3] = make_float4((float)R[r][0], (float)R[r][1], (float)R[r][2], (float)R[r][3]);   // {w, u, v}
    if (BDPT_PRIM_Q > 3) hp[BDPT_PRIM_Q - 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    hs[0] = make_float4((float)d[9], (float)d[10], (float)d[11], idb.f);
    hs[1] = make_float4((float)d[12], (float)d[13], (float)d[14], matb.f);
    hs[2] = make_float4((float)d[15], (float)d[16], (float)d[17], 0.f);
  } else {
    matb.i = p.material | (int)0x80000000;
    union { unsigned u; float f; } nan; nan.u = 0x7fc00000u;
    hp[0] = make_float4(0.f, 0.f, 0.f, nan.f);
    hp[1] = make_float4((float)d[0], (float)d[1], (float)d[2], (float)(d[3] * d[3]));
    hp[2] = hp[BDPT_PRIM_Q - 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    hs[0] = make_float4(0.f, 0.f, 0.f, idb.f);
    hs[1] = make_float4(0.f, 0.f, 0.f, matb.f);
    hs[2] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__global__ void k_bvh_records(const BdptPrim* prims, const uint32_t* prim_order, int n, float4* hp, float4* hs, double* h64) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t id = prim_order[k];
  make_leaf_record(prims[id], id, hp + BDPT_PRIM_Q * (size_t)k, hs + 3 * (size_t)k, h64 + 10 * (size_t)k);
}

}  // namespace bdpt
