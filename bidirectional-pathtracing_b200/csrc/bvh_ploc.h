// bvh_ploc.h — per-element steps of the parallel locally-ordered clustering build (Meister & Bittner 2018, PLOC),
// shared by the device kernels (bvh_gpu.cuh) and the host restatement the CPU tests run (bvh8.cpp, build_bvh8_ploc).
//
// The reference builds its BVH on one host thread by recursive spatial-midpoint splits (src/scene/bvh.cpp:51-129);
// closest-hit results do not depend on the tree (SURVEY Q7 aside), so this path builds its own on the GPU:
//   1. 63-bit Morton codes of the primitive centroids, radix sort
//   2. PLOC: the sorted primitives are the initial clusters; every round each cluster looks at the RADIUS clusters on
//      either side of it in the array, picks the one whose union with it has the smallest surface area, and mutual
//      picks merge (the lower position keeps the new cluster); the array is compacted; until one cluster is left.
//      Bottom-up agglomeration by surface area: tree quality on a par with a binned-SAH top-down build.
//   3. the SAH-optimal collapse into 8-wide nodes and their quantisation (bvh8_wide.h), bottom-up then per level
#pragma once
#include "bvh8_wide.h"

namespace bdpt {

#define BDPT_PLOC_RADIUS 16

BDPT_HD inline uint64_t ploc_expand21(uint32_t v) {   // 21 bits -> every third bit of 63
  uint64_t x = v & 0x1fffffu;
  x = (x | (x << 32)) & 0x1f00000000ffffull;
  x = (x | (x << 16)) & 0x1f0000ff0000ffull;
  x = (x | (x << 8)) & 0x100f00f00f00f00full;
  x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}

// centroid (cx, cy, cz) inside the scene's centroid box [clo, chi]
BDPT_HD inline uint64_t ploc_morton(const float* c, const float* clo, const float* chi) {
  uint32_t q[3];
  for (int a = 0; a < 3; a++) {
    const float ext = chi[a] - clo[a];
#ifdef __CUDA_ARCH__   // (the library is built with -prec-div=false; the key must be the host restatement's bit for bit)
    float u = ext > 0.f ? __fdiv_rn(c[a] - clo[a], ext) : 0.f;
#else
    float u = ext > 0.f ? (c[a] - clo[a]) / ext : 0.f;
#endif
    u = u < 0.f ? 0.f : u > 1.f ? 1.f : u;
    double s = (double)u * 2097151.0;
    q[a] = (uint32_t)s;
  }
  return (ploc_expand21(q[0]) << 2) | (ploc_expand21(q[1]) << 1) | ploc_expand21(q[2]);
}

// Nearest neighbour of cluster position i among positions [i - R, i + R]: the one minimising the half area of the
// union box; ties by the pair's (lower, higher) position, so that the globally best pair always picks each other and
// every round merges at least one pair.  lo / hi: boxes of the clusters in ARRAY order, 4 floats apiece.
BDPT_HD inline int ploc_nearest(int i, int n, const float* lo4, const float* hi4, int radius = BDPT_PLOC_RADIUS) {
  const float* li = lo4 + 4 * (size_t)i; const float* hi_ = hi4 + 4 * (size_t)i;
  int best = -1; float best_a = FLT_MAX; int best_lo = 0, best_hi = 0;
  const int j0 = i - radius > 0 ? i - radius : 0;
  const int j1 = i + radius < n - 1 ? i + radius : n - 1;
  for (int j = j0; j <= j1; j++) {
    if (j == i) continue;
    const float* lj = lo4 + 4 * (size_t)j; const float* hj = hi4 + 4 * (size_t)j;
    float ulo[3], uhi[3];
    for (int a = 0; a < 3; a++) { ulo[a] = li[a] < lj[a] ? li[a] : lj[a]; uhi[a] = hi_[a] > hj[a] ? hi_[a] : hj[a]; }
    const float A = wide_half_area(ulo, uhi);
    const int pl = i < j ? i : j, ph = i < j ? j : i;
    if (A < best_a || (A == best_a && (pl < best_lo || (pl == best_lo && ph < best_hi)))) { best = j; best_a = A; best_lo = pl; best_hi = ph; }
  }
  return best;
}

}  // namespace bdpt
