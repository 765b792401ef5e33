// bvh8.h — host-side build of the compressed 8-wide BVH the traversal kernels walk.
//
// The reference builds a binary pointer tree by spatial-midpoint splits with leaves of up to 4
// primitives and always descends into both children (reference src/scene/bvh.cpp:51-129,161-188).
// Closest-hit results do not depend on the tree (only exact ties do), so this path
// builds its own structure for the GPU: a binned-SAH binary tree collapsed into 8-wide nodes whose
// child boxes are quantised to 8 bits per plane relative to the node's own box (80-byte nodes, five
// 16-byte loads), children laid out so that a ray's direction octant gives a front-to-back visiting
// order without sorting, and primitives re-ordered so that each node's leaf primitives are contiguous.
//
// Node layout (80 B = 5 x uint4):
//   q0 = { px, py, pz (float bits: box origin),  ex | ey<<8 | ez<<16 | imask<<24 }
//        e* are biased fp32 exponents: child plane = p + q * 2^(e-127)
//        imask bit s = 1 if slot s holds an internal node
//   q1 = { child_base, prim_base, meta[0..3], meta[4..7] }
//        internal child in slot s = nodes[child_base + popcount(imask & ((1<<s)-1))]
//        meta[s]: 0 = empty; internal = 0x20 | (24 + s); leaf = (unary count 1/3/7)<<5 | first prim offset (0..23)
//        leaf prims of slot s = prims[prim_base + offset ... + count)
//   q2 = { qlo_x[0..3], qlo_x[4..7], qlo_y[0..3], qlo_y[4..7] }
//   q3 = { qlo_z[0..3], qlo_z[4..7], qhi_x[0..3], qhi_x[4..7] }
//   q4 = { qhi_y[0..3], qhi_y[4..7], qhi_z[0..3], qhi_z[4..7] }
#pragma once
#include <stdint.h>

#include <vector>

namespace bdpt {

struct BuildPrim {
  float lo[3], hi[3];  // conservative fp32 bounds
};

struct Bvh8 {
  std::vector<uint32_t> nodes;       // 20 words per node
  std::vector<uint32_t> prim_order;  // leaf-ordered -> original primitive index
  uint32_t max_depth = 0;
};

// Host build (binned SAH, top down).  Never fails for n >= 1; n == 0 yields a single empty node.  Since round 2 the
// product builds on the device (bvh_gpu.cuh); this one is the checker and the BDPT_BVH=host path.
void build_bvh8(const std::vector<BuildPrim>& prims, Bvh8* out);
// Host restatement of the device build (PLOC over Morton order + the same collapse): for the CPU tests.
void build_bvh8_ploc(const std::vector<BuildPrim>& prims, Bvh8* out);

}  // namespace bdpt
