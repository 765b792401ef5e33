// traverse.cuh — closest-hit / any-hit traversal of the compressed 8-wide BVH (bvh8.h).
//
// Replaces BVHAccel::intersect (reference src/scene/bvh.cpp:161-188, wrapper bvh.h:104-107),
// BBox::intersect (src/scene/bbox.cpp:10-56), Triangle::intersect (src/scene/triangle.cpp:57-95)
// and Sphere::intersect (src/scene/sphere.cpp:61-93).  Acceptance rules of the primitive tests are
// the reference's: min_t <= t <= max_t inclusive, b1 >= 0, b2 >= 0, b1 + b2 <= 1, no backface or
// determinant cull; a sphere reports its nearer root inside the interval, else the farther, and
// only if t > 0.  The box test only culls, so the answer is the reference's closest hit (exact
// ties between primitives excepted).
//
// Software traversal (B200 has no RT cores): a per-thread stack of "node group" entries
// {child base index, hit bits | internal mask}; one entry per visited node regardless of how many
// of its 8 children were hit, children popped nearest-first by count-leading-zeros on the hit bits,
// which the builder's octant slot assignment makes a front-to-back order.
#pragma once
#include "common.cuh"

namespace bdpt {

#define BDPT_STACK 32

struct Hit {
  float t;
  int leaf;  // index into the leaf-ordered primitive arrays, -1 = miss
};   // (no barycentrics: finish_hit computes them for the one winner — two registers less in the traversal loops)

// byte k of w as the float 2^15 + bb, without an int->float conversion: PRMT builds 0x4700bb00.
// The box test folds the 2^15 into its per-axis constant, so a quantised plane costs one PRMT and
// one FFMA; the fold's rounding error (< 1/512 quantum) is inside the slack the builder leaves
// (bvh8.cpp, "quantisation grid").
#define BDPT_QBIAS 32768.0f
__device__ inline float qbyte(uint32_t w, int k, uint32_t bias_bits) {   // bias_bits = DScene::qbias_bits
  return __uint_as_float(__byte_perm(w, bias_bits, 0x7504u + ((uint32_t)k << 4)));
}

// Intersect one leaf primitive; returns true and shrinks tmax on an accepted hit.  (Record layouts: bvh_gpu.cuh,
// make_leaf_record.)
template <int TR = 0>
__device__ inline bool intersect_prim(const float4* __restrict__ prims, int k, f3 o, f3 d, float tmin, float& tmax) {
  float4 c, a;
#if BDPT_PRIM_Q == 4
  ld256nc(prims + BDPT_PRIM_Q * k, &c, &a);
#else
  c = ldg4(prims + BDPT_PRIM_Q * k); a = ldg4(prims + BDPT_PRIM_Q * k + 1);
#endif
  if ((TR & BDPT_TRAIT_TRIS) || c.w == c.w) {
    // triangle: rows of world -> (u, v, w); the ray meets w = 0 at t, and (u, v) of that point are the barycentrics
    const float4 b = ldg4(prims + BDPT_PRIM_Q * k + 2);
    const float ow = fmaf(c.x, o.x, fmaf(c.y, o.y, fmaf(c.z, o.z, c.w)));
    const float dw = fmaf(c.x, d.x, fmaf(c.y, d.y, c.z * d.z));
    const float t = -ow * rcp_fast(dw);            // the winner is re-intersected in fp64 (finish_hit)
    // (hit point first, then its two coordinates: the form O + t D per coordinate, and an early exit on t, were both slower)
    const float px = fmaf(t, d.x, o.x), py = fmaf(t, d.y, o.y), pz = fmaf(t, d.z, o.z);
    const float b1 = fmaf(a.x, px, fmaf(a.y, py, fmaf(a.z, pz, a.w)));
    const float b2 = fmaf(b.x, px, fmaf(b.y, py, fmaf(b.z, pz, b.w)));
    if (t >= tmin && t <= tmax && b1 >= 0.f && b2 >= 0.f && b1 + b2 <= 1.f) {
      tmax = t;
      return true;
    }
    return false;
  }
  // sphere: a = {centre, r^2}
  f3 oc = o - mk3(a.x, a.y, a.z);
  float A = dot(d, d), hb = dot(oc, d);
  f3 l = oc - d * (hb / A);
  float disc = A * (a.w - dot(l, l));
  if (disc < 0.f) return false;
  float sd = sqrtf(disc), ia = 1.0f / A;
  float t1 = (-hb - sd) * ia, t2 = (-hb + sd) * ia, t = -1.f;
  if (t1 >= tmin && t1 <= tmax) t = t1;
  else if (t2 >= tmin && t2 <= tmax) t = t2;
  if (t > 0.f) { tmax = t; return true; }
  return false;
}

// Resumable traversal state: begin() then step() until it returns true.  One step = pop the nearest
// pending internal child, test its 8 children, test the primitives of the leaf children that were
// hit.  Being resumable is what lets the persistent kernels refill a lane the moment its ray ends
// instead of idling until the slowest ray of the warp is done.
struct Trav {
  f3 o, d, idir;
  float tmin;
  uint32_t oct;
  uint2 group;
  uint2 tg;   // closest-hit only: a copy of stack[sp - 1], loaded one pop early (below)
  int sp;
  Hit hit;

  __device__ inline void begin(f3 o_, f3 d_, float tmin_, float tmax_) {
    o = o_; d = d_; tmin = tmin_;
    hit.t = tmax_; hit.leaf = -1;
    oct = (d.x < 0.f ? 1u : 0u) | (d.y < 0.f ? 2u : 0u) | (d.z < 0.f ? 4u : 0u);
    // MUFU.RCP (<= 2 ulp): the reciprocals only feed the box cull, whose slack factor below covers them
    idir.x = rcp_fast(fabsf(d.x) > 1e-20f ? d.x : copysignf(1e-20f, d.x));
    idir.y = rcp_fast(fabsf(d.y) > 1e-20f ? d.y : copysignf(1e-20f, d.y));
    idir.z = rcp_fast(fabsf(d.z) > 1e-20f ? d.z : copysignf(1e-20f, d.z));
    sp = 0;
    group = make_uint2(0u, 0x80000000u);  // root: node 0, top priority bit
  }

  // `stack` is the caller's per-thread array of BDPT_STACK entries (kept out of this struct so that the
  // scalar state stays in registers)
  // `top` (optional): the first `ntop` nodes of the tree — its upper levels, the builder emits nodes breadth
  // first — staged in shared memory by the calling kernel.
  // One node visit: pop the nearest pending internal child of the current group, test its 8 children; the hit internal
  // children become the current group; the hit leaf primitives are returned as a bit set (bit i = primitive *pbase + i).
  template <bool ANY_HIT>
  __device__ inline uint32_t node_visit(const DScene& sc, uint2* stack, const uint4* top, const uint32_t ntop, int* pbase) {
    const uint32_t octinv = oct ^ 7u;
    // ---- pop the nearest internal child of the current group
    uint32_t bit = 31u - (uint32_t)__clz(group.y);
    group.y &= ~(1u << bit);
    uint32_t base = group.x;
    uint32_t imask_prev = group.y & 0xffu;
    if (group.y > 0x00ffffffu) { stack[sp++] = group; if (!ANY_HIT) tg = group; }
    uint32_t slot = (bit - 24u) ^ octinv;
    uint32_t rel = __popc(imask_prev & ((1u << slot) - 1u));
    const uint32_t ni = base + rel;
    uint4 q0, q1, q2, q3, q4;
    if (ni < ntop) {
      const uint4* np = top + (size_t)ni * 5;
      q0 = np[0]; q1 = np[1]; q2 = np[2]; q3 = np[3]; q4 = np[4];
    } else {
      const uint4* np = sc.nodes + (size_t)ni * BDPT_NODE_Q;
#if BDPT_NODE_Q == 6
      ld256nc(np, &q0, &q1); ld256nc(np + 2, &q2, &q3); q4 = ldg4(np + 4);
#else
      q0 = ldg4(np); q1 = ldg4(np + 1); q2 = ldg4(np + 2); q3 = ldg4(np + 3); q4 = ldg4(np + 4);
#endif
    }

    f3 adj, org;   // plane q of an axis is hit at t = (2^15 + q) * adj + org
    adj.x = __uint_as_float((q0.w & 0xffu) << 23) * idir.x;
    adj.y = __uint_as_float(((q0.w >> 8) & 0xffu) << 23) * idir.y;
    adj.z = __uint_as_float(((q0.w >> 16) & 0xffu) << 23) * idir.z;
    org.x = fmaf(-BDPT_QBIAS, adj.x, (__uint_as_float(q0.x) - o.x) * idir.x);
    org.y = fmaf(-BDPT_QBIAS, adj.y, (__uint_as_float(q0.y) - o.y) * idir.y);
    org.z = fmaf(-BDPT_QBIAS, adj.z, (__uint_as_float(q0.z) - o.z) * idir.z);
    const uint32_t imask = q0.w >> 24;

    // near / far plane words per axis, chosen by the direction sign
    uint32_t nx0 = (oct & 1u) ? q3.z : q2.x, nx1 = (oct & 1u) ? q3.w : q2.y;
    uint32_t fx0 = (oct & 1u) ? q2.x : q3.z, fx1 = (oct & 1u) ? q2.y : q3.w;
    uint32_t ny0 = (oct & 2u) ? q4.x : q2.z, ny1 = (oct & 2u) ? q4.y : q2.w;
    uint32_t fy0 = (oct & 2u) ? q2.z : q4.x, fy1 = (oct & 2u) ? q2.w : q4.y;
    uint32_t nz0 = (oct & 4u) ? q4.z : q3.x, nz1 = (oct & 4u) ? q4.w : q3.y;
    uint32_t fz0 = (oct & 4u) ? q3.x : q4.z, fz1 = (oct & 4u) ? q3.y : q4.w;

    // per child byte of the meta words, four children per 32-bit operation: the bit position a hit child
    // sets (internal children: 24 + slot, mirrored by the ray octant so that the highest bit is the
    // nearest child; leaves: their primitive offset) and the bits it sets there (1 / 11 / 111)
    const uint32_t octinv4 = octinv * 0x01010101u;
    const uint32_t in0 = (((q1.z & (q1.z << 1)) & 0x10101010u) >> 4) * 0xffu;
    const uint32_t in1 = (((q1.w & (q1.w << 1)) & 0x10101010u) >> 4) * 0xffu;
    const uint32_t pos0 = q1.z ^ (octinv4 & in0), pos1 = q1.w ^ (octinv4 & in1);
    const uint32_t cnt0 = (q1.z >> 5) & 0x07070707u, cnt1 = (q1.w >> 5) & 0x07070707u;

    const uint32_t qb = sc.qbias_bits;
    uint32_t hitmask = 0u;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int k = i & 3;
      uint32_t nxw = i < 4 ? nx0 : nx1, fxw = i < 4 ? fx0 : fx1;
      uint32_t nyw = i < 4 ? ny0 : ny1, fyw = i < 4 ? fy0 : fy1;
      uint32_t nzw = i < 4 ? nz0 : nz1, fzw = i < 4 ? fz0 : fz1;
      float t0 = fmaxf(fmaxf(fmaf(qbyte(nxw, k, qb), adj.x, org.x), fmaf(qbyte(nyw, k, qb), adj.y, org.y)),
                       fmaxf(fmaf(qbyte(nzw, k, qb), adj.z, org.z), tmin));
      float t1 = fminf(fminf(fmaf(qbyte(fxw, k, qb), adj.x, org.x), fmaf(qbyte(fyw, k, qb), adj.y, org.y)),
                       fminf(fmaf(qbyte(fzw, k, qb), adj.z, org.z), hit.t));
      t1 *= 1.000001f;  // keep the cull conservative under fp32 rounding (8 ulp: two 2-ulp reciprocals + the FMAs)
      const uint32_t pos = (i < 4 ? pos0 : pos1) >> (8 * k);             // the shift below uses its low 5 bits
      const uint32_t cnt = ((i < 4 ? cnt0 : cnt1) >> (8 * k)) & 0xffu;
      const uint32_t bits = cnt << (pos & 31u);
      hitmask |= t0 <= t1 ? bits : 0u;
    }
    group = make_uint2(q1.x, (hitmask & 0xff000000u) | imask);

    *pbase = (int)q1.y;
    return hitmask & 0x00ffffffu;
  }

  // Primitive tests of a leaf set; true = an any-hit ray found its hit.
  template <bool ANY_HIT, int TR>
  __device__ inline bool leaf_tests(const DScene& sc, uint32_t tris, int pbase) {
    while (tris) {
      int i = __ffs(tris) - 1;
      tris &= tris - 1u;
      int k = pbase + i;
      if (intersect_prim<TR>(sc.prims, k, o, d, tmin, hit.t)) {
        hit.leaf = k;
        if (ANY_HIT) return true;
      }
    }
    return false;
  }

  // The current group has no pending child left: take the next one from the stack.  True = the traversal is over.
  template <bool ANY_HIT>
  __device__ inline bool next_group(uint2* stack) {
    if (group.y <= 0x00ffffffu) {
      if (sp == 0) return true;
      if (!ANY_HIT) {
        // the popped entry comes from a register; the one below it is requested now and not needed before the next pop
        // or push — the local-memory latency hides behind the node visits in between (walk -1.8 %; the any-hit kernel
        // has no registers to spare for it: +5 %)
        group = tg; --sp;
        if (sp > 0) tg = stack[sp - 1];
      } else {
        group = stack[--sp];
      }
    }
    return false;
  }

  template <bool ANY_HIT, int TR = 0>
  __device__ inline bool step(const DScene& sc, uint2* stack, const uint4* top = nullptr, const uint32_t ntop = 0) {
    int pbase;
    const uint32_t tris = node_visit<ANY_HIT>(sc, stack, top, ntop, &pbase);
    if (leaf_tests<ANY_HIT, TR>(sc, tris, pbase)) return true;
    return next_group<ANY_HIT>(stack);
  }

};

template <bool ANY_HIT, int TR = 0>
__device__ inline Hit traverse(const DScene& sc, f3 o, f3 d, float tmin, float tmax) {
  Trav T;
  uint2 stack[BDPT_STACK];
  T.begin(o, d, tmin, tmax);
  while (!T.template step<ANY_HIT, TR>(sc, stack)) {}
  return T.hit;
}

// Shading data of an accepted closest hit: interpolated (unflipped) unit normal, primitive id in
// the reference's order and material id (Triangle::intersect triangle.cpp:80-87, Sphere :78-87).
//
// The winner of the fp32 traversal is re-intersected ONCE in fp64 against the fp64 geometry the
// loader delivered (B200 has full-rate-class FP64 units; this is ~60 DFMA per ray, not per test):
// t and the barycentrics then agree with the fp64 reference to ~1e-12 relative instead of the
// ~1e-4 an fp32 Moller-Trumbore gives at grazing incidence, which keeps hit points on their
// surface and the EPS_F = 1e-5 self-intersection offsets meaningful.
template <int TR = 0>
__device__ inline void finish_hit(const DScene& sc, Hit& h, f3 o, f3 d, f3* n, int* prim_id, int* mat, f3* hit_p) {
  const float4 n1 = ldg4(sc.shading + 3 * h.leaf), n2 = ldg4(sc.shading + 3 * h.leaf + 1);
  *prim_id = __float_as_int(n1.w);
  int mw = __float_as_int(n2.w);
  *mat = mw & 0x7fffffff;
  const double* g = sc.prims64 + 10 * (size_t)h.leaf;
  const double ox = o.x, oy = o.y, oz = o.z, dx = d.x, dy = d.y, dz = d.z;
  if ((TR & BDPT_TRAIT_TRIS) || mw >= 0) {
    const double2 g0 = __ldg((const double2*)g), g1 = __ldg((const double2*)g + 1), g2 = __ldg((const double2*)g + 2),
                  g3 = __ldg((const double2*)g + 3), g4 = __ldg((const double2*)g + 4);
    const double p1x = g0.x, p1y = g0.y, p1z = g1.x;
    const double e1x = g1.y - p1x, e1y = g2.x - p1y, e1z = g2.y - p1z;
    const double e2x = g3.x - p1x, e2y = g3.y - p1y, e2z = g4.x - p1z;
    const double sx = ox - p1x, sy = oy - p1y, sz = oz - p1z;
    const double s1x = dy * e2z - dz * e2y, s1y = dz * e2x - dx * e2z, s1z = dx * e2y - dy * e2x;
    const double s2x = sy * e1z - sz * e1y, s2y = sz * e1x - sx * e1z, s2z = sx * e1y - sy * e1x;
    const double inv = 1.0 / (s1x * e1x + s1y * e1y + s1z * e1z);
    const double t = (s2x * e2x + s2y * e2y + s2z * e2z) * inv;
    const double b1 = (s1x * sx + s1y * sy + s1z * sz) * inv;
    const double b2 = (s2x * dx + s2y * dy + s2z * dz) * inv;
    float hb1, hb2;
    if (t == t && fabs(t - (double)h.t) <= 1e-3 * fabs((double)h.t)) {  // sanity: same root
      h.t = (float)t; hb1 = (float)b1; hb2 = (float)b2;
      *hit_p = mk3((float)(ox + dx * t), (float)(oy + dy * t), (float)(oz + dz * t));
    } else {
      // (never seen on the fixture scenes) keep the fp32 hit: its barycentrics from the record's u and v rows
      *hit_p = o + d * h.t;
      const float4 ru = ldg4(sc.prims + BDPT_PRIM_Q * h.leaf + 1), rv = ldg4(sc.prims + BDPT_PRIM_Q * h.leaf + 2);
      hb1 = fmaf(ru.x, hit_p->x, fmaf(ru.y, hit_p->y, fmaf(ru.z, hit_p->z, ru.w)));
      hb2 = fmaf(rv.x, hit_p->x, fmaf(rv.y, hit_p->y, fmaf(rv.z, hit_p->z, rv.w)));
    }
    const float4 n3 = ldg4(sc.shading + 3 * h.leaf + 2);
    float b0 = 1.f - hb1 - hb2;
    f3 nn = mk3(n1.x, n1.y, n1.z) * b0 + hb1 * mk3(n2.x, n2.y, n2.z) + hb2 * mk3(n3.x, n3.y, n3.z);
    *n = normalized(nn);
  } else {
    const double2 g0 = __ldg((const double2*)g), g1 = __ldg((const double2*)g + 1);
    const double cx = g0.x, cy = g0.y, cz = g1.x, r = g1.y;
    const double ocx = ox - cx, ocy = oy - cy, ocz = oz - cz;
    const double A = dx * dx + dy * dy + dz * dz, hb = ocx * dx + ocy * dy + ocz * dz;
    const double k = hb / A;
    const double lx = ocx - dx * k, ly = ocy - dy * k, lz = ocz - dz * k;
    const double disc = A * (r * r - (lx * lx + ly * ly + lz * lz));
    double t = (double)h.t;
    if (disc >= 0.0) {
      const double sd = sqrt(disc);
      const double t1 = (-hb - sd) / A, t2 = (-hb + sd) / A;
      t = fabs(t1 - t) <= fabs(t2 - t) ? t1 : t2;  // the root the fp32 test picked
    }
    h.t = (float)t;
    const double px = ox + dx * t, py = oy + dy * t, pz = oz + dz * t;
    *hit_p = mk3((float)px, (float)py, (float)pz);
    const double nx = px - cx, ny = py - cy, nz = pz - cz;
    const double il = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);
    *n = mk3((float)(nx * il), (float)(ny * il), (float)(nz * il));
  }
}

}  // namespace bdpt
