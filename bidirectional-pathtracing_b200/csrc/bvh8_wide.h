// bvh8_wide.h — binary BVH -> compressed 8-wide BVH: the parts shared by the host builder (bvh8.cpp) and the
// device builder (bvh_gpu.cuh).  Plain functions over a "tree view", compiled for both sides.
//
//   wide_dp_node    one step of the SAH-optimal collapse (Ylitie, Karras, Laine 2017, section 3.2), bottom-up
//   wide_plan_node  the children of one wide node (the roots its binary subtree is cut into), their octant slots
//   wide_write_node quantisation grid, 8-bit child boxes, meta bytes: the 20 words of bvh8.h
//
// A tree view T offers:  box(n, lo[3], hi[3]);  left(n) / right(n) (-1 for a one-primitive node);  count(n);
// gather(n, out) — the primitive indices under n, for nodes of at most 3 primitives.
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define BDPT_HD __host__ __device__
#else
#define BDPT_HD
#endif

namespace bdpt {

struct WideParams {
  int max_leaf;      // primitives per leaf child (<= 3: the node's meta byte counts them in unary)
  float prim_cost;   // SAH cost of one primitive test relative to one visit of an 8-wide node
};

BDPT_HD inline float wide_half_area(const float* lo, const float* hi) {
  float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
  if (dx < 0) return 0.f;
  return dx * dy + dy * dz + dz * dx;
}

// C(n, i), i = 1..7: cheapest way to represent the subtree of binary node n as at most i roots.  One root is either a
// leaf child (<= max_leaf primitives) or a wide node whose <= 8 children are the roots its two subtrees are
// distributed into:  C(n,1) = min(A_n P_n c_p,  A_n + min_k C(l,k) + C(r,8-k)),
// C(n,i) = min(C(n,i-1), min_k C(l,k) + C(r,i-k)).
// c = C + 8 n (entries 1..7), sp = split + 9 n (entries 2..8: roots given to the left subtree; 0 = "use one root fewer").
// cl / cr: the children's rows (entries 1..7); c / sp / leaf: this node's outputs
BDPT_HD inline void wide_dp_combine(float A, int cnt, const float* cl, const float* cr, const WideParams& P, float* c, uint8_t* sp,
                                    uint8_t* leaf) {
  float dist[9];
  for (int j = 2; j <= 8; j++) {
    float best = FLT_MAX; int bk = 1;
    const int k0 = j - 7 > 1 ? j - 7 : 1, k1 = j - 1 < 7 ? j - 1 : 7;
    for (int k = k0; k <= k1; k++) {
      float v = cl[k] + cr[j - k];
      if (v < best) { best = v; bk = k; }
    }
    dist[j] = best; sp[j] = (uint8_t)bk;
  }
  const float c_leaf = cnt <= P.max_leaf ? A * (float)cnt * P.prim_cost : FLT_MAX;
  const float c_int = dist[8] + A;
  if (c_leaf <= c_int) { c[1] = c_leaf; *leaf = 1; } else { c[1] = c_int; *leaf = 0; }
  for (int i = 2; i <= 7; i++) {
    if (dist[i] < c[i - 1]) c[i] = dist[i];
    else { c[i] = c[i - 1]; sp[i] = 0; }
  }
}

template <class T>
BDPT_HD inline void wide_dp_node(const T& t, int n, const WideParams& P, float* C, uint8_t* split, uint8_t* is_leaf) {
  float lo[3], hi[3];
  t.box(n, lo, hi);
  const float A = wide_half_area(lo, hi);
  float* c = C + (size_t)n * 8;
  const int l = t.left(n);
  if (l < 0) {   // a single primitive
    for (int i = 1; i <= 7; i++) c[i] = A * P.prim_cost;
    is_leaf[n] = 1;
    return;
  }
  wide_dp_combine(A, t.count(n), C + (size_t)l * 8, C + (size_t)t.right(n) * 8, P, c, split + (size_t)n * 9, is_leaf + n);
}

// the roots that represent binary node n with a budget of i (iterative form of the recursion over `split`)
template <class T>
BDPT_HD inline void wide_collect(const T& t, const uint8_t* split, int n, int i, int* out, int& nout) {
  int sn[16], si[16], sp = 0;
  sn[0] = n; si[0] = i; sp = 1;
  while (sp > 0) {
    sp--;
    int cn = sn[sp], ci = si[sp];
    while (ci > 1 && t.left(cn) >= 0 && split[(size_t)cn * 9 + ci] == 0) ci--;
    if (ci == 1 || t.left(cn) < 0) { out[nout++] = cn; continue; }
    const int k = split[(size_t)cn * 9 + ci];
    // left first: push right, then left
    sn[sp] = t.right(cn); si[sp] = ci - k; sp++;
    sn[sp] = t.left(cn); si[sp] = k; sp++;
  }
}

struct WidePlan {
  int ch[8];           // binary nodes that become the children (collection order)
  int nch;
  int slot_child[8];   // slot -> index into ch, -1 = empty
  int n_internal;      // children that are wide nodes themselves
  int n_leaf_prims;    // primitives of the leaf children
};

// Children of the wide node that stands for binary node `bn`, and their slots: the assignment that maximises the
// total of (child centroid - node centroid) . (slot's octant direction), so that a ray's direction octant gives a
// front-to-back visiting order without sorting (Hungarian method on the 8 x 8 matrix, absent children = zero rows).
template <class T>
BDPT_HD inline void wide_plan_node(const T& t, const uint8_t* split, const uint8_t* is_leaf, int bn, WidePlan* P) {
  P->nch = 0;
  if (t.left(bn) < 0) P->ch[P->nch++] = bn;   // a one-primitive scene
  else {
    const int k = split[(size_t)bn * 9 + 8];
    wide_collect(t, split, t.left(bn), k, P->ch, P->nch);
    wide_collect(t, split, t.right(bn), 8 - k, P->ch, P->nch);
  }
  const int nch = P->nch;
  float tlo[3], thi[3], nc[3];
  t.box(bn, tlo, thi);
  for (int a = 0; a < 3; a++) nc[a] = 0.5f * (tlo[a] + thi[a]);
  float cost[8][8];
  P->n_internal = 0; P->n_leaf_prims = 0;
  for (int c = 0; c < nch; c++) {
    float lo[3], hi[3], off[3];
    t.box(P->ch[c], lo, hi);
    for (int a = 0; a < 3; a++) off[a] = 0.5f * (lo[a] + hi[a]) - nc[a];
    for (int s = 0; s < 8; s++)
      cost[c][s] = ((s & 1) ? off[0] : -off[0]) + ((s & 2) ? off[1] : -off[1]) + ((s & 4) ? off[2] : -off[2]);
    if (is_leaf[P->ch[c]]) P->n_leaf_prims += t.count(P->ch[c]); else P->n_internal++;
  }
  for (int s = 0; s < 8; s++) P->slot_child[s] = -1;
  const int N = 8;
  double a[N + 1][N + 1];
  for (int i = 1; i <= N; i++) for (int j = 1; j <= N; j++) a[i][j] = i <= nch ? -(double)cost[i - 1][j - 1] : 0.0;
  double u[N + 1], v[N + 1]; int pm[N + 1], way[N + 1];
  for (int j = 0; j <= N; j++) { u[j] = 0; v[j] = 0; pm[j] = 0; way[j] = 0; }
  for (int i = 1; i <= N; i++) {
    pm[0] = i; int j0 = 0; double minv[N + 1]; bool used[N + 1];
    for (int j = 0; j <= N; j++) { minv[j] = 1e300; used[j] = false; }
    do {
      used[j0] = true; int i0 = pm[j0], j1 = 0; double delta = 1e300;
      for (int j = 1; j <= N; j++) if (!used[j]) {
        double cur = a[i0][j] - u[i0] - v[j];
        if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
        if (minv[j] < delta) { delta = minv[j]; j1 = j; }
      }
      for (int j = 0; j <= N; j++) if (used[j]) { u[pm[j]] += delta; v[j] -= delta; } else minv[j] -= delta;
      j0 = j1;
    } while (pm[j0] != 0);
    do { int j1 = way[j0]; pm[j0] = pm[j1]; j0 = j1; } while (j0);
  }
  for (int j = 1; j <= N; j++) if (pm[j] >= 1 && pm[j] <= nch) P->slot_child[j - 1] = pm[j] - 1;
}

BDPT_HD inline uint32_t wide_fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
BDPT_HD inline uint32_t wide_pack4(const uint8_t* b) { return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24); }

// Writes the 20 words of the wide node (bvh8.h) and the leaf children's primitives into prim_order[prim_base ...).
// Internal children take the node indices child_base + (their rank in slot order); next_bn[rank] receives their
// binary nodes.  Returns the internal mask.
//
// Quantisation grid: every decoded plane keeps at least 1/64 of a quantum of slack to the box it bounds (the grid
// origin sits 1/32 quantum below the node's min corner): the traversal decodes a plane as
// fma(2^15 + q, a, b - 2^15 a), whose rounding error is below 1/512 quantum (traverse.cuh).
template <class T>
BDPT_HD inline uint32_t wide_write_node(const T& t, const uint8_t* is_leaf, int bn, const WidePlan& P, uint32_t child_base,
                                        uint32_t prim_base, uint32_t* w, uint32_t* prim_order, int* next_bn) {
  float tlo[3], thi[3];
  t.box(bn, tlo, thi);
  float p[3], scale[3], slack[3]; uint32_t ebyte[3];
  for (int a = 0; a < 3; a++) {
    const float lo = tlo[a], hi = thi[a];
    float ext = hi - lo;
    int e = -126;
    if (ext > 0.f) {
      e = (int)ceil(log2((double)ext / 254.9));
      e = e < -126 ? -126 : e > 126 ? 126 : e;
    }
    for (;; e++) {
      scale[a] = ldexpf(1.0f, e);
      slack[a] = scale[a] * (1.0f / 64.0f);
      p[a] = lo - 2.0f * slack[a];
      if (e >= 126 || p[a] + 255.0f * scale[a] >= hi + slack[a]) break;
    }
    ebyte[a] = (uint32_t)(e + 127);
  }
  uint8_t meta[8], q[6][8];
  for (int s = 0; s < 8; s++) { meta[s] = 0; for (int r = 0; r < 6; r++) q[r][s] = 0; }
  uint32_t imask = 0, prim_off = 0;
  int rank = 0;
  for (int s = 0; s < 8; s++) {
    const int c = P.slot_child[s];
    if (c < 0) continue;
    const int cn = P.ch[c];
    float clo[3], chi[3];
    t.box(cn, clo, chi);
    for (int a = 0; a < 3; a++) {
      double lo = ((double)clo[a] - (double)p[a]) / (double)scale[a];
      double hi = ((double)chi[a] - (double)p[a]) / (double)scale[a];
      int qlo = (int)floor(lo), qhi = (int)ceil(hi);
      qlo = qlo < 0 ? 0 : qlo > 255 ? 255 : qlo; qhi = qhi < 0 ? 0 : qhi > 255 ? 255 : qhi;
      // keep the decoded box conservative in fp32, with the slack the traversal's decode relies on
      while (qlo > 0 && p[a] + (float)qlo * scale[a] > clo[a] - slack[a]) qlo--;
      while (qhi < 255 && p[a] + (float)qhi * scale[a] < chi[a] + slack[a]) qhi++;
      q[a][s] = (uint8_t)qlo; q[3 + a][s] = (uint8_t)qhi;
    }
    if (!is_leaf[cn]) {
      imask |= 1u << s;
      meta[s] = (uint8_t)(0x20 | (24 + s));
      next_bn[rank++] = cn;
    } else {
      const int cnt = t.count(cn);
      const uint32_t unary = cnt == 1 ? 1u : cnt == 2 ? 3u : 7u;
      meta[s] = (uint8_t)((unary << 5) | prim_off);
      t.gather(cn, prim_order + prim_base + prim_off);
      prim_off += (uint32_t)cnt;
    }
  }
  w[0] = wide_fbits(p[0]); w[1] = wide_fbits(p[1]); w[2] = wide_fbits(p[2]);
  w[3] = ebyte[0] | (ebyte[1] << 8) | (ebyte[2] << 16) | (imask << 24);
  w[4] = child_base; w[5] = prim_base;
  w[6] = wide_pack4(meta); w[7] = wide_pack4(meta + 4);
  for (int r = 0; r < 6; r++) { w[8 + 2 * r] = wide_pack4(q[r]); w[9 + 2 * r] = wide_pack4(q[r] + 4); }
  return imask;
}

}  // namespace bdpt
