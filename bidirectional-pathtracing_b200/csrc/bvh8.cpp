// bvh8.cpp — binned-SAH binary build, collapse to 8-wide, octant slot assignment, 8-bit
// quantisation.  Host code; see bvh8.h for the node layout and for how this relates to the
// reference's BVHAccel (reference src/scene/bvh.cpp:51-129).
#include "bvh8.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <deque>

namespace bdpt {
namespace {

struct Box {
  float lo[3], hi[3];
  void reset() { for (int a = 0; a < 3; a++) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; } }
  void grow(const float* l, const float* h) {
    for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], l[a]); hi[a] = std::max(hi[a], h[a]); }
  }
  void grow(const Box& b) { grow(b.lo, b.hi); }
  float half_area() const {
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    if (dx < 0) return 0.f;
    return dx * dy + dy * dz + dz * dx;
  }
};

struct BNode {
  Box box;
  int left = -1, right = -1;  // -1 => leaf
  int first = 0, count = 0;
};

struct Builder {
  const std::vector<BuildPrim>& prims;
  std::vector<uint32_t> idx;
  std::vector<float> cent;  // 3 per prim
  std::vector<BNode> bn;
  static constexpr int kBins = 16;
  int kMaxLeaf = 3;        // primitives per leaf child (<= 3: the node's meta byte counts them in unary)
  float kTravCost = 0.5f;  // SAH cost of one more traversal step relative to one primitive test

  explicit Builder(const std::vector<BuildPrim>& p) : prims(p) {
    idx.resize(p.size());
    cent.resize(p.size() * 3);
    for (size_t i = 0; i < p.size(); i++) {
      idx[i] = (uint32_t)i;
      for (int a = 0; a < 3; a++) cent[3 * i + a] = 0.5f * (p[i].lo[a] + p[i].hi[a]);
    }
    bn.reserve(p.size() * 2 + 1);
    if (const char* e = getenv("BDPT_BVH_LEAF")) kMaxLeaf = std::max(1, std::min(3, atoi(e)));   // tuning knobs
    if (const char* e = getenv("BDPT_BVH_CT")) kTravCost = (float)atof(e);
  }

  int build(int first, int count) {
    BNode n;
    n.box.reset();
    Box cb; cb.reset();
    for (int i = first; i < first + count; i++) {
      const BuildPrim& p = prims[idx[i]];
      n.box.grow(p.lo, p.hi);
      const float* c = &cent[3 * idx[i]];
      cb.grow(c, c);
    }
    n.first = first; n.count = count;
    int me = (int)bn.size();
    bn.push_back(n);
    if (count == 1) return me;

    // binned SAH over the three axes
    float best_cost = FLT_MAX; int best_axis = -1, best_split = -1;
    for (int a = 0; a < 3; a++) {
      float ext = cb.hi[a] - cb.lo[a];
      if (!(ext > 0.f)) continue;
      Box bb[kBins]; int bc[kBins];
      for (int b = 0; b < kBins; b++) { bb[b].reset(); bc[b] = 0; }
      float k = kBins * (1.f - 1e-6f) / ext;
      for (int i = first; i < first + count; i++) {
        int b = std::min(kBins - 1, std::max(0, (int)((cent[3 * idx[i] + a] - cb.lo[a]) * k)));
        bb[b].grow(prims[idx[i]].lo, prims[idx[i]].hi); bc[b]++;
      }
      float la[kBins]; int lc[kBins];
      Box acc; acc.reset(); int cnt = 0;
      for (int b = 0; b < kBins; b++) { acc.grow(bb[b]); cnt += bc[b]; la[b] = acc.half_area(); lc[b] = cnt; }
      acc.reset(); cnt = 0;
      for (int b = kBins - 1; b >= 1; b--) {
        acc.grow(bb[b]); cnt += bc[b];
        if (lc[b - 1] == 0 || cnt == 0) continue;
        float cost = la[b - 1] * lc[b - 1] + acc.half_area() * cnt;
        if (cost < best_cost) { best_cost = cost; best_axis = a; best_split = b; }
      }
    }
    float area = n.box.half_area();
    if (count <= kMaxLeaf) {
      float leaf_cost = area * count;
      float split_cost = best_axis < 0 ? FLT_MAX : kTravCost * area + best_cost;
      if (leaf_cost <= split_cost) return me;
    }
    int mid;
    if (best_axis < 0) {
      mid = first + count / 2;  // identical centroids: split by index
    } else {
      float ext = cb.hi[best_axis] - cb.lo[best_axis];
      float k = kBins * (1.f - 1e-6f) / ext;
      float lo = cb.lo[best_axis];
      auto it = std::partition(idx.begin() + first, idx.begin() + first + count, [&](uint32_t id) {
        int b = std::min(kBins - 1, std::max(0, (int)((cent[3 * id + best_axis] - lo) * k)));
        return b < best_split;
      });
      mid = (int)(it - idx.begin());
      if (mid == first || mid == first + count) mid = first + count / 2;
    }
    int l = build(first, mid - first);
    int r = build(mid, first + count - mid);
    bn[me].left = l; bn[me].right = r;
    return me;
  }
};

inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

}  // namespace

void build_bvh8(const std::vector<BuildPrim>& prims, Bvh8* out) {
  out->nodes.clear(); out->prim_order.clear(); out->max_depth = 1;
  if (prims.empty()) { out->nodes.assign(20, 0u); return; }
  Builder B(prims);
  int root = B.build(0, (int)prims.size());

  struct Item { int wide; int bnode; uint32_t depth; };
  std::deque<Item> queue;
  out->nodes.assign(20, 0u);
  queue.push_back({0, root, 1});
  while (!queue.empty()) {
    Item it = queue.front(); queue.pop_front();
    out->max_depth = std::max(out->max_depth, it.depth);
    const BNode& top = B.bn[it.bnode];
    // collapse: expand the largest internal child until 8 children
    int ch[8]; int nch = 0;
    if (top.left < 0) { ch[nch++] = it.bnode; }
    else { ch[nch++] = top.left; ch[nch++] = top.right; }
    while (nch < 8) {
      int pick = -1; float best = -1.f;
      for (int i = 0; i < nch; i++) {
        const BNode& c = B.bn[ch[i]];
        if (c.left < 0) continue;
        float a = c.box.half_area();
        if (a > best) { best = a; pick = i; }
      }
      if (pick < 0) break;
      const BNode& c = B.bn[ch[pick]];
      int l = c.left, r = c.right;
      ch[pick] = l; ch[nch++] = r;
    }
    // octant slot assignment (greedy maximum of centroid offset . slot direction)
    float nc[3];
    for (int a = 0; a < 3; a++) nc[a] = 0.5f * (top.box.lo[a] + top.box.hi[a]);
    float cost[8][8];
    for (int c = 0; c < nch; c++) {
      const Box& b = B.bn[ch[c]].box;
      float off[3];
      for (int a = 0; a < 3; a++) off[a] = 0.5f * (b.lo[a] + b.hi[a]) - nc[a];
      for (int s = 0; s < 8; s++)
        cost[c][s] = ((s & 1) ? off[0] : -off[0]) + ((s & 2) ? off[1] : -off[1]) + ((s & 4) ? off[2] : -off[2]);
    }
    int slot_child[8]; for (int s = 0; s < 8; s++) slot_child[s] = -1;
    bool cdone[8] = {false, false, false, false, false, false, false, false};
    for (int k = 0; k < nch; k++) {
      float best = -FLT_MAX; int bc = -1, bs = -1;
      for (int c = 0; c < nch; c++) {
        if (cdone[c]) continue;
        for (int s = 0; s < 8; s++) {
          if (slot_child[s] >= 0) continue;
          if (cost[c][s] > best) { best = cost[c][s]; bc = c; bs = s; }
        }
      }
      cdone[bc] = true; slot_child[bs] = bc;
    }
    // quantisation grid.  Every decoded plane keeps at least 1/64 of a quantum of slack to the box it
    // bounds (the grid origin sits 1/32 quantum below the node's min corner): the traversal decodes a
    // plane as fma(2^15 + q, a, b - 2^15 a), whose rounding error is below 1/512 quantum (traverse.cuh).
    float p[3]; uint32_t ebyte[3]; float scale[3], slack[3];
    for (int a = 0; a < 3; a++) {
      const float lo = top.box.lo[a], hi = top.box.hi[a];
      float ext = hi - lo;
      int e = -126;
      if (ext > 0.f) {
        e = (int)std::ceil(std::log2((double)ext / 254.9));
        e = std::max(-126, std::min(126, e));
      }
      for (;; e++) {
        scale[a] = std::ldexp(1.0f, e);
        slack[a] = scale[a] * (1.0f / 64.0f);
        p[a] = lo - 2.0f * slack[a];
        if (e >= 126 || p[a] + 255.0f * scale[a] >= hi + slack[a]) break;
      }
      ebyte[a] = (uint32_t)(e + 127);
    }
    uint8_t meta[8] = {0}, q[6][8];
    memset(q, 0, sizeof q);
    uint32_t imask = 0;
    uint32_t child_base = (uint32_t)(out->nodes.size() / 20);
    uint32_t prim_base = (uint32_t)out->prim_order.size();
    uint32_t prim_off = 0;
    for (int s = 0; s < 8; s++) {
      int c = slot_child[s];
      if (c < 0) continue;
      const BNode& cn = B.bn[ch[c]];
      for (int a = 0; a < 3; a++) {
        double lo = ((double)cn.box.lo[a] - (double)p[a]) / (double)scale[a];
        double hi = ((double)cn.box.hi[a] - (double)p[a]) / (double)scale[a];
        int qlo = (int)std::floor(lo), qhi = (int)std::ceil(hi);
        qlo = std::max(0, std::min(255, qlo)); qhi = std::max(0, std::min(255, qhi));
        // keep the decoded box conservative in fp32, with the slack the traversal's decode relies on
        while (qlo > 0 && p[a] + (float)qlo * scale[a] > cn.box.lo[a] - slack[a]) qlo--;
        while (qhi < 255 && p[a] + (float)qhi * scale[a] < cn.box.hi[a] + slack[a]) qhi++;
        q[a][s] = (uint8_t)qlo; q[3 + a][s] = (uint8_t)qhi;
      }
      if (cn.left >= 0) {
        imask |= 1u << s;
        meta[s] = (uint8_t)(0x20 | (24 + s));
      } else {
        uint32_t unary = cn.count == 1 ? 1u : cn.count == 2 ? 3u : 7u;
        meta[s] = (uint8_t)((unary << 5) | prim_off);
        for (int i = 0; i < cn.count; i++) out->prim_order.push_back(B.idx[cn.first + i]);
        prim_off += (uint32_t)cn.count;
      }
    }
    // children in slot order get consecutive node indices
    for (int s = 0; s < 8; s++) {
      if (!(imask & (1u << s))) continue;
      int wide = (int)(out->nodes.size() / 20);
      out->nodes.resize(out->nodes.size() + 20, 0u);
      queue.push_back({wide, ch[slot_child[s]], it.depth + 1});
    }
    uint32_t* w = &out->nodes[(size_t)it.wide * 20];
    w[0] = fbits(p[0]); w[1] = fbits(p[1]); w[2] = fbits(p[2]);
    w[3] = ebyte[0] | (ebyte[1] << 8) | (ebyte[2] << 16) | (imask << 24);
    w[4] = child_base; w[5] = prim_base;
    auto pack4 = [](const uint8_t* b) { return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24); };
    w[6] = pack4(meta); w[7] = pack4(meta + 4);
    for (int r = 0; r < 6; r++) { w[8 + 2 * r] = pack4(q[r]); w[9 + 2 * r] = pack4(q[r] + 4); }
  }
}

}  // namespace bdpt
