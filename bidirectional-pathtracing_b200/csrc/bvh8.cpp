// bvh8.cpp — binned-SAH binary build, collapse to 8-wide, octant slot assignment, 8-bit
// quantisation.  Host code; see bvh8.h for the node layout and for how this relates to the
// reference's BVHAccel (reference src/scene/bvh.cpp:51-129).
#include "bvh8.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <deque>
#include <thread>

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
  std::vector<BNode> bn;    // preallocated (2n - 1 nodes at most): subtrees are built by several threads
  std::atomic<int> next_node{0};
  std::atomic<int> spare_threads{0};
  static constexpr int kMaxBins = 64;
  int kBins = 16;          // SAH bins per axis (BDPT_BVH_BINS, tuning only)
  static constexpr int kParallelMin = 2048;   // subtrees smaller than this are not worth a thread
  int kMaxLeaf = 3;        // primitives per leaf child (<= 3: the node's meta byte counts them in unary)
  float kPrimCost = 0.8f;  // SAH cost of one primitive test relative to one visit of an 8-wide node (0.4 … 1.0: within 1.5 %)

  explicit Builder(const std::vector<BuildPrim>& p) : prims(p) {
    idx.resize(p.size());
    cent.resize(p.size() * 3);
    for (size_t i = 0; i < p.size(); i++) {
      idx[i] = (uint32_t)i;
      for (int a = 0; a < 3; a++) cent[3 * i + a] = 0.5f * (p[i].lo[a] + p[i].hi[a]);
    }
    bn.resize(p.size() * 2 + 1);
    if (const char* e = getenv("BDPT_BVH_LEAF")) kMaxLeaf = std::max(1, std::min(3, atoi(e)));   // tuning knobs
    if (const char* e = getenv("BDPT_BVH_CP")) kPrimCost = (float)atof(e);
    if (const char* e = getenv("BDPT_BVH_BINS")) kBins = std::max(2, std::min(kMaxBins, atoi(e)));
    int hw = (int)std::thread::hardware_concurrency();
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) hw = std::max(1, hw / std::max(1, atoi(e)));   // one process per GPU: share the host
    if (const char* e = getenv("BDPT_BVH_THREADS")) hw = atoi(e);
    spare_threads = std::max(0, std::min(hw, 16) - 1);
  }

  // Nodes of at most kExactMax primitives: exact sweep SAH over the centroid order of each axis.  (Resetting and
  // sweeping 3 x 16 bins for every one of the ~2n small nodes was 80 % of the build time.)
  static constexpr int kExactMax = 12;
  int build_small(int me, int first, int count, const Box& cb) {
    if (count == 2) {   // nothing to choose
      const int l = build(first, 1), r = build(first + 1, 1);
      bn[me].left = l; bn[me].right = r;
      return me;
    }
    uint32_t ord[kExactMax], best_ord[kExactMax];
    float right_area[kExactMax];
    float best_cost = FLT_MAX; int best_split = -1;
    for (int a = 0; a < 3; a++) {
      if (!(cb.hi[a] - cb.lo[a] > 0.f)) continue;
      for (int i = 0; i < count; i++) ord[i] = idx[first + i];
      // insertion sort by centroid (stable: ties keep the incoming order)
      for (int i = 1; i < count; i++) {
        uint32_t v = ord[i]; float c = cent[3 * v + a]; int j = i - 1;
        while (j >= 0 && cent[3 * ord[j] + a] > c) { ord[j + 1] = ord[j]; j--; }
        ord[j + 1] = v;
      }
      Box acc; acc.reset();
      for (int i = count - 1; i >= 1; i--) { acc.grow(prims[ord[i]].lo, prims[ord[i]].hi); right_area[i] = acc.half_area(); }
      acc.reset();
      for (int i = 1; i < count; i++) {   // left = ord[0..i), right = ord[i..count)
        acc.grow(prims[ord[i - 1]].lo, prims[ord[i - 1]].hi);
        float cost = acc.half_area() * i + right_area[i] * (count - i);
        if (cost < best_cost) { best_cost = cost; best_split = i; for (int k = 0; k < count; k++) best_ord[k] = ord[k]; }
      }
    }
    int mid;
    if (best_split < 0) mid = first + count / 2;   // identical centroids: split by index
    else { for (int k = 0; k < count; k++) idx[first + k] = best_ord[k]; mid = first + best_split; }
    const int l = build(first, mid - first);
    const int r = build(mid, first + count - mid);
    bn[me].left = l; bn[me].right = r;
    return me;
  }

  // Bounds of the primitives and of their centroids over idx[first, first + count)
  void bounds(int first, int count, Box* box, Box* cb) const {
    box->reset(); cb->reset();
    for (int i = first; i < first + count; i++) {
      const BuildPrim& p = prims[idx[i]];
      box->grow(p.lo, p.hi);
      const float* c = &cent[3 * idx[i]];
      cb->grow(c, c);
    }
  }

  int build(int first, int count) {
    Box box, cb;
    bounds(first, count, &box, &cb);
    return build(first, count, box, cb);
  }

  // `box` / `cb`: bounds of the range's primitives / centroids (the parent computes them while it partitions)
  int build(int first, int count, const Box& box, const Box& cb) {
    BNode n;
    n.box = box;
    n.first = first; n.count = count;
    const int me = next_node.fetch_add(1);
    bn[me] = n;
    if (count == 1) return me;

    if (count <= kExactMax) return build_small(me, first, count, cb);

    // binned SAH over the three axes, one pass over the primitives
    float best_cost = FLT_MAX; int best_axis = -1, best_split = -1;
    Box bb[3][kMaxBins]; int bc[3][kMaxBins]; float kk[3]; bool use[3];
    for (int a = 0; a < 3; a++) {
      float ext = cb.hi[a] - cb.lo[a];
      use[a] = ext > 0.f;
      kk[a] = use[a] ? kBins * (1.f - 1e-6f) / ext : 0.f;
      for (int b = 0; b < kBins; b++) { bb[a][b].reset(); bc[a][b] = 0; }
    }
    for (int i = first; i < first + count; i++) {
      const uint32_t id = idx[i];
      const BuildPrim& p = prims[id];
      for (int a = 0; a < 3; a++) {
        if (!use[a]) continue;
        int b = std::min(kBins - 1, std::max(0, (int)((cent[3 * id + a] - cb.lo[a]) * kk[a])));
        bb[a][b].grow(p.lo, p.hi); bc[a][b]++;
      }
    }
    for (int a = 0; a < 3; a++) {
      if (!use[a]) continue;
      float la[kMaxBins]; int lc[kMaxBins];
      Box acc; acc.reset(); int cnt = 0;
      for (int b = 0; b < kBins; b++) { acc.grow(bb[a][b]); cnt += bc[a][b]; la[b] = acc.half_area(); lc[b] = cnt; }
      acc.reset(); cnt = 0;
      for (int b = kBins - 1; b >= 1; b--) {
        acc.grow(bb[a][b]); cnt += bc[a][b];
        if (lc[b - 1] == 0 || cnt == 0) continue;
        float cost = la[b - 1] * lc[b - 1] + acc.half_area() * cnt;
        if (cost < best_cost) { best_cost = cost; best_axis = a; best_split = b; }
      }
    }
    int mid;
    Box lbox, lcb, rbox, rcb;
    bool have_bounds = false;
    if (best_axis < 0) {
      mid = first + count / 2;  // identical centroids: split by index
    } else {
      // partition by bin and collect both halves' bounds in the same pass
      const float k = kk[best_axis], lo = cb.lo[best_axis];
      lbox.reset(); lcb.reset(); rbox.reset(); rcb.reset();
      int i = first, j = first + count - 1;
      while (i <= j) {
        const uint32_t id = idx[i];
        const int b = std::min(kBins - 1, std::max(0, (int)((cent[3 * id + best_axis] - lo) * k)));
        const BuildPrim& p = prims[id];
        const float* c = &cent[3 * id];
        if (b < best_split) { lbox.grow(p.lo, p.hi); lcb.grow(c, c); i++; }
        else { rbox.grow(p.lo, p.hi); rcb.grow(c, c); std::swap(idx[i], idx[j]); j--; }
      }
      mid = i;
      have_bounds = true;
      if (mid == first || mid == first + count) { mid = first + count / 2; have_bounds = false; }
    }
    if (!have_bounds) { bounds(first, mid - first, &lbox, &lcb); bounds(mid, first + count - mid, &rbox, &rcb); }
    // the two halves touch disjoint ranges of idx and claim their nodes from the atomic counter: the left one may
    // run on another thread.  The tree does not depend on which thread built what (node numbers do; the collapse
    // below only follows left / right links).
    int l, r;
    const int nl = mid - first, nr = first + count - mid;
    if (std::min(nl, nr) >= kParallelMin && spare_threads.fetch_sub(1) > 0) {
      std::thread t([&] { l = build(first, nl, lbox, lcb); });
      r = build(mid, nr, rbox, rcb);
      t.join();
      spare_threads.fetch_add(1);
    } else {
      if (std::min(nl, nr) >= kParallelMin) spare_threads.fetch_add(1);   // undo the failed claim
      l = build(first, nl, lbox, lcb);
      r = build(mid, nr, rbox, rcb);
    }
    bn[me].left = l; bn[me].right = r;
    return me;
  }
};

inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

}  // namespace

void build_bvh8(const std::vector<BuildPrim>& prims, Bvh8* out) {
  out->nodes.clear(); out->prim_order.clear(); out->max_depth = 1;
  if (prims.empty()) { out->nodes.assign(20, 0u); return; }
  auto T0 = std::chrono::steady_clock::now();
  Builder B(prims);
  auto T1 = std::chrono::steady_clock::now();
  int root = B.build(0, (int)prims.size());
  auto T2 = std::chrono::steady_clock::now();

  // ---- SAH-optimal collapse of the binary tree into 8-wide nodes (Ylitie, Karras, Laine 2017, section 3.2).
  // C(n, i): cheapest way to represent the subtree of binary node n as at most i roots (i = 1..7).  One root is
  // either a leaf child (<= kMaxLeaf primitives) or a wide node whose <= 8 children are the roots its two subtrees
  // are distributed into:  C(n,1) = min(A_n P_n c_p,  A_n + min_k C(l,k) + C(r,8-k)),
  // C(n,i) = min(C(n,i-1), min_k C(l,k) + C(r,i-k)).  The greedy "open the largest child" collapse this replaces
  // left the average node half empty (4.1 of 8 slots on CBbunny, a quarter of the nodes with two children).
  const int nbn = B.next_node.load();
  std::vector<float> C((size_t)nbn * 8, FLT_MAX);          // C[n*8 + i], i = 1..7
  std::vector<uint8_t> split((size_t)nbn * 9, 0);          // split[n*9 + j], j = 2..8: roots given to the left subtree; 0 = "use j-1 roots"
  std::vector<uint8_t> is_leaf((size_t)nbn, 0);            // the one-root form of n is a leaf child
  {
    std::vector<int> order; order.reserve(nbn);
    std::vector<int> st; st.push_back(root);
    while (!st.empty()) { int n = st.back(); st.pop_back(); order.push_back(n); if (B.bn[n].left >= 0) { st.push_back(B.bn[n].left); st.push_back(B.bn[n].right); } }
    for (size_t oi = order.size(); oi-- > 0;) {            // children before parents
      const int n = order[oi];
      const BNode& bnode = B.bn[n];
      const float A = bnode.box.half_area();
      float* c = &C[(size_t)n * 8];
      if (bnode.left < 0) {                                // a single primitive
        for (int i = 1; i <= 7; i++) c[i] = A * B.kPrimCost;
        is_leaf[n] = 1;
        continue;
      }
      const float* cl = &C[(size_t)bnode.left * 8];
      const float* cr = &C[(size_t)bnode.right * 8];
      uint8_t* sp = &split[(size_t)n * 9];
      float dist[9];
      for (int j = 2; j <= 8; j++) {
        float best = FLT_MAX; int bk = 1;
        for (int k = std::max(1, j - 7); k <= std::min(7, j - 1); k++) {
          float v = cl[k] + cr[j - k];
          if (v < best) { best = v; bk = k; }
        }
        dist[j] = best; sp[j] = (uint8_t)bk;
      }
      const float c_leaf = bnode.count <= B.kMaxLeaf ? A * (float)bnode.count * B.kPrimCost : FLT_MAX;
      const float c_int = dist[8] + A;
      if (c_leaf <= c_int) { c[1] = c_leaf; is_leaf[n] = 1; } else { c[1] = c_int; }
      for (int i = 2; i <= 7; i++) {
        if (dist[i] < c[i - 1]) c[i] = dist[i];
        else { c[i] = c[i - 1]; sp[i] = 0; }
      }
    }
  }
  // the roots that represent binary node n with a budget of i
  struct Collect {
    const Builder& B; const std::vector<uint8_t>& split;
    void run(int n, int i, int* out, int& nout) const {
      while (i > 1 && B.bn[n].left >= 0 && split[(size_t)n * 9 + i] == 0) i--;
      if (i == 1 || B.bn[n].left < 0) { out[nout++] = n; return; }
      const int k = split[(size_t)n * 9 + i];
      run(B.bn[n].left, k, out, nout);
      run(B.bn[n].right, i - k, out, nout);
    }
  } collect{B, split};

  auto T2b = std::chrono::steady_clock::now();
  struct Item { int wide; int bnode; uint32_t depth; };
  std::deque<Item> queue;
  out->nodes.assign(20, 0u);
  queue.push_back({0, root, 1});
  while (!queue.empty()) {
    Item it = queue.front(); queue.pop_front();
    out->max_depth = std::max(out->max_depth, it.depth);
    const BNode& top = B.bn[it.bnode];
    int ch[8]; int nch = 0;
    if (top.left < 0) { ch[nch++] = it.bnode; }            // a one-primitive scene
    else {
      const int k = split[(size_t)it.bnode * 9 + 8];
      collect.run(top.left, k, ch, nch);
      collect.run(top.right, 8 - k, ch, nch);
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
    {
      // optimal assignment (maximum total of cost[child][slot]): Hungarian method on the 8 x 8 matrix, absent
      // children padded with zero rows
      const int N = 8;
      double a[N + 1][N + 1];
      for (int i = 1; i <= N; i++) for (int j = 1; j <= N; j++) a[i][j] = i <= nch ? -(double)cost[i - 1][j - 1] : 0.0;
      double u[N + 1] = {0}, v[N + 1] = {0}; int pm[N + 1] = {0}, way[N + 1] = {0};
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
      for (int j = 1; j <= N; j++) if (pm[j] >= 1 && pm[j] <= nch) slot_child[j - 1] = pm[j] - 1;
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
      if (!is_leaf[ch[c]]) {
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
  if (getenv("BDPT_BVH_TIMING")) {
    auto T3 = std::chrono::steady_clock::now();
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    fprintf(stderr, "bvh8: setup %.1f ms, sah %.1f ms, collapse dp %.1f ms, emit %.1f ms\n", ms(T0, T1), ms(T1, T2), ms(T2, T2b), ms(T2b, T3));
  }
}

}  // namespace bdpt
