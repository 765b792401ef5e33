// bvh8.cpp — binned-SAH binary build, collapse to 8-wide, octant slot assignment, 8-bit
// quantisation.  Host code; see bvh8.h for the node layout and for how this relates to the
// reference's BVHAccel (reference src/scene/bvh.cpp:51-129).
#include "bvh8.h"
#include "bvh_ploc.h"

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

// ---- tree views for bvh8_wide.h
struct SahView {
  const Builder& B;
  void box(int n, float* lo, float* hi) const { for (int a = 0; a < 3; a++) { lo[a] = B.bn[n].box.lo[a]; hi[a] = B.bn[n].box.hi[a]; } }
  int left(int n) const { return B.bn[n].left; }
  int right(int n) const { return B.bn[n].right; }
  int count(int n) const { return B.bn[n].count; }
  void gather(int n, uint32_t* out) const { for (int i = 0; i < B.bn[n].count; i++) out[i] = B.idx[B.bn[n].first + i]; }
};

// Binary tree as flat arrays: what the device builder produces (bvh_gpu.cuh) and build_bvh8_ploc restates on the host.
struct FlatView {
  const float* lo4; const float* hi4;      // 4 floats per node
  const int* l; const int* r; const int* cnt;
  const uint32_t* leaf_prim;               // nodes [0, n_leaves) are the leaves: their primitive
  void box(int n, float* lo, float* hi) const { for (int a = 0; a < 3; a++) { lo[a] = lo4[4 * (size_t)n + a]; hi[a] = hi4[4 * (size_t)n + a]; } }
  int left(int n) const { return l[n]; }
  int right(int n) const { return r[n]; }
  int count(int n) const { return cnt[n]; }
  void gather(int n, uint32_t* out) const {   // at most 3 primitives: a subtree of at most 5 nodes
    int st[4], sp = 0, k = 0;
    st[sp++] = n;
    while (sp > 0) {
      const int c = st[--sp];
      if (l[c] < 0) out[k++] = leaf_prim[c];
      else { st[sp++] = r[c]; st[sp++] = l[c]; }
    }
  }
};

// The collapse and the emission, for any tree view: SAH-optimal cut into 8-wide nodes bottom-up (`order`: parents
// before children), then the wide nodes breadth first (children in slot order get consecutive node indices).
template <class T>
void emit_wide_tree(const T& view, int root, const std::vector<int>& order, int n_bin, const WideParams& P, Bvh8* out) {
  std::vector<float> C((size_t)n_bin * 8, FLT_MAX);          // C[n*8 + i], i = 1..7
  std::vector<uint8_t> split((size_t)n_bin * 9, 0);          // split[n*9 + j], j = 2..8
  std::vector<uint8_t> is_leaf((size_t)n_bin, 0);            // the one-root form of n is a leaf child
  for (size_t oi = order.size(); oi-- > 0;) wide_dp_node(view, order[oi], P, C.data(), split.data(), is_leaf.data());   // children before parents
  struct Item { int wide; int bnode; uint32_t depth; };
  std::deque<Item> queue;
  out->nodes.assign(20, 0u);
  out->prim_order.assign((size_t)view.count(root), 0u);
  uint32_t n_prims_out = 0;
  queue.push_back({0, root, 1});
  while (!queue.empty()) {
    Item it = queue.front(); queue.pop_front();
    out->max_depth = std::max(out->max_depth, it.depth);
    WidePlan plan;
    wide_plan_node(view, split.data(), is_leaf.data(), it.bnode, &plan);
    const uint32_t child_base = (uint32_t)(out->nodes.size() / 20);
    const uint32_t prim_base = n_prims_out;
    n_prims_out += (uint32_t)plan.n_leaf_prims;
    out->nodes.resize(out->nodes.size() + 20 * (size_t)plan.n_internal, 0u);
    int next_bn[8];
    wide_write_node(view, is_leaf.data(), it.bnode, plan, child_base, prim_base, &out->nodes[(size_t)it.wide * 20], out->prim_order.data(), next_bn);
    for (int k = 0; k < plan.n_internal; k++) queue.push_back({(int)child_base + k, next_bn[k], it.depth + 1});
  }
}

WideParams wide_params_from_env() {
  WideParams P; P.max_leaf = 3; P.prim_cost = 0.8f;   // 0.4 ... 1.0: within 1.5 % (measured on B200, round 1)
  if (const char* e = getenv("BDPT_BVH_LEAF")) P.max_leaf = std::max(1, std::min(3, atoi(e)));   // tuning knobs
  if (const char* e = getenv("BDPT_BVH_CP")) P.prim_cost = (float)atof(e);
  return P;
}

}  // namespace

void build_bvh8(const std::vector<BuildPrim>& prims, Bvh8* out) {
  out->nodes.clear(); out->prim_order.clear(); out->max_depth = 1;
  if (prims.empty()) { out->nodes.assign(20, 0u); return; }
  auto T0 = std::chrono::steady_clock::now();
  Builder B(prims);
  auto T1 = std::chrono::steady_clock::now();
  int root = B.build(0, (int)prims.size());
  auto T2 = std::chrono::steady_clock::now();
  const int nbn = B.next_node.load();
  std::vector<int> order; order.reserve(nbn);
  {
    std::vector<int> st; st.push_back(root);
    while (!st.empty()) { int n = st.back(); st.pop_back(); order.push_back(n); if (B.bn[n].left >= 0) { st.push_back(B.bn[n].left); st.push_back(B.bn[n].right); } }
  }
  WideParams P; P.max_leaf = B.kMaxLeaf; P.prim_cost = B.kPrimCost;
  SahView view{B};
  emit_wide_tree(view, root, order, nbn, P, out);
  if (getenv("BDPT_BVH_TIMING")) {
    auto T3 = std::chrono::steady_clock::now();
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    fprintf(stderr, "bvh8: setup %.1f ms, sah %.1f ms, collapse + emit %.1f ms\n", ms(T0, T1), ms(T1, T2), ms(T2, T3));
  }
}

// Host restatement of the DEVICE build (bvh_gpu.cuh): the same per-element steps (bvh_ploc.h, bvh8_wide.h) run as
// plain loops.  It exists so that the CPU tests can check the algorithm's logic — tree soundness, quality against the
// binned-SAH build — where there is no GPU; the product builds on the device.
void build_bvh8_ploc(const std::vector<BuildPrim>& prims, Bvh8* out) {
  out->nodes.clear(); out->prim_order.clear(); out->max_depth = 1;
  const int N = (int)prims.size();
  if (N == 0) { out->nodes.assign(20, 0u); return; }
  // Morton order of the centroids
  float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  std::vector<float> cent((size_t)N * 3);
  for (int i = 0; i < N; i++)
    for (int a = 0; a < 3; a++) {
      const float c = 0.5f * (prims[i].lo[a] + prims[i].hi[a]);
      cent[3 * (size_t)i + a] = c; clo[a] = std::min(clo[a], c); chi[a] = std::max(chi[a], c);
    }
  std::vector<std::pair<uint64_t, uint32_t>> keys(N);
  for (int i = 0; i < N; i++) keys[i] = {ploc_morton(&cent[3 * (size_t)i], clo, chi), (uint32_t)i};
  std::stable_sort(keys.begin(), keys.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  const int n_bin = 2 * N - 1;
  std::vector<float> lo4((size_t)n_bin * 4, 0.f), hi4((size_t)n_bin * 4, 0.f);
  std::vector<int> l(n_bin, -1), r(n_bin, -1), cnt(n_bin, 1);
  std::vector<uint32_t> leaf_prim(N);
  for (int k = 0; k < N; k++) {
    const BuildPrim& p = prims[keys[k].second];
    leaf_prim[k] = keys[k].second;
    for (int a = 0; a < 3; a++) { lo4[4 * (size_t)k + a] = p.lo[a]; hi4[4 * (size_t)k + a] = p.hi[a]; }
  }
  // PLOC rounds over the cluster array (cluster boxes kept in array order, as the device does)
  std::vector<int> cluster(N), next_cluster(N), nn(N);
  std::vector<float> clo4((size_t)N * 4), chi4((size_t)N * 4), nlo4((size_t)N * 4), nhi4((size_t)N * 4);
  for (int k = 0; k < N; k++) { cluster[k] = k; for (int a = 0; a < 4; a++) { clo4[4 * (size_t)k + a] = lo4[4 * (size_t)k + a]; chi4[4 * (size_t)k + a] = hi4[4 * (size_t)k + a]; } }
  int n = N, next_node = N;
  while (n > 1) {
    for (int i = 0; i < n; i++) nn[i] = ploc_nearest(i, n, clo4.data(), chi4.data());
    int m = 0;
    for (int i = 0; i < n; i++) {
      const int j = nn[i];
      const bool mutual = nn[j] == i;
      if (mutual && i > j) continue;       // merged into position j
      int id = cluster[i];
      float blo[4], bhi[4];
      for (int a = 0; a < 4; a++) { blo[a] = clo4[4 * (size_t)i + a]; bhi[a] = chi4[4 * (size_t)i + a]; }
      if (mutual) {
        id = next_node++;
        l[id] = cluster[i]; r[id] = cluster[j]; cnt[id] = cnt[cluster[i]] + cnt[cluster[j]];
        for (int a = 0; a < 3; a++) { blo[a] = std::min(blo[a], clo4[4 * (size_t)j + a]); bhi[a] = std::max(bhi[a], chi4[4 * (size_t)j + a]); }
        for (int a = 0; a < 4; a++) { lo4[4 * (size_t)id + a] = blo[a]; hi4[4 * (size_t)id + a] = bhi[a]; }
      }
      next_cluster[m] = id;
      for (int a = 0; a < 4; a++) { nlo4[4 * (size_t)m + a] = blo[a]; nhi4[4 * (size_t)m + a] = bhi[a]; }
      m++;
    }
    cluster.swap(next_cluster); clo4.swap(nlo4); chi4.swap(nhi4);
    n = m;
  }
  const int root = cluster[0];
  std::vector<int> order; order.reserve(n_bin);
  {
    std::vector<int> st; st.push_back(root);
    while (!st.empty()) { int c = st.back(); st.pop_back(); order.push_back(c); if (l[c] >= 0) { st.push_back(l[c]); st.push_back(r[c]); } }
  }
  FlatView view{lo4.data(), hi4.data(), l.data(), r.data(), cnt.data(), leaf_prim.data()};
  emit_wide_tree(view, root, order, n_bin, wide_params_from_env(), out);
}

}  // namespace bdpt
