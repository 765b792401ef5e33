// api.cu — the C-ABI of libbdpt_b200.so (include/bdpt_b200.h): context, scene upload (host BVH
// build + fp32 narrowing), batch scheduling of the three wavefront kernels, frame buffers, stats
// and the parity test hooks.  No CPU fallback: without a CUDA device every call fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bdpt_b200.h"
#include "bvh8.h"
#include "bvh_gpu.cuh"
#include "kernels.cuh"

using namespace bdpt;

namespace {
thread_local std::string g_create_error = "";
}

// ---- build-time tuning of the persistent kernels (A/B: scripts/gpu_ab_lib.sh)
// k_walk_pt's thresholds: lanes that must want shading before the warp shades / idle lanes it tolerates before it fetches
#ifndef BDPT_WALK_ST
#define BDPT_WALK_ST 14
#endif
#ifndef BDPT_WALK_FT
#define BDPT_WALK_FT 6
#endif
#ifndef BDPT_WALK_MINB
#define BDPT_WALK_MINB 6   // resident CTAs/SM k_walk_pt is compiled for (80 registers)
#endif

#ifndef BDPT_SHADOW_MINB
#define BDPT_SHADOW_MINB 9    // resident CTAs/SM the triangle build of k_shadow_pt is compiled for: 56 registers, nothing spilled
                              // (10: 48 registers, 36 B of spills in the hot loop, +3.8 %; 8: +0.2 %; 11: +15 %)
#endif
#ifndef BDPT_SHADOW_GEN_MINB
#define BDPT_SHADOW_GEN_MINB 9   // ... and the build with the sphere test compiled in (also 56 registers, nothing spilled; 8: +2.3 % on CBspheres)
#endif
#define BDPT_SMALL_SCENE 1024u   // primitives up to which the host builds the (same) PLOC tree; crossover measured near 1,500

struct BdptContext {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;

  // scene
  bool have_scene = false;
  BdptCamera cam_base{};
  DScene sc{};
  uint4* d_nodes = nullptr;        // the builders' output: 80 bytes a node (what bdpt_device_bvh exports)
  uint4* d_nodes_wide = nullptr;   // the same nodes at a stride of 96 bytes: what the kernels traverse (DScene::nodes)
  size_t cap_nodes_wide = 0;
  float4 *d_prims = nullptr, *d_shading = nullptr, *d_mats = nullptr, *d_lights = nullptr;
  double* d_prims64 = nullptr;
  size_t cap_scene[6] = {0, 0, 0, 0, 0, 0};
  uint32_t bvh_nodes = 0, bvh_depth = 0;
  double ms_bvh_build = 0;
  float4* d_envmap = nullptr; int env_w = 0, env_h = 0;   // lat-long map of the environment light (bdpt_set_environment_map)
  // device build (bvh_gpu.cuh): workspace (grow-only), pinned counters, the leaf order it produced
  char* bvh_ws = nullptr; size_t bvh_ws_bytes = 0;
  unsigned int* h_bvh = nullptr;
  uint32_t* d_prim_order = nullptr;     // device build: points into bvh_ws; null after a host build, whose order is kept in
  std::vector<uint32_t> h_prim_order;   // ... here (what bdpt_device_bvh exports)
  // where the BVH is built: 2 (default) = on the device unless the scene has at most BDPT_SMALL_SCENE primitives, for which
  // the host restatement of the same PLOC build (same tree, node for node) is quicker than ~50 tiny launches
  // (252 primitives: 0.26 against 1.05 ms); 1 = always on the device (BDPT_BVH=device); 0 = binned SAH on the host
  // (BDPT_BVH=host, bvh8.cpp: the checker)
  int bvh_on_device = 2;
  uint32_t bvh_rounds = 0;              // PLOC rounds of the last device build

  // frame
  int W = 0, H = 0;
  DCamera cam{};
  // un-normalised accumulators: ONE allocation of 2 * W * H float4 {r, g, b, -}, eye image first, light image second,
  // so that a multi-GPU frame is reduced with a single collective (bdpt_reduce) and a splat is one 16-byte reduction
  // The kernels' reductions go to a second pair (acc_batch) that k_fold adds to `acc` after every batch with a
  // compensation term (acc_comp): all three live in the one allocation, `acc` first.
  float4* acc = nullptr;
  float4 *acc_eye = nullptr, *acc_light = nullptr;
  float4 *acc_comp = nullptr, *acc_batch = nullptr;
  float* resolve_tmp = nullptr;
  uint8_t* rgba_tmp = nullptr;
  int ns_last = 1;
  float lens_radius = 0.f, focal_dist = 4.7f;   // bdpt_set_lens
  int cell[4] = {0, 0, 0, 0};   // x, y, dx, dy of a cell render (bdpt_set_cell); dx = 0: the whole frame

  // batch storage
  size_t cap_B = 0; int cap_m = -1; size_t cap_q = 0;
  char* slab = nullptr;        // backing store of v0..v3, q0..q2, counts
  size_t slab_bytes = 0, slab_sz_v = 0, slab_sz_q = 0, slab_sz_c = 0;
  int slab_align_kb = 2048;    // BDPT_SLAB_ALIGN_KB, tuning only
  float4 *v0 = nullptr, *q0 = nullptr, *q2 = nullptr;
  int* counts = nullptr;
  uint2* chunks = nullptr;          // the shadow queue's chunk list (kernels.cuh, Job::chunks)
  size_t cap_chunks = 0;
  char* counters = nullptr;    // 2 MiB block holding qctr and d_stats
  unsigned int* qctr = nullptr;
  unsigned long long* d_stats = nullptr;
  unsigned int* h_qctr = nullptr;  // pinned

  // stats
  BdptStats stats{};
  std::vector<cudaEvent_t> ev;
  int sm_count = 148;
  // walk scheduler: 0 = one subpath per lane, 1 = persistent per-lane state machine; -1 = choose by BVH size at
  // upload (tiny trees have no traversal divergence to hide).  Override: BDPT_SCHED.  (A per-bounce wavefront walk
  // was measured in round 1 and removed: its state round trips cost more HBM time than the divergence it removed.)
  int sched = -1, sched_req = -1;
  int variant = 0;      // k_walk_pt build: 0 = 6 CTAs/SM (default), 1 = 4 CTAs/SM without spills (BDPT_VARIANT, tuning only)
  // BVH nodes the persistent kernels stage in shared memory (BDPT_TOP_NODES; 73 = root + 8 + 64 = three levels).
  // Off by default: measured on B200 it changes nothing (-0.5 % at 73, -1.6 % at 300) — L1 already holds the top
  // of the tree (77 % L1 hit rate in k_walk_pt).
  int top_nodes = 0;
  int connect_minb = 0; // __launch_bounds__ min blocks of k_connect: 0 = the build's own (BDPT_CONNECT_MINB, tuning only)
  int traits = 0;       // BDPT_TRAIT_* of the uploaded scene
  int point_lights = 0; // the uploaded scene has a point light
  int spec = 7;         // which kernels may use their trait-specialised build: 1 walk, 2 connect, 4 shadow (BDPT_SPEC, tuning only)
  // batch size cap (samples) and its memory budget (BDPT_BATCH_LOG2, BDPT_BATCH_GIB).  The persistent kernels end
  // with a tail of long paths and every batch starts with cold caches: CBbunny4 at 2^22 / 2^23 / 2^24 / 2^25 samples per
  // batch renders 927 / 966 / 988 / 999 M samples/s, and at m = 16 (16 KB per sample: the budget binds, not the cap)
  // CBgems gains 4 % from 48 -> 96 GiB.  96 of the 180 GB are there to be used; storage is only ever allocated for
  // the samples a render actually has, and a device that cannot give it gets a smaller batch (bdpt_render).
  int batch_log2 = 25, batch_gib = 96;
  int walk_ctas = BDPT_WALK_MINB, shadow_ctas = 0;  // resident CTAs per SM of the persistent kernels (shadow: 0 = the build's own, 9 or 8)
  // multi-GPU (comm.inc): an ncclComm_t, bound at run time
  void* comm = nullptr; int comm_ranks = 0, comm_rank = 0;
};

namespace {

#define CK(call)                                                                   \
  do {                                                                             \
    cudaError_t e_ = (call);                                                       \
    if (e_ != cudaSuccess) {                                                       \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);              \
      return BDPT_ERR_CUDA;                                                        \
    }                                                                              \
  } while (0)

int fail(BdptContext* ctx, int code, const std::string& msg) { ctx->err = msg; return code; }

template <class T>
void dfree(T*& p) { if (p) { cudaFree(p); p = nullptr; } }

float ibits(int i) { float f; memcpy(&f, &i, 4); return f; }
float f_down(double v) { float f = (float)v; if ((double)f > v) f = std::nextafterf(f, -INFINITY); return f; }
float f_up(double v) { float f = (float)v; if ((double)f < v) f = std::nextafterf(f, INFINITY); return f; }

void inverse3(const double* c, double* r /* rows of the inverse */) {
  // c: three column vectors
  const double *a = c, *b = c + 3, *cc = c + 6;
  double r0[3] = {b[1] * cc[2] - b[2] * cc[1], b[2] * cc[0] - b[0] * cc[2], b[0] * cc[1] - b[1] * cc[0]};
  double r1[3] = {cc[1] * a[2] - cc[2] * a[1], cc[2] * a[0] - cc[0] * a[2], cc[0] * a[1] - cc[1] * a[0]};
  double r2[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
  double det = a[0] * r0[0] + a[1] * r0[1] + a[2] * r0[2];
  for (int k = 0; k < 3; k++) { r[k] = r0[k] / det; r[3 + k] = r1[k] / det; r[6 + k] = r2[k] / det; }
}

void make_dcamera(const BdptCamera& c, int W, int H, DCamera* out) {
  const double PI = 3.14159265358979323;
  double hfov = c.hfov_deg, vfov = c.vfov_deg;
  if (W != (int)c.screen_w || H != (int)c.screen_h) {
    // Camera::set_screen_size, reference camera.cpp:83-89
    hfov = 2 * (std::atan(((double)W) / (2 * c.screen_dist)) * (180 / PI));
    vfov = 2 * (std::atan(((double)H) / (2 * c.screen_dist)) * (180 / PI));
  }
  out->pos = mk3((float)c.pos[0], (float)c.pos[1], (float)c.pos[2]);
  out->c2w_x = mk3((float)c.c2w[0], (float)c.c2w[1], (float)c.c2w[2]);
  out->c2w_y = mk3((float)c.c2w[3], (float)c.c2w[4], (float)c.c2w[5]);
  out->c2w_z = mk3((float)c.c2w[6], (float)c.c2w[7], (float)c.c2w[8]);
  double inv[9];
  inverse3(c.c2w, inv);
  out->w2c_r0 = mk3((float)inv[0], (float)inv[1], (float)inv[2]);
  out->w2c_r1 = mk3((float)inv[3], (float)inv[4], (float)inv[5]);
  out->w2c_r2 = mk3((float)inv[6], (float)inv[7], (float)inv[8]);
  out->tan_h = (float)std::tan(hfov * PI / 360);
  out->tan_v = (float)std::tan(vfov * PI / 360);
  out->nclip = (float)c.nclip; out->fclip = (float)c.fclip;
  out->W = W; out->H = H;
  out->lens_radius = 0.f; out->focal_dist = 1.f;   // (bdpt_set_lens fills these in)
}

// fp32 bounds of what the kernels actually test: the fp32-narrowed triangle, a slightly padded sphere
void make_build_prims(const BdptPrim* prims, uint32_t n_prims, std::vector<BuildPrim>* out) {
  std::vector<BuildPrim>& bp = *out;
  bp.resize(n_prims);
  for (uint32_t i = 0; i < n_prims; i++) {
    const double* d = prims[i].data;
    if (prims[i].type == BDPT_PRIM_TRIANGLE) {
      for (int a = 0; a < 3; a++) {
        float v0 = (float)d[a], v1 = (float)d[3 + a], v2 = (float)d[6 + a];
        bp[i].lo[a] = std::min(v0, std::min(v1, v2));
        bp[i].hi[a] = std::max(v0, std::max(v1, v2));
      }
    } else {
      double r = std::fabs(d[3]) * (1.0 + 1e-6) + 1e-7;
      for (int a = 0; a < 3; a++) { bp[i].lo[a] = f_down(d[a] - r); bp[i].hi[a] = f_up(d[a] + r); }
    }
  }
}

// Points the context's vertex / queue / count arrays into a slab.
void carve_slab(BdptContext* ctx, char* slab) {
  const size_t align = (size_t)ctx->slab_align_kb << 10;
  char* base = (char*)(((uintptr_t)slab + align - 1) / align * align);
  const size_t sz_v = ctx->slab_sz_v, sz_q = ctx->slab_sz_q;
  // one 64-byte record per vertex (v0..v3 are its four float4): a lane's four stores in the walk fill two whole
  // 32-byte sectors (as four separate planes they were partial-sector writes and DRAM saw 2x the bytes)
  ctx->v0 = (float4*)base;
  // shadow queue: 32-byte ray records (q0, stride 2), contributions q2 as a plane
  ctx->q0 = (float4*)(base + 4 * sz_v); ctx->q2 = (float4*)(base + 4 * sz_v + 2 * sz_q);
  ctx->counts = (int*)(base + 4 * sz_v + 3 * sz_q);
  ctx->chunks = (uint2*)(base + 4 * sz_v + 3 * sz_q + ctx->slab_sz_c);
}

int ensure_batch_storage(BdptContext* ctx, size_t B, int m, size_t qcap) {
  if (B > ctx->cap_B || m > ctx->cap_m || qcap > ctx->cap_q) {
    // One slab for the vertex records, the counts and the shadow queue, carved at 2 MiB boundaries.  (Tiling the
    // records per warp, so that a warp reads one 24 KB window, measured 13 % slower: the path positions being far
    // apart is what spreads a warp's loads over the HBM channels.)
    const size_t nB = std::max(B, ctx->cap_B); const int nm = std::max(m, ctx->cap_m); const size_t nq = std::max(qcap, ctx->cap_q);
    dfree(ctx->slab);
    ctx->v0 = nullptr; ctx->q0 = ctx->q2 = nullptr; ctx->counts = nullptr; ctx->chunks = nullptr;
    ctx->cap_B = 0; ctx->cap_m = -1; ctx->cap_q = 0;   // nothing is held until the allocation below succeeds
    const size_t n = 2 * (size_t)(nm + 1) * nB;
    const size_t align = (size_t)ctx->slab_align_kb << 10;
    auto up = [&](size_t x) { return (x + align - 1) / align * align; };
    const size_t sz_v = up(n * sizeof(float4)), sz_q = up(nq * sizeof(float4)), sz_c = up(2 * nB * sizeof(int));
    // chunk list: every block's segment cut into chunks of 256 entries
    const size_t n_chunks = (nB / BDPT_CONNECT_NT + 1) * ((size_t)BDPT_CONNECT_NT * (nm + 1) * (nm + 1) / 256 + 1);
    const size_t sz_k = up(n_chunks * sizeof(uint2));
    const size_t total = 4 * sz_v + 3 * sz_q + sz_c + sz_k + align;
    CK(cudaMalloc(&ctx->slab, total));
    ctx->slab_bytes = total; ctx->slab_sz_v = sz_v; ctx->slab_sz_q = sz_q; ctx->slab_sz_c = sz_c; ctx->cap_chunks = n_chunks;
    carve_slab(ctx, ctx->slab);
    ctx->cap_B = nB; ctx->cap_m = nm; ctx->cap_q = nq;
  }
  if (!ctx->qctr) {
    // The queue counter is the hottest address of k_connect (one returning atomic per warp and strategy, ~12 M per
    // batch).  As a 16-byte cudaMalloc it shared a small-allocation page with the materials, the lights and the
    // statistics, all read or updated by the same kernel, and k_connect took 5.9 to 8.2 ms per 2^24 samples depending
    // on how those happened to sit together (different per context, stable within one).  In a block of its own,
    // statistics a megabyte away, it is 5.9 ms every time.
    CK(cudaMalloc(&ctx->counters, (size_t)2 << 20));
    ctx->qctr = (unsigned int*)ctx->counters;
    ctx->d_stats = (unsigned long long*)(ctx->counters + ((size_t)1 << 20));
    CK(cudaMemset(ctx->d_stats, 0, 4 * sizeof(unsigned long long)));
    CK(cudaMemset(ctx->qctr, 0, 8 * sizeof(unsigned int)));
    CK(cudaMallocHost(&ctx->h_qctr, 8 * sizeof(unsigned int)));
  }
  return BDPT_OK;
}

cudaEvent_t get_event(BdptContext* ctx, size_t i) {
  while (ctx->ev.size() <= i) { cudaEvent_t e; cudaEventCreate(&e); ctx->ev.push_back(e); }
  return ctx->ev[i];
}

void fill_job(BdptContext* ctx, const BdptRenderParams* p, Job* J) {
  memset(J, 0, sizeof(Job));
  J->sc = ctx->sc; J->cam = ctx->cam; J->point_lights = ctx->point_lights;
  J->cam.lens_radius = ctx->lens_radius; J->cam.focal_dist = ctx->focal_dist;
  J->v0 = ctx->v0; J->counts = ctx->counts;
  J->q0 = ctx->q0; J->q2 = ctx->q2; J->qctr = ctx->qctr; J->chunks = ctx->chunks;
  J->acc_eye = ctx->acc_batch; J->acc_light = ctx->acc_batch + (size_t)ctx->W * ctx->H; J->stats = ctx->d_stats; J->dbg = nullptr;
  // the walk stores a vertex before it tests the depth (bidirection.cpp:82-86): m = 0 behaves as m = 1
  J->seed = p->seed; J->m = std::max(1, p->max_ray_depth); J->first_sample = p->first_sample;
  const bool cell = ctx->cell[2] > 0;
  J->cx0 = cell ? ctx->cell[0] : 0; J->cy0 = cell ? ctx->cell[1] : 0;
  J->cx1 = cell ? ctx->cell[0] + ctx->cell[2] : ctx->W; J->cy1 = cell ? ctx->cell[1] + ctx->cell[3] : ctx->H;
  J->tiles_x = (J->cx1 - J->cx0 + 7) / 8;
  J->npix_pad = J->tiles_x * ((J->cy1 - J->cy0 + 3) / 4) * 32;
  J->div_tiles_x = make_fastdiv((uint32_t)J->tiles_x); J->div_npix_pad = make_fastdiv((uint32_t)J->npix_pad);
  J->roulette = (p->flags & BDPT_FLAG_ROULETTE) ? 1 : 0;
  J->no_mis = (p->flags & BDPT_FLAG_NO_MIS) ? 1 : 0;
  J->balance = (p->flags & BDPT_FLAG_BALANCE) ? 1 : 0;
  J->ntop = std::min(ctx->top_nodes, ctx->sc.n_nodes);
  J->only_t = p->only_t; J->only_s = p->only_s;
  J->single_x = -1; J->single_y = -1;
}

// Launches k_connect in the build the scene's traits allow (the trait never depends on the tuning knob).
template <bool TRACE, int TR>
void launch_connect(BdptContext* ctx, const Job& J, unsigned grid, int minb) {
  constexpr int NT = BDPT_CONNECT_NT;
  switch (minb) {
    case 3: k_connect<TRACE, 3, TR><<<grid, NT, 0, ctx->stream>>>(J); break;
    case 4: k_connect<TRACE, 4, TR><<<grid, NT, 0, ctx->stream>>>(J); break;
    case 5: k_connect<TRACE, 5, TR><<<grid, NT, 0, ctx->stream>>>(J); break;
    case 8: k_connect<TRACE, 8, TR><<<grid, NT, 0, ctx->stream>>>(J); break;
    default: k_connect<TRACE, 6, TR><<<grid, NT, 0, ctx->stream>>>(J); break;
  }
}

// Runs one batch on ctx->stream.  evbase >= 0: record 4 events around the three kernels.
template <bool TRACE>
int run_batch(BdptContext* ctx, Job& J, long evbase) {
  // queue tail, shadow cursor, walk cursor; the overflow flag (word 4) is per render
  CK(cudaMemsetAsync(ctx->qctr, 0, 4 * sizeof(unsigned int), ctx->stream));
  const int threads = 128;
  if (evbase >= 0) cudaEventRecord(get_event(ctx, evbase + 0), ctx->stream);
  const bool lens = ctx->lens_radius > 0.0;   // thin lens: the generic build that carries the extension
  const int wtr = lens ? BDPT_TRAIT_LENS : (ctx->spec & 1) ? ctx->traits : 0;
  if (ctx->sched) {
    const int g = ctx->sm_count * ctx->walk_ctas;
    // 6 CTAs/SM (80 registers, ~0.2 KB of spills) beats 4 CTAs at 101 registers by 6 %: the state machine is
    // latency-bound between votes.  BDPT_VARIANT=1 selects the unspilled build for comparison.
    const size_t shm = (size_t)J.ntop * 80;
    if (ctx->variant == 1 && !lens) k_walk_pt<16, 8, 4><<<g, threads, shm, ctx->stream>>>(J);
    else switch (wtr) {
      case 1: k_walk_pt<BDPT_WALK_ST, BDPT_WALK_FT, BDPT_WALK_MINB, 1><<<g, threads, shm, ctx->stream>>>(J); break;
      case 2: k_walk_pt<BDPT_WALK_ST, BDPT_WALK_FT, BDPT_WALK_MINB, 2><<<g, threads, shm, ctx->stream>>>(J); break;
      case 3: k_walk_pt<BDPT_WALK_ST, BDPT_WALK_FT, BDPT_WALK_MINB, 3><<<g, threads, shm, ctx->stream>>>(J); break;
      case 4: k_walk_pt<BDPT_WALK_ST, BDPT_WALK_FT, BDPT_WALK_MINB, 4><<<g, threads, shm, ctx->stream>>>(J); break;
      case 5: k_walk_pt<BDPT_WALK_ST, BDPT_WALK_FT, BDPT_WALK_MINB, 5><<<g, threads, shm, ctx->stream>>>(J); break;
      case BDPT_TRAIT_LENS: k_walk_pt<BDPT_WALK_ST, BDPT_WALK_FT, BDPT_WALK_MINB, BDPT_TRAIT_LENS><<<g, threads, shm, ctx->stream>>>(J); break;
      default: k_walk_pt<BDPT_WALK_ST, BDPT_WALK_FT, BDPT_WALK_MINB, 0><<<g, threads, shm, ctx->stream>>>(J); break;
    }
  } else {
    const unsigned g = (unsigned)((2 * J.B + threads - 1) / threads);
    switch (wtr) {
      case 1: k_walk<1><<<g, threads, 0, ctx->stream>>>(J); break;
      case 2: k_walk<2><<<g, threads, 0, ctx->stream>>>(J); break;
      case 3: k_walk<3><<<g, threads, 0, ctx->stream>>>(J); break;
      case 4: k_walk<4><<<g, threads, 0, ctx->stream>>>(J); break;
      case 5: k_walk<5><<<g, threads, 0, ctx->stream>>>(J); break;
      case BDPT_TRAIT_LENS: k_walk<BDPT_TRAIT_LENS><<<g, threads, 0, ctx->stream>>>(J); break;
      default: k_walk<0><<<g, threads, 0, ctx->stream>>>(J); break;
    }
  }
  if (evbase >= 0) cudaEventRecord(get_event(ctx, evbase + 1), ctx->stream);
  // k_connect at 5 CTAs/SM (96 registers, nothing spilled).  Re-measured after the division / square-root change made
  // the kernel shorter — per 2^24 samples, bounds 5 / 6 / 4 / 8: CBbunny4 3.47 / 3.51 / 3.85 / 4.38 ms, CBspheres 3.69 / 3.95
  // / 4.14, glass bunny 3.69 / 3.93 / 4.13, CBgems m=16 1.41 / 1.57 / 1.54 (before it: 6 won, 4.00 against 4.24 on CBbunny).
  // BDPT_CONNECT_MINB overrides the bound only, never the trait.
  const unsigned cgrid = (unsigned)((J.B + BDPT_CONNECT_NT - 1) / BDPT_CONNECT_NT);
  // a sample pushes at most (m+1)^2 shadow rays (t = 1..m+1, s = 1..m+1): a block's segment cannot overflow
  J.seg_cap = (unsigned)(BDPT_CONNECT_NT * (J.m + 1) * (J.m + 1));
  J.chunk_cap = (unsigned)std::min<size_t>(ctx->cap_chunks, 0xffffffffu);
  const int ctr = (!TRACE && (ctx->spec & 2)) ? (ctx->traits & (BDPT_TRAIT_LAMBERT | BDPT_TRAIT_NOMF)) : 0;
  if (lens) launch_connect<TRACE, BDPT_TRAIT_LENS>(ctx, J, cgrid, ctx->connect_minb ? ctx->connect_minb : 5);
  else if (ctr & BDPT_TRAIT_LAMBERT) launch_connect<TRACE, BDPT_TRAIT_LAMBERT>(ctx, J, cgrid, ctx->connect_minb ? ctx->connect_minb : 5);
  else if (ctr & BDPT_TRAIT_NOMF) launch_connect<TRACE, BDPT_TRAIT_NOMF>(ctx, J, cgrid, ctx->connect_minb ? ctx->connect_minb : 5);
  else launch_connect<TRACE, 0>(ctx, J, cgrid, ctx->connect_minb ? ctx->connect_minb : 5);
  if (evbase >= 0) cudaEventRecord(get_event(ctx, evbase + 2), ctx->stream);
  {
    // triangles only: 48 registers (28 B of spills) and 10 CTAs/SM; with the sphere test compiled in, 64 and 8.
    // (The persistent form wins even on one-node BVHs: 2.9 vs 4.9 ms per 2^24 samples.)
    const size_t shm = (size_t)J.ntop * 80;
    if ((ctx->spec & 4) && (ctx->traits & BDPT_TRAIT_TRIS)) {
      const int c = ctx->shadow_ctas > 0 ? ctx->shadow_ctas : BDPT_SHADOW_MINB;
      k_shadow_pt<BDPT_SHADOW_MINB, BDPT_TRAIT_TRIS><<<ctx->sm_count * c, threads, shm, ctx->stream>>>(J);
    } else {
      const int c = ctx->shadow_ctas > 0 ? ctx->shadow_ctas : BDPT_SHADOW_GEN_MINB;
      k_shadow_pt<BDPT_SHADOW_GEN_MINB><<<ctx->sm_count * c, threads, shm, ctx->stream>>>(J);
    }
  }
  if (evbase >= 0) cudaEventRecord(get_event(ctx, evbase + 3), ctx->stream);
  ctx->stats.kernel_launches += 3;
  CK(cudaGetLastError());
  return BDPT_OK;
}

int check_ready(BdptContext* ctx, bool need_frame) {
  if (!ctx) return BDPT_ERR_ARG;
  if (!ctx->have_scene) return fail(ctx, BDPT_ERR_ARG, "no scene uploaded (bdpt_upload_scene)");
  if (need_frame && (ctx->W <= 0 || ctx->H <= 0)) return fail(ctx, BDPT_ERR_ARG, "no frame size set (bdpt_set_frame)");
  return BDPT_OK;
}

// generic "host doubles in -> float kernel -> host doubles out" helper for the test hooks
template <class Launch>
int run_hook(BdptContext* ctx, const double* in, size_t n, int n_in, double* out, int n_out, Launch launch) {
  if (n == 0) return BDPT_OK;
  std::vector<float> hin(n * n_in), hout(n * n_out);
  for (size_t i = 0; i < n * (size_t)n_in; i++) hin[i] = (float)in[i];
  float *din = nullptr, *dout = nullptr;
  CK(cudaMalloc(&din, hin.size() * sizeof(float)));
  CK(cudaMalloc(&dout, hout.size() * sizeof(float)));
  CK(cudaMemcpyAsync(din, hin.data(), hin.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  launch(din, dout);
  ctx->stats.kernel_launches += 1;
  cudaError_t e = cudaMemcpyAsync(hout.data(), dout, hout.size() * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(din); cudaFree(dout);
  if (e != cudaSuccess) return fail(ctx, BDPT_ERR_CUDA, cudaGetErrorString(e));
  for (size_t i = 0; i < hout.size(); i++) out[i] = (double)hout[i];
  return BDPT_OK;
}


// ---- device build of the wide BVH and of the leaf-ordered scene records (bvh_gpu.cuh).  On return the scene arrays
// d_nodes / d_prims / d_shading / d_prims64 are filled (stream-ordered) and *n_nodes / *depth are known.
int ensure_dev(void** p, size_t* cap, size_t bytes) {
  if (bytes <= *cap && *p) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr; *cap = 0;
  if (cudaMalloc(p, bytes) != cudaSuccess) return 1;
  *cap = bytes;
  return 0;
}

int build_scene_on_device(BdptContext* ctx, const BdptPrim* prims, uint32_t n_prims, uint32_t* n_nodes, uint32_t* depth) {
  const int N = (int)n_prims;
  const size_t n2 = 2 * (size_t)N;
  cudaStream_t st = ctx->stream;
  // scene arrays (grow-only)
  if (ensure_dev((void**)&ctx->d_nodes, &ctx->cap_scene[0], (size_t)N * 80) || ensure_dev((void**)&ctx->d_prims, &ctx->cap_scene[1], BDPT_PRIM_Q * (size_t)N * sizeof(float4)) ||
      ensure_dev((void**)&ctx->d_shading, &ctx->cap_scene[2], 3 * (size_t)N * sizeof(float4)) || ensure_dev((void**)&ctx->d_prims64, &ctx->cap_scene[5], 10 * (size_t)N * sizeof(double)))
    return fail(ctx, BDPT_ERR_CUDA, "cudaMalloc (scene arrays)");
  // workspace
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, N, 0, 63, st);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
  const size_t o_raw = take((size_t)N * sizeof(BdptPrim)), o_blo = take(N * sizeof(float4)), o_bhi = take(N * sizeof(float4));
  const size_t o_k0 = take(N * 8), o_k1 = take(N * 8), o_v0 = take(N * 4), o_v1 = take(N * 4), o_sort = take(sort_bytes);
  const size_t o_tlo = take(n2 * sizeof(float4)), o_thi = take(n2 * sizeof(float4)), o_l = take(n2 * 4), o_r = take(n2 * 4), o_p = take(n2 * 4), o_c = take(n2 * 4), o_lp = take(N * 4);
  const size_t o_cl = take(N * 4), o_clo = take(N * sizeof(float4)), o_chi = take(N * sizeof(float4)), o_nn = take(N * 4);
  const size_t o_oid = take(N * 4), o_olo = take(N * sizeof(float4)), o_ohi = take(N * sizeof(float4)), o_keep = take(N);
  const size_t nb = (size_t)(N + BDPT_PLOC_BLOCK - 1) / BDPT_PLOC_BLOCK;
  const size_t o_bc = take(nb * 4), o_bo = take(nb * 4), o_ctr = take(64);
  const size_t o_C = take(n2 * 8 * sizeof(float)), o_split = take(n2 * 9), o_leaf = take(n2), o_visit = take(n2 * 4);
  const size_t o_it0 = take((size_t)N * sizeof(int2)), o_it1 = take((size_t)N * sizeof(int2)), o_po = take(N * 4);
  if (ensure_dev((void**)&ctx->bvh_ws, &ctx->bvh_ws_bytes, off)) return fail(ctx, BDPT_ERR_CUDA, "cudaMalloc (BVH build workspace)");
  char* W = ctx->bvh_ws;
  BdptPrim* d_raw = (BdptPrim*)(W + o_raw);
  float4 *blo = (float4*)(W + o_blo), *bhi = (float4*)(W + o_bhi);
  unsigned long long *k0 = (unsigned long long*)(W + o_k0), *k1 = (unsigned long long*)(W + o_k1);
  uint32_t *v0 = (uint32_t*)(W + o_v0), *v1 = (uint32_t*)(W + o_v1);
  DevTree T{(float4*)(W + o_tlo), (float4*)(W + o_thi), (int*)(W + o_l), (int*)(W + o_r), (int*)(W + o_p), (int*)(W + o_c), (uint32_t*)(W + o_lp)};
  int* cluster = (int*)(W + o_cl); float4 *clo = (float4*)(W + o_clo), *chi = (float4*)(W + o_chi); int* nn = (int*)(W + o_nn);
  int* out_id = (int*)(W + o_oid); float4 *olo = (float4*)(W + o_olo), *ohi = (float4*)(W + o_ohi); unsigned char* keep = (unsigned char*)(W + o_keep);
  unsigned *bcount = (unsigned*)(W + o_bc), *boff = (unsigned*)(W + o_bo), *ctr = (unsigned*)(W + o_ctr);
  float* C = (float*)(W + o_C); uint8_t *split = (uint8_t*)(W + o_split), *is_leaf = (uint8_t*)(W + o_leaf); unsigned* visit = (unsigned*)(W + o_visit);
  int2* items[2] = {(int2*)(W + o_it0), (int2*)(W + o_it1)};
  uint32_t* prim_order = (uint32_t*)(W + o_po);
  ctx->d_prim_order = prim_order;
  unsigned* cb = ctr + 8;      // centroid bounds, 6 encoded words
  unsigned* ectr = ctr + 4;    // emission counters

  const int TB = 256, G = (N + TB - 1) / TB;
  CK(cudaMemcpyAsync(d_raw, prims, (size_t)N * sizeof(BdptPrim), cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(cb, 0xff, 3 * sizeof(unsigned), st));
  CK(cudaMemsetAsync(cb + 3, 0, 3 * sizeof(unsigned), st));
  k_bvh_bounds<<<G, TB, 0, st>>>(d_raw, N, blo, bhi, cb);
  k_bvh_morton<<<G, TB, 0, st>>>(blo, bhi, N, cb, k0, v0);
  cub::DeviceRadixSort::SortPairs(W + o_sort, sort_bytes, (const unsigned long long*)k0, k1, (const uint32_t*)v0, v1, N, 0, 63, st);
  k_bvh_leaves<<<G, TB, 0, st>>>(blo, bhi, v1, N, T, cluster, clo, chi);
  ctx->stats.kernel_launches += 4;
  ctx->h_bvh[0] = (unsigned)N; ctx->h_bvh[1] = (unsigned)N; ctx->h_bvh[2] = (unsigned)N; ctx->h_bvh[3] = 0;
  CK(cudaMemcpyAsync(ctr, ctx->h_bvh, 4 * sizeof(unsigned), cudaMemcpyHostToDevice, st));
  // PLOC rounds; the cluster count comes back every few rounds (the kernels of a finished build are no-ops)
  int n_upper = N; uint32_t rounds = 0;
  int radius = BDPT_PLOC_RADIUS;
  if (const char* e = getenv("BDPT_PLOC_RADIUS")) radius = std::max(1, std::min(256, atoi(e)));   // tuning only
  bool tail = false;
  while (n_upper > 1) {
    if (n_upper <= 1024) {   // the rest in one launch (bvh_gpu.cuh, k_ploc_tail); its round count comes back with the emission's
      k_ploc_tail<<<1, 1024, 0, st>>>(ctr, cluster, clo, chi, T, radius);
      CK(cudaMemcpyAsync(ctx->h_bvh + 12, ctr + 3, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
      ctx->stats.kernel_launches += 1; tail = true;
      break;
    }
    for (int r = 0; r < 4; r++) {
      const int g = (n_upper + BDPT_PLOC_BLOCK - 1) / BDPT_PLOC_BLOCK;
      k_ploc_nearest<<<g, BDPT_PLOC_BLOCK, 0, st>>>(ctr, clo, chi, nn, radius);
      k_ploc_merge<<<g, BDPT_PLOC_BLOCK, 0, st>>>(ctr, nn, cluster, clo, chi, T, out_id, olo, ohi, keep, bcount);
      k_ploc_scan<<<1, 1024, 0, st>>>(ctr, bcount, boff);
      k_ploc_scatter<<<g, BDPT_PLOC_BLOCK, 0, st>>>(ctr, keep, boff, out_id, olo, ohi, cluster, clo, chi);
      ctx->stats.kernel_launches += 4; rounds++;
    }
    CK(cudaMemcpyAsync(ctx->h_bvh + 8, ctr, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int n_now = (int)ctx->h_bvh[8];
    if (n_now < 1 || n_now > n_upper) return fail(ctx, BDPT_ERR_CUDA, "device BVH build: cluster count went wrong");
    if (n_now == n_upper && n_now > 1 && rounds > 4096) return fail(ctx, BDPT_ERR_CUDA, "device BVH build does not converge");
    n_upper = n_now;
  }
  ctx->bvh_rounds = rounds;   // (+ the tail's, below)
  // collapse DP, bottom-up
  CK(cudaMemsetAsync(visit, 0, n2 * 4, st));
  WideParams P; P.max_leaf = 3; P.prim_cost = 0.8f;
  if (const char* e = getenv("BDPT_BVH_LEAF")) P.max_leaf = std::max(1, std::min(3, atoi(e)));
  if (const char* e = getenv("BDPT_BVH_CP")) P.prim_cost = (float)atof(e);
  k_bvh_collapse_dp<<<G, TB, 0, st>>>(T, N, P, C, split, is_leaf, visit);
  // wide nodes, one level per launch (items of level 0: {node 0, root = the last cluster})
  k_bvh_root<<<1, 1, 0, st>>>(cluster, items[0], ectr);
  ctx->stats.kernel_launches += 2;
  uint32_t n_items = 1, levels = 0; int cur = 0;
  while (n_items > 0) {
    if (levels >= BDPT_STACK) return fail(ctx, BDPT_ERR_SCENE, "BVH deeper than the traversal stack");
    k_bvh_emit_level<<<(n_items + 63) / 64, 64, 0, st>>>(T, split, is_leaf, items[cur], (int)n_items, (uint32_t*)ctx->d_nodes, prim_order, items[cur ^ 1], ectr);
    ctx->stats.kernel_launches += 1; levels++;
    CK(cudaMemcpyAsync(ctx->h_bvh + 8, ectr, 3 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CK(cudaMemsetAsync(ectr + 2, 0, sizeof(unsigned), st));
    CK(cudaStreamSynchronize(st));
    n_items = ctx->h_bvh[10];
    if (ctx->h_bvh[8] > (unsigned)N || n_items > (unsigned)N) return fail(ctx, BDPT_ERR_CUDA, "device BVH build: node count went wrong");
    cur ^= 1;
  }
  if (ctx->h_bvh[9] != (unsigned)N) return fail(ctx, BDPT_ERR_CUDA, "device BVH build: not every primitive reached a leaf");
  if (tail) ctx->bvh_rounds += ctx->h_bvh[12];   // (copied back before the emission's synchronisations)
  *n_nodes = ctx->h_bvh[8]; *depth = levels;
  k_bvh_records<<<G, TB, 0, st>>>(d_raw, prim_order, N, ctx->d_prims, ctx->d_shading, ctx->d_prims64);
  ctx->stats.kernel_launches += 1;
  CK(cudaGetLastError());
  return BDPT_OK;
}

}  // namespace

extern "C" {

int bdpt_create(int device, BdptContext** out) {
  if (!out) return BDPT_ERR_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                     " (libbdpt_b200 has no CPU fallback)";
    return BDPT_ERR_CUDA;
  }
  if (device < 0 || device >= n) { g_create_error = "bad device index"; return BDPT_ERR_ARG; }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return BDPT_ERR_CUDA; }
  BdptContext* ctx = new BdptContext();
  ctx->device = device;
  e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete ctx; return BDPT_ERR_CUDA; }
  ctx->own_stream = true;
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  // (measurement only: the L1 / shared-memory split the driver picks for kernels that use almost no shared memory)
  if (const char* e = getenv("BDPT_CACHE_CONFIG")) {
    const int c = atoi(e);
    cudaDeviceSetCacheConfig(c == 1 ? cudaFuncCachePreferL1 : c == 2 ? cudaFuncCachePreferShared : c == 3 ? cudaFuncCachePreferEqual : cudaFuncCachePreferNone);
  }
  if (const char* e = getenv("BDPT_SCHED")) ctx->sched_req = std::min(1, atoi(e));
  ctx->sched = ctx->sched_req >= 0 ? ctx->sched_req : 1;
  if (const char* e = getenv("BDPT_BATCH_LOG2")) ctx->batch_log2 = std::min(26, std::max(10, atoi(e)));
  if (const char* e = getenv("BDPT_BATCH_GIB")) ctx->batch_gib = std::min(144, std::max(1, atoi(e)));
  if (const char* e = getenv("BDPT_TOP_NODES")) ctx->top_nodes = std::max(0, std::min(512, atoi(e)));
  if (const char* e = getenv("BDPT_CONNECT_MINB")) ctx->connect_minb = atoi(e);
  if (const char* e = getenv("BDPT_SPEC")) ctx->spec = atoi(e);
  if (const char* e = getenv("BDPT_SLAB_ALIGN_KB")) ctx->slab_align_kb = std::max(1, atoi(e));
  if (const char* e = getenv("BDPT_VARIANT")) ctx->variant = atoi(e);
  if (const char* e = getenv("BDPT_WALK_CTAS")) ctx->walk_ctas = std::max(1, atoi(e));
  if (const char* e = getenv("BDPT_SHADOW_CTAS")) ctx->shadow_ctas = std::max(1, atoi(e));
  if (const char* e = getenv("BDPT_BVH")) ctx->bvh_on_device = !strcmp(e, "host") ? 0 : !strcmp(e, "device") ? 1 : 2;
  if (cudaMallocHost(&ctx->h_bvh, 16 * sizeof(unsigned int)) != cudaSuccess) { g_create_error = "cudaMallocHost failed"; cudaStreamDestroy(ctx->stream); delete ctx; return BDPT_ERR_CUDA; }
  *out = ctx;
  return BDPT_OK;
}

void bdpt_destroy(BdptContext* ctx) {
  if (!ctx) return;
  bdpt_comm_destroy(ctx);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  dfree(ctx->d_nodes); dfree(ctx->d_nodes_wide); dfree(ctx->d_prims); dfree(ctx->d_shading); dfree(ctx->d_mats); dfree(ctx->d_lights); dfree(ctx->d_prims64);
  dfree(ctx->acc); dfree(ctx->resolve_tmp); dfree(ctx->rgba_tmp);
  dfree(ctx->slab); dfree(ctx->counters); ctx->qctr = nullptr; ctx->d_stats = nullptr;
  if (ctx->h_qctr) cudaFreeHost(ctx->h_qctr);
  if (ctx->h_bvh) cudaFreeHost(ctx->h_bvh);
  dfree(ctx->bvh_ws); dfree(ctx->d_envmap);
  for (cudaEvent_t e : ctx->ev) cudaEventDestroy(e);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* bdpt_last_error(const BdptContext* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int bdpt_set_stream(BdptContext* ctx, void* cuda_stream) {
  if (!ctx) return BDPT_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->own_stream = false;
  return BDPT_OK;
}

int bdpt_upload_scene(BdptContext* ctx, const BdptCamera* camera, const BdptMaterial* materials, uint32_t n_materials,
                      const BdptLight* lights, uint32_t n_lights, const BdptPrim* prims, uint32_t n_prims) {
  if (!ctx || !camera || (!materials && n_materials) || (!lights && n_lights) || (!prims && n_prims)) return BDPT_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (n_lights == 0) return fail(ctx, BDPT_ERR_SCENE, "scene has no light: BDPT needs at least one (bidirection.cpp:109)");
  for (uint32_t i = 0; i < n_lights; i++)
    if (lights[i].type != BDPT_LIGHT_AREA && lights[i].type != BDPT_LIGHT_POINT && lights[i].type != BDPT_LIGHT_ENV)
      return fail(ctx, BDPT_ERR_SCENE, "light type not supported by BDPT (only area and point lights implement sample_Le, light.cpp)");
  for (uint32_t i = 0; i < n_materials; i++)
    if (materials[i].type < BDPT_MAT_DIFFUSE || materials[i].type > BDPT_MAT_MICROFACET)
      return fail(ctx, BDPT_ERR_SCENE, "unknown material type");
  for (uint32_t i = 0; i < n_prims; i++) {
    if (prims[i].material < 0 || (uint32_t)prims[i].material >= n_materials) return fail(ctx, BDPT_ERR_SCENE, "primitive references a missing material");
    if (prims[i].type != BDPT_PRIM_TRIANGLE && prims[i].type != BDPT_PRIM_SPHERE) return fail(ctx, BDPT_ERR_SCENE, "unknown primitive type");
  }

  auto t0 = std::chrono::steady_clock::now();
  CK(cudaStreamSynchronize(ctx->stream));
  // grow-only device storage: re-uploading a scene of the same size allocates nothing
  auto ensure = [&](void** p, size_t* cap, size_t bytes) -> cudaError_t {
    if (bytes <= *cap && *p) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaSuccess) *cap = bytes;
    return e;
  };
  // (one extra material: the emitter the environment's sphere carries, index n_materials)
  std::vector<float4> hm(5 * ((size_t)n_materials + 1), make_float4(0.f, 0.f, 0.f, 0.f));
  hm[5 * (size_t)n_materials + 0] = make_float4(0.f, 0.f, 0.f, ibits(BDPT_MAT_EMISSION));
  hm[5 * (size_t)n_materials + 2] = make_float4(1.f, 1.f, 1.f, 0.f);
  for (uint32_t i = 0; i < n_materials; i++) {
    const BdptMaterial& m = materials[i];
    hm[5 * i + 0] = make_float4((float)m.reflectance[0], (float)m.reflectance[1], (float)m.reflectance[2], ibits(m.type));
    hm[5 * i + 1] = make_float4((float)m.transmittance[0], (float)m.transmittance[1], (float)m.transmittance[2], (float)m.ior);
    hm[5 * i + 2] = make_float4((float)m.radiance[0], (float)m.radiance[1], (float)m.radiance[2], (float)m.alpha);
    hm[5 * i + 3] = make_float4((float)m.eta[0], (float)m.eta[1], (float)m.eta[2], 0.f);
    hm[5 * i + 4] = make_float4((float)m.k[0], (float)m.k[1], (float)m.k[2], 0.f);
  }
  std::vector<float4> hl(5 * (size_t)n_lights);
  int env_light = -1;
  for (uint32_t i = 0; i < n_lights; i++) {
    const BdptLight& l = lights[i];
    if (l.type == BDPT_LIGHT_ENV) {
      // the far sphere sits around the scene's bounding sphere (centre and half diagonal of the primitives' box)
      double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
      for (uint32_t k = 0; k < n_prims; k++) {
        const double* d = prims[k].data;
        const int nv = prims[k].type == BDPT_PRIM_TRIANGLE ? 3 : 1;
        const double r = prims[k].type == BDPT_PRIM_TRIANGLE ? 0.0 : d[3];
        for (int v = 0; v < nv; v++) for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], d[3 * v + a] - r); hi[a] = std::max(hi[a], d[3 * v + a] + r); }
      }
      if (n_prims == 0) for (int a = 0; a < 3; a++) { lo[a] = -1; hi[a] = 1; }
      const double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
      const double rs = std::max(1e-6, 0.5 * std::sqrt(dx * dx + dy * dy + dz * dz)), R = 1000.0 * rs;
      hl[5 * i + 0] = make_float4((float)l.radiance[0], (float)l.radiance[1], (float)l.radiance[2], ibits(l.type));
      hl[5 * i + 1] = make_float4((float)(0.5 * (lo[0] + hi[0])), (float)(0.5 * (lo[1] + hi[1])), (float)(0.5 * (lo[2] + hi[2])), (float)(4.0 * 3.14159265358979323 * R * R));
      hl[5 * i + 2] = make_float4(0.f, 0.f, 0.f, 0.f);
      hl[5 * i + 3] = make_float4((float)R, (float)rs, 0.f, 0.f);
      hl[5 * i + 4] = make_float4(0.f, 0.f, 0.f, 0.f);
      env_light = (int)i;
      continue;
    }
    double nx = std::sqrt(l.dim_x[0] * l.dim_x[0] + l.dim_x[1] * l.dim_x[1] + l.dim_x[2] * l.dim_x[2]);
    double ny = std::sqrt(l.dim_y[0] * l.dim_y[0] + l.dim_y[1] * l.dim_y[1] + l.dim_y[2] * l.dim_y[2]);
    hl[5 * i + 0] = make_float4((float)l.radiance[0], (float)l.radiance[1], (float)l.radiance[2], ibits(l.type));
    hl[5 * i + 1] = make_float4((float)l.position[0], (float)l.position[1], (float)l.position[2], (float)(nx * ny));  // light.cpp:199-203
    hl[5 * i + 2] = make_float4((float)l.direction[0], (float)l.direction[1], (float)l.direction[2], 0.f);
    hl[5 * i + 3] = make_float4((float)l.dim_x[0], (float)l.dim_x[1], (float)l.dim_x[2], 0.f);
    hl[5 * i + 4] = make_float4((float)l.dim_y[0], (float)l.dim_y[1], (float)l.dim_y[2], 0.f);
  }
  CK(ensure((void**)&ctx->d_mats, &ctx->cap_scene[3], hm.size() * sizeof(float4)));
  CK(ensure((void**)&ctx->d_lights, &ctx->cap_scene[4], hl.size() * sizeof(float4)));
  CK(cudaMemcpyAsync(ctx->d_mats, hm.data(), hm.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_lights, hl.data(), hl.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));

  uint32_t n_nodes = 0, depth = 0;
  const bool small_scene = ctx->bvh_on_device == 2 && n_prims <= BDPT_SMALL_SCENE;
  if (ctx->bvh_on_device && !small_scene && n_prims >= 1) {
    // the raw primitives go up as they are; bounds, Morton order, PLOC, the 8-wide collapse and the leaf-ordered
    // records are all made on the device (bvh_gpu.cuh) — what BVHAccel::construct_bvh does on one host thread
    int rc = build_scene_on_device(ctx, prims, n_prims, &n_nodes, &depth);
    if (rc) return rc;
  } else {
    // BDPT_BVH=host (and the empty scene): binned-SAH build on the host (bvh8.cpp), kept as the checker; small scenes:
    // the PLOC build, on the host
    std::vector<BuildPrim> bp;
    make_build_prims(prims, n_prims, &bp);
    Bvh8 bvh;
    {
      const char* algo = getenv("BDPT_BVH_HOST_ALGO");
      if (small_scene || (algo && !strcmp(algo, "ploc"))) build_bvh8_ploc(bp, &bvh); else build_bvh8(bp, &bvh);
    }
    if (bvh.max_depth >= BDPT_STACK) return fail(ctx, BDPT_ERR_SCENE, "BVH deeper than the traversal stack");
    // leaf-ordered primitive + shading records
    std::vector<float4> hp(BDPT_PRIM_Q * (size_t)std::max<uint32_t>(n_prims, 1)), hs(3 * (size_t)std::max<uint32_t>(n_prims, 1));
    std::vector<double> h64(10 * (size_t)std::max<uint32_t>(n_prims, 1), 0.0);
    for (size_t k = 0; k < bvh.prim_order.size(); k++)
      make_leaf_record(prims[bvh.prim_order[k]], bvh.prim_order[k], hp.data() + BDPT_PRIM_Q * k, hs.data() + 3 * k, h64.data() + 10 * k);
    CK(ensure((void**)&ctx->d_nodes, &ctx->cap_scene[0], bvh.nodes.size() * sizeof(uint32_t)));
    CK(ensure((void**)&ctx->d_prims, &ctx->cap_scene[1], hp.size() * sizeof(float4)));
    CK(ensure((void**)&ctx->d_shading, &ctx->cap_scene[2], hs.size() * sizeof(float4)));
    CK(ensure((void**)&ctx->d_prims64, &ctx->cap_scene[5], h64.size() * sizeof(double)));
    CK(cudaMemcpyAsync(ctx->d_prims64, h64.data(), h64.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_nodes, bvh.nodes.data(), bvh.nodes.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_prims, hp.data(), hp.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_shading, hs.data(), hs.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));   // the staging vectors go out of scope
    ctx->d_prim_order = nullptr; ctx->h_prim_order = bvh.prim_order;
    n_nodes = (uint32_t)(bvh.nodes.size() / 20); depth = bvh.max_depth;
  }
  // the traversal's copy of the nodes: 96 bytes apart, so that a node is three aligned 32-byte loads
  if (ensure_dev((void**)&ctx->d_nodes_wide, &ctx->cap_nodes_wide, (size_t)std::max<uint32_t>(n_nodes, 1) * BDPT_NODE_Q * sizeof(uint4)))
    return fail(ctx, BDPT_ERR_CUDA, "cudaMalloc (nodes)");
  if (n_nodes) CK(cudaMemcpy2DAsync(ctx->d_nodes_wide, BDPT_NODE_Q * sizeof(uint4), ctx->d_nodes, 80, 80, n_nodes, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  // (historical name: the whole scene ingest — build, records, copies — since the build moved to the device)
  ctx->ms_bvh_build = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();

  ctx->sc.nodes = ctx->d_nodes_wide; ctx->sc.prims = ctx->d_prims; ctx->sc.shading = ctx->d_shading; ctx->sc.prims64 = ctx->d_prims64;
  ctx->sc.materials = ctx->d_mats; ctx->sc.lights = ctx->d_lights;
  ctx->sc.n_lights = (int)n_lights; ctx->sc.n_materials = (int)n_materials;
  ctx->sc.env_light = env_light;
  ctx->sc.envmap = ctx->d_envmap; ctx->sc.env_w = ctx->env_w; ctx->sc.env_h = ctx->env_h;
  ctx->sc.qbias_bits = 0x47000000u;
  ctx->sc.n_nodes = (int)n_nodes; ctx->sc.n_prims = (int)n_prims;
  ctx->bvh_nodes = n_nodes; ctx->bvh_depth = depth;
  ctx->sched = ctx->sched_req >= 0 ? ctx->sched_req : (ctx->bvh_nodes > 8 ? 1 : 0);
  // what this scene promises the specialised kernels (common.cuh, BDPT_TRAIT_*)
  int traits = BDPT_TRAIT_TRIS | BDPT_TRAIT_LAMBERT;
  for (uint32_t i = 0; i < n_prims; i++) if (prims[i].type != BDPT_PRIM_TRIANGLE) { traits &= ~BDPT_TRAIT_TRIS; break; }
  for (uint32_t i = 0; i < n_materials; i++)
    if (materials[i].type != BDPT_MAT_DIFFUSE && materials[i].type != BDPT_MAT_EMISSION) { traits &= ~BDPT_TRAIT_LAMBERT; break; }
  if (!(traits & BDPT_TRAIT_LAMBERT)) {
    traits |= BDPT_TRAIT_NOMF;
    for (uint32_t i = 0; i < n_materials; i++) if (materials[i].type == BDPT_MAT_MICROFACET) { traits &= ~BDPT_TRAIT_NOMF; break; }
  }
  ctx->traits = traits;
  ctx->point_lights = 0;
  for (uint32_t i = 0; i < n_lights; i++) if (lights[i].type == BDPT_LIGHT_POINT) ctx->point_lights = 1;
  ctx->cam_base = *camera;
  ctx->have_scene = true;
  if (ctx->W > 0) make_dcamera(ctx->cam_base, ctx->W, ctx->H, &ctx->cam);
  return BDPT_OK;
}

int bdpt_set_environment_map(BdptContext* ctx, const float* rgb, int width, int height) {
  if (!ctx || ((width > 0 || height > 0) && (!rgb || width <= 0 || height < 2))) return BDPT_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  dfree(ctx->d_envmap); ctx->env_w = ctx->env_h = 0;
  if (width > 0) {
    std::vector<float4> h((size_t)width * height);
    for (size_t i = 0; i < h.size(); i++) h[i] = make_float4(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], 0.f);
    CK(cudaMalloc(&ctx->d_envmap, h.size() * sizeof(float4)));
    CK(cudaMemcpy(ctx->d_envmap, h.data(), h.size() * sizeof(float4), cudaMemcpyHostToDevice));
    ctx->env_w = width; ctx->env_h = height;
  }
  ctx->sc.envmap = ctx->d_envmap; ctx->sc.env_w = ctx->env_w; ctx->sc.env_h = ctx->env_h;
  return BDPT_OK;
}

int bdpt_set_camera(BdptContext* ctx, const BdptCamera* camera) {
  if (!ctx || !camera) return BDPT_ERR_ARG;
  ctx->cam_base = *camera;
  if (ctx->W > 0) make_dcamera(ctx->cam_base, ctx->W, ctx->H, &ctx->cam);
  return BDPT_OK;
}

int bdpt_set_frame(BdptContext* ctx, int width, int height) {
  if (!ctx || width <= 0 || height <= 0 || (long long)width * height > (1ll << 30)) return BDPT_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (width != ctx->W || height != ctx->H) {
    CK(cudaStreamSynchronize(ctx->stream));
    // no frame is held until every allocation below has succeeded
    ctx->W = ctx->H = 0;
    dfree(ctx->acc); ctx->acc_eye = ctx->acc_light = ctx->acc_comp = ctx->acc_batch = nullptr;
    dfree(ctx->resolve_tmp); dfree(ctx->rgba_tmp);
    const size_t npix = (size_t)width * height;
    cudaError_t e = cudaMalloc(&ctx->acc, 6 * npix * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->resolve_tmp, npix * 3 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->rgba_tmp, npix * 4);
    if (e != cudaSuccess) {
      dfree(ctx->acc); dfree(ctx->resolve_tmp); dfree(ctx->rgba_tmp);
      return fail(ctx, BDPT_ERR_CUDA, std::string("frame buffers: ") + cudaGetErrorString(e));
    }
    ctx->acc_eye = ctx->acc; ctx->acc_light = ctx->acc + npix;
    ctx->acc_comp = ctx->acc + 2 * npix; ctx->acc_batch = ctx->acc + 4 * npix;
    ctx->W = width; ctx->H = height;
    ctx->cell[0] = ctx->cell[1] = ctx->cell[2] = ctx->cell[3] = 0;
  }
  make_dcamera(ctx->cam_base, ctx->W, ctx->H, &ctx->cam);
  return bdpt_clear(ctx);
}

int bdpt_set_lens(BdptContext* ctx, double lens_radius, double focal_distance) {
  if (!ctx) return BDPT_ERR_ARG;
  if (!(lens_radius >= 0.0) || (lens_radius > 0.0 && !(focal_distance > 0.0))) return fail(ctx, BDPT_ERR_ARG, "bad lens parameters");
  ctx->lens_radius = (float)lens_radius; ctx->focal_dist = (float)(lens_radius > 0.0 ? focal_distance : 1.0);
  return BDPT_OK;
}

int bdpt_set_cell(BdptContext* ctx, int x, int y, int dx, int dy) {
  if (!ctx) return BDPT_ERR_ARG;
  if (dx == 0 && dy == 0) { ctx->cell[0] = ctx->cell[1] = ctx->cell[2] = ctx->cell[3] = 0; return BDPT_OK; }
  if (ctx->W <= 0) return fail(ctx, BDPT_ERR_ARG, "no frame size set (bdpt_set_frame)");
  if (x < 0 || y < 0 || dx <= 0 || dy <= 0 || x + dx > ctx->W || y + dy > ctx->H) return fail(ctx, BDPT_ERR_ARG, "cell outside the frame");
  ctx->cell[0] = x; ctx->cell[1] = y; ctx->cell[2] = dx; ctx->cell[3] = dy;
  return BDPT_OK;
}

int bdpt_clear(BdptContext* ctx) {
  if (!ctx) return BDPT_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (ctx->acc && ctx->W > 0)
    CK(cudaMemsetAsync(ctx->acc, 0, 6 * (size_t)ctx->W * ctx->H * sizeof(float4), ctx->stream));
  if (ctx->d_stats) CK(cudaMemsetAsync(ctx->d_stats, 0, 4 * sizeof(unsigned long long), ctx->stream));
  BdptStats z{};
  z.ms_bvh_build = ctx->ms_bvh_build; z.bvh_nodes = ctx->bvh_nodes;
  ctx->stats = z;
  return BDPT_OK;
}

int bdpt_render(BdptContext* ctx, const BdptRenderParams* p) {
  int rc = check_ready(ctx, true);
  if (rc) return rc;
  if (!p || p->ns_total <= 0 || p->max_ray_depth < 0 || p->n_samples < 0 || p->first_sample < 0 || p->max_ray_depth > 64)
    return fail(ctx, BDPT_ERR_ARG, "bad render parameters");
  CK(cudaSetDevice(ctx->device));
  ctx->ns_last = p->ns_total;
  if (p->n_samples == 0) return BDPT_OK;

  Job J;
  // batch size: bounded by the vertex planes (64 B/vertex) and the worst-case shadow queue (48 B/entry)
  const int m = std::max(1, p->max_ray_depth);
  const size_t per_sample = (size_t)2 * (m + 1) * 64 + (size_t)(m + 1) * (m + 1) * 48 + 8;
  size_t B = (size_t)((size_t)ctx->batch_gib << 30) / per_sample;
  B = std::min<size_t>(B, (size_t)1 << ctx->batch_log2);
  // a point light's pixel takes one splat per sample: its batch sum stays an exact multiple of the splat up to 2^24 (k_fold)
  if (ctx->point_lights) B = std::min<size_t>(B, (size_t)1 << 24);
  B = std::max<size_t>(B & ~(size_t)1023, 1024);
  const int cw = ctx->cell[2] > 0 ? ctx->cell[2] : ctx->W, ch = ctx->cell[2] > 0 ? ctx->cell[3] : ctx->H;
  const int tiles_x = (cw + 7) / 8, npix_pad = tiles_x * ((ch + 3) / 4) * 32;
  const unsigned long long total = (unsigned long long)npix_pad * (unsigned long long)p->n_samples;
  if (total < B) B = (size_t)((total + 1023) & ~1023ull);
  // If the device cannot give the slab (another context holds the memory), halve the batch until it can: the image
  // does not depend on the batch size, only the kernels' tails do.
  for (;;) {
    const size_t qcap = B * (size_t)(m + 1) * (size_t)(m + 1);
    rc = ensure_batch_storage(ctx, B, m, std::max<size_t>(qcap, 1024));
    if (rc == BDPT_OK) break;
    cudaGetLastError();   // clear the allocation error
    if (B <= 65536) return rc;
    B = std::max<size_t>((B / 2) & ~(size_t)1023, 1024);
  }
  fill_job(ctx, p, &J);
  J.B = (int)B; J.g_end = total;
  // the planes are indexed with the job's own B and m (VSLOT), capacity may be larger

  const unsigned long long n_batches = (total + B - 1) / B;
  const long timed = (long)std::min<unsigned long long>(n_batches, 64);
  cudaEvent_t e0 = get_event(ctx, 0), e1 = get_event(ctx, 1);
  CK(cudaEventRecord(e0, ctx->stream));
  CK(cudaMemsetAsync(ctx->qctr + 4, 0, sizeof(unsigned int), ctx->stream));   // queue overflow: sticky over the batches
  for (unsigned long long b = 0; b < n_batches; b++) {
    J.g0 = b * B;
    J.k0 = (unsigned int)(J.g0 / (unsigned long long)npix_pad); J.p0 = (unsigned int)(J.g0 % (unsigned long long)npix_pad);
    rc = run_batch<false>(ctx, J, b < (unsigned long long)timed ? 2 + 4 * (long)b : -1);
    if (rc) return rc;
    // the batch's splats join the frame's sums (compensated: kernels.cuh, k_fold)
    const size_t nacc = 2 * (size_t)ctx->W * ctx->H;
    k_fold<<<(unsigned)((nacc + 255) / 256), 256, 0, ctx->stream>>>(ctx->acc, ctx->acc_comp, ctx->acc_batch, nacc);
    ctx->stats.kernel_launches += 1;
  }
  CK(cudaMemcpyAsync(ctx->h_qctr, ctx->qctr, 8 * sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaEventRecord(e1, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ctx->stats.ms_total = ms;
  double w = 0, c = 0, s = 0;
  for (long b = 0; b < timed; b++) {
    float a;
    cudaEventElapsedTime(&a, ctx->ev[2 + 4 * b], ctx->ev[3 + 4 * b]); w += a;
    cudaEventElapsedTime(&a, ctx->ev[3 + 4 * b], ctx->ev[4 + 4 * b]); c += a;
    cudaEventElapsedTime(&a, ctx->ev[4 + 4 * b], ctx->ev[5 + 4 * b]); s += a;
  }
  ctx->stats.ms_walk = w; ctx->stats.ms_connect = c; ctx->stats.ms_shadow = s;
  ctx->stats.timed_batches = (uint32_t)timed; ctx->stats.n_batches = (uint32_t)n_batches;
  ctx->stats.batch_samples = (uint32_t)B;
  ctx->stats.samples += (unsigned long long)cw * ch * (unsigned long long)p->n_samples;
  if (ctx->h_qctr[4]) return fail(ctx, BDPT_ERR_OVERFLOW, "shadow queue overflow");
  return BDPT_OK;
}

int bdpt_framebuffer_device(BdptContext* ctx, int which, void** device_ptr, size_t* n_floats) {
  if (!ctx || !device_ptr || which < BDPT_BUFFER_SUM || which > BDPT_BUFFER_LIGHT) return BDPT_ERR_ARG;
  if (!ctx->acc) return fail(ctx, BDPT_ERR_ARG, "no frame size set");
  const size_t n = (size_t)ctx->W * ctx->H * 4;
  // BDPT_BUFFER_SUM names the whole allocation (eye image then light image): what a single collective reduces
  *device_ptr = which == BDPT_BUFFER_LIGHT ? ctx->acc_light : ctx->acc_eye;
  if (n_floats) *n_floats = which == BDPT_BUFFER_SUM ? 2 * n : n;
  return BDPT_OK;
}

int bdpt_read_framebuffer(BdptContext* ctx, int which, float* rgb_out) {
  if (!ctx || !rgb_out || which < 0 || which > 2) return BDPT_ERR_ARG;
  if (!ctx->acc) return fail(ctx, BDPT_ERR_ARG, "no frame size set");
  CK(cudaSetDevice(ctx->device));
  const size_t npix = (size_t)ctx->W * ctx->H;
  k_resolve<<<(unsigned)((npix + 255) / 256), 256, 0, ctx->stream>>>(ctx->acc_eye, ctx->acc_light, which, 1.0f / (float)ctx->ns_last, npix, ctx->resolve_tmp);
  ctx->stats.kernel_launches += 1;
  CK(cudaMemcpyAsync(rgb_out, ctx->resolve_tmp, npix * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return BDPT_OK;
}

int bdpt_tonemap_rgba8(BdptContext* ctx, int which, uint8_t* rgba_out) {
  if (!ctx || !rgba_out || which < 0 || which > 2) return BDPT_ERR_ARG;
  if (!ctx->acc) return fail(ctx, BDPT_ERR_ARG, "no frame size set");
  CK(cudaSetDevice(ctx->device));
  int n = ctx->W * ctx->H;
  k_tonemap<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->acc_eye, ctx->acc_light, which, 1.0f / (float)ctx->ns_last, ctx->W, ctx->H, ctx->rgba_tmp);
  ctx->stats.kernel_launches += 1;
  CK(cudaMemcpyAsync(rgba_out, ctx->rgba_tmp, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return BDPT_OK;
}

int bdpt_get_stats(BdptContext* ctx, BdptStats* out) {
  if (!ctx || !out) return BDPT_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (ctx->d_stats) {
    unsigned long long h[4];
    CK(cudaMemcpyAsync(h, ctx->d_stats, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stats.rays_closest = h[0]; ctx->stats.rays_shadow = h[1]; ctx->stats.connections = h[2];
  }
  ctx->stats.ms_bvh_build = ctx->ms_bvh_build; ctx->stats.bvh_nodes = ctx->bvh_nodes;
  *out = ctx->stats;
  return BDPT_OK;
}

// ------------------------------------------------------------------------------------ test hooks
int bdpt_host_build_bvh(const BdptPrim* prims, uint32_t n_prims, uint32_t* nodes_out, uint32_t nodes_cap,
                        uint32_t* n_nodes, uint32_t* prim_order_out, uint32_t* max_depth) {
  if (!prims && n_prims) return BDPT_ERR_ARG;
  std::vector<BuildPrim> bp;
  make_build_prims(prims, n_prims, &bp);
  Bvh8 bvh;
  // BDPT_BVH_HOST_ALGO=ploc: the host restatement of the device build (bvh8.cpp, build_bvh8_ploc)
  const char* algo = getenv("BDPT_BVH_HOST_ALGO");
  if (algo && !strcmp(algo, "ploc")) build_bvh8_ploc(bp, &bvh); else build_bvh8(bp, &bvh);
  uint32_t nn = (uint32_t)(bvh.nodes.size() / 20);
  if (n_nodes) *n_nodes = nn;
  if (max_depth) *max_depth = bvh.max_depth;
  if (nodes_out) {
    if (nodes_cap < nn) return BDPT_ERR_ARG;
    memcpy(nodes_out, bvh.nodes.data(), bvh.nodes.size() * sizeof(uint32_t));
  }
  if (prim_order_out) memcpy(prim_order_out, bvh.prim_order.data(), bvh.prim_order.size() * sizeof(uint32_t));
  return BDPT_OK;
}

int bdpt_device_bvh(BdptContext* ctx, uint32_t* nodes_out, uint32_t nodes_cap, uint32_t* n_nodes, uint32_t* prim_order_out,
                    uint32_t* max_depth, uint32_t* rounds) {
  int rc = check_ready(ctx, false);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  if (n_nodes) *n_nodes = ctx->bvh_nodes;
  if (max_depth) *max_depth = ctx->bvh_depth;
  if (rounds) *rounds = ctx->d_prim_order ? ctx->bvh_rounds : 0;
  if (nodes_out) {
    if (nodes_cap < ctx->bvh_nodes) return fail(ctx, BDPT_ERR_ARG, "nodes_cap too small");
    CK(cudaMemcpyAsync(nodes_out, ctx->d_nodes, (size_t)ctx->bvh_nodes * 80, cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (prim_order_out) {
    if (ctx->d_prim_order) CK(cudaMemcpyAsync(prim_order_out, ctx->d_prim_order, (size_t)ctx->sc.n_prims * 4, cudaMemcpyDeviceToHost, ctx->stream));
    else memcpy(prim_order_out, ctx->h_prim_order.data(), ctx->h_prim_order.size() * 4);   // built on the host (small scene, or BDPT_BVH=host)
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return BDPT_OK;
}

int bdpt_intersect_batch(BdptContext* ctx, const double* rays, size_t n, double* hits) {
  int rc = check_ready(ctx, false);
  if (rc) return rc;
  if ((!rays || !hits) && n) return BDPT_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  DScene sc = ctx->sc; cudaStream_t st = ctx->stream;
  rc = run_hook(ctx, rays, n, 8, hits, 6, [&](float* din, float* dout) {
    k_intersect_batch<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(sc, din, n, dout);
  });
  if (rc) return rc;
  for (size_t i = 0; i < n; i++) {  // ids travelled as int bits in a float lane
    float a = (float)hits[6 * i], b = (float)hits[6 * i + 5];
    int ia, ib; memcpy(&ia, &a, 4); memcpy(&ib, &b, 4);
    if (hits[6 * i] == -1.0) { hits[6 * i + 1] = INFINITY; continue; }
    hits[6 * i] = (double)ia; hits[6 * i + 5] = (double)ib;
  }
  return BDPT_OK;
}

int bdpt_occluded_batch(BdptContext* ctx, const double* rays, size_t n, int32_t* occluded) {
  int rc = check_ready(ctx, false);
  if (rc) return rc;
  if ((!rays || !occluded) && n) return BDPT_ERR_ARG;
  if (n == 0) return BDPT_OK;
  CK(cudaSetDevice(ctx->device));
  std::vector<float> h(n * 8);
  for (size_t i = 0; i < n * 8; i++) h[i] = (float)rays[i];
  float* din = nullptr; int* dout = nullptr;
  CK(cudaMalloc(&din, h.size() * sizeof(float)));
  CK(cudaMalloc(&dout, n * sizeof(int)));
  CK(cudaMemcpyAsync(din, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  k_occluded_batch<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->sc, din, n, dout);
  ctx->stats.kernel_launches += 1;
  cudaError_t e = cudaMemcpyAsync(occluded, dout, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(din); cudaFree(dout);
  if (e != cudaSuccess) return fail(ctx, BDPT_ERR_CUDA, cudaGetErrorString(e));
  return BDPT_OK;
}

int bdpt_bsdf_eval_batch(BdptContext* ctx, const double* in, size_t n, double* out) {
  int rc = check_ready(ctx, false);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  DScene sc = ctx->sc; cudaStream_t st = ctx->stream;
  return run_hook(ctx, in, n, 7, out, 8, [&](float* din, float* dout) {
    k_bsdf_eval<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(sc, din, n, dout);
  });
}

int bdpt_bsdf_sample_batch(BdptContext* ctx, const double* in, size_t n, double* out) {
  int rc = check_ready(ctx, false);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  DScene sc = ctx->sc; cudaStream_t st = ctx->stream;
  return run_hook(ctx, in, n, 7, out, 7, [&](float* din, float* dout) {
    k_bsdf_sample<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(sc, din, n, dout);
  });
}

int bdpt_light_eval_batch(BdptContext* ctx, const double* in, size_t n, double* out) {
  int rc = check_ready(ctx, false);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  DScene sc = ctx->sc; cudaStream_t st = ctx->stream;
  return run_hook(ctx, in, n, 7, out, 6, [&](float* din, float* dout) {
    k_light_eval<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(sc, din, n, dout);
  });
}

int bdpt_camera_eval_batch(BdptContext* ctx, const double* in, size_t n, double* out) {
  int rc = check_ready(ctx, true);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  DCamera cam = ctx->cam; cudaStream_t st = ctx->stream;
  return run_hook(ctx, in, n, 5, out, 11, [&](float* din, float* dout) {
    k_camera_eval<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(cam, din, n, dout);
  });
}

int bdpt_trace_sample(BdptContext* ctx, const BdptRenderParams* p, int x, int y, int sample, double* out, size_t cap, size_t* n_out) {
  int rc = check_ready(ctx, true);
  if (rc) return rc;
  if (!p || !out || x < 0 || y < 0 || x >= ctx->W || y >= ctx->H) return fail(ctx, BDPT_ERR_ARG, "bad trace arguments");
  CK(cudaSetDevice(ctx->device));
  const int m = std::max(1, p->max_ray_depth);
  const size_t B = 1024;
  rc = ensure_batch_storage(ctx, B, m, B * (size_t)(m + 1) * (size_t)(m + 1));
  if (rc) return rc;
  // the trace must not disturb the frame: render into scratch accumulators
  float4 *se = nullptr, *sl = nullptr; double* dbg = nullptr;
  size_t nacc = (size_t)ctx->W * ctx->H * sizeof(float4);
  size_t ndbg = (size_t)(m + 1) * (m + 2) * 8;
  CK(cudaMalloc(&se, nacc)); CK(cudaMalloc(&sl, nacc)); CK(cudaMalloc(&dbg, ndbg * sizeof(double)));
  CK(cudaMemsetAsync(se, 0, nacc, ctx->stream)); CK(cudaMemsetAsync(sl, 0, nacc, ctx->stream));
  std::vector<double> hdbg(ndbg, -1.0);
  CK(cudaMemcpyAsync(dbg, hdbg.data(), ndbg * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  BdptRenderParams q = *p;
  q.first_sample = sample; q.n_samples = 1;
  Job J;
  fill_job(ctx, &q, &J);
  J.acc_eye = se; J.acc_light = sl; J.dbg = dbg;
  J.B = (int)B; J.g0 = 0; J.g_end = 1; J.single_x = x; J.single_y = y;
  rc = run_batch<true>(ctx, J, -1);
  std::vector<float4> hv; std::vector<int> hc(2 * B);
  size_t nv = 2 * (size_t)(m + 1) * B;
  if (rc == BDPT_OK) {
    hv.resize(4 * nv);   // 64-byte records: four float4 per vertex
    cudaMemcpyAsync(hv.data(), ctx->v0, 4 * nv * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(hc.data(), ctx->counts, 2 * B * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(hdbg.data(), dbg, ndbg * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  }
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(se); cudaFree(sl); cudaFree(dbg);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(ctx, BDPT_ERR_CUDA, cudaGetErrorString(e));

  std::vector<double> rec;
  const int nE = hc[0] & 0xff, nL = hc[B] & 0xff;   // (bit 8: the subpath ends on an emitter)
  rec.push_back(nE + 1); rec.push_back(nL + 1);
  const double nan = NAN;
  for (int kind = 0; kind < 2; kind++) {
    int nk = kind == 0 ? nE : nL;
    for (int j = 0; j <= nk; j++) {
      double r[13];
      for (double& v : r) v = nan;
      if (j == 0) { r[6] = r[7] = r[8] = 1.0; r[9] = 1.0; r[10] = -1; }
      else if (j == 1 && kind == 0) {
        r[0] = ctx->cam_base.pos[0]; r[1] = ctx->cam_base.pos[1]; r[2] = ctx->cam_base.pos[2];
        r[6] = r[7] = r[8] = 1.0; r[9] = 1.0; r[10] = -1; r[12] = 1.0;
      } else {
        size_t s = (size_t)(kind * (m + 1) + (j - 1)) * B;
        float4 a = hv[4 * s], b = hv[4 * s + 1], c = hv[4 * s + 2], d = hv[4 * s + 3];
        int mat; memcpy(&mat, &a.w, 4);
        r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = b.x; r[4] = b.y; r[5] = b.z;
        r[6] = c.x; r[7] = c.y; r[8] = c.z; r[10] = mat;
        if (j == 1) { r[9] = b.w; r[12] = d.x; }
      }
      rec.insert(rec.end(), r, r + 13);
    }
  }
  for (int t = 1; t <= nE; t++)
    for (int s = 0; s <= nL; s++) {
      const double* r = &hdbg[8 * ((size_t)(t - 1) * (m + 2) + s)];
      if (r[0] < 0) continue;
      rec.insert(rec.end(), r, r + 8);
    }
  if (n_out) *n_out = rec.size();
  memcpy(out, rec.data(), std::min(cap, rec.size()) * sizeof(double));
  return BDPT_OK;
}

}  // extern "C"

#include "comm.inc"
