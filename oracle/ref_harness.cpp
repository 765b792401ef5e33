// TEST INFRASTRUCTURE (oracle).  Not part of the product; nothing shipped links it.
//
// ref_harness — drives the UNMODIFIED reference objects (compiled in place from
// /root/reference by oracle/Makefile) so that the CUDA path and the fp64 port
// in oracle/bdpt_oracle.cpp can be pinned against the reference itself:
//
//   dump      scene.dae out.fscene
//       Loads the scene exactly like src/application/main.cpp:155-176
//       (ColladaParser::load -> Application::init/load -> set_up_pathtracer)
//       and writes the primitive list in BVHAccel order (bvh.cpp:17), BSDFs,
//       lights and the placed camera as a flat scene (include/bdpt_flatscene.h).
//   intersect scene.dae rays.bin hits.bin
//       rays: N x {o[3], d[3], tmin, tmax} fp64.  Calls BVHAccel::intersect
//       (bvh.h:104-107) per ray; hits: N x {prim_id(-1 = miss), t, n[3], material}
//       as 6 fp64 each.
//   bsdf      scene.dae in.bin out.bin
//       in: N x {material, wo[3], wi[3]} fp64 (7 each); out: N x {f[3], sample_pdf(0,wi),
//       is_delta, emission[3]} (8 each) via BSDF::f / sample_pdf (bsdf.h:59-127).
//   bsdf_sample scene.dae in.bin out.bin
//       in: N x {material, wo[3]} (4 each); out: N x {wi[3], pdf, f[3]} (7 each): BSDF::sample_f called
//       in order from this fresh process, so the draws are the reference's default-seeded streams
//       (random_util.h:10-22).  wi and pdf are preset to NaN: RefractionBSDF::sample_f returns
//       without writing them on total internal reflection (advanced_bsdf.cpp:166-168).
//   envdir    W H map.bin dirs.bin out.bin
//       map: W x H x 3 fp64 radiance (row 0 = up); dirs: N x 3; out: N x 3 = EnvironmentLight::sample_dir
//       (environment_light.cpp:157-165: dir -> theta, phi -> x, y -> bilerp) of a reference EnvironmentLight built
//       on that map.  (Its constructor writes probability_debug.png into the working directory.)
//   exr       file.exr out.bin
//       Decodes the file with the reference's own EXR reader (the vendored tinyexr, called the way load_exr does,
//       src/application/main.cpp:40-78: half channels requested as float) and writes {width, height, then R G B per
//       pixel, rows top first} as fp64 — the pin of host/envmap_read.h.
//   lensray   scene.dae W H lensRadius focalDistance in.bin out.bin
//       in: N x {x, y, rndR, rndTheta}; out: N x {o[3], d[3]} = Camera::generate_ray_for_thin_lens
//       (camera_lens.cpp:22-43) of the scene's placed camera.
//   light     scene.dae in.bin out.bin
//       in: N x {light, p[3], wi[3]} (7 each); out: N x {contain_point, Le[3] (sample_pdf), point_pdf, dir_pdf}
//   camera    scene.dae W H in.bin out.bin
//       in: N x {x, y (generate_ray args), p[3]} (5 each); out: N x {ray o[3], d[3],
//       We, dir_pdf, raster x, raster y, dist} (11 each) via Camera::generate_ray
//       / sample_ray_pdf (camera.cpp:191-248).
//   render    scene.dae W H ns m threads out_prefix [t s]
//       Runs RaytracedRenderer::start_raytracing (tile queue + T std::threads,
//       raytraced_renderer.cpp:267-328) and dumps sampleBuffer / eyeBuffer /
//       lightBuffer as raw fp64 RGB (row 0 = bottom) to out_prefix.{sum,eye,light}.f64.
//       With "t s" only strategy (i_eye == t, i_light == s) is accumulated — the
//       author's own debug filter at bidirection.cpp:492-493, applied by
//       overriding the virtual est_radiance_global_illumination with the loop of
//       bidirection.cpp:472-500.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#define TINYEXR_IMPLEMENTATION   /* the reference compiles tinyexr into its main.cpp (:4-5); the harness has its own main */
#include "CGL/tinyexr.h"

// The reference keeps the members we must read (BSDF parameters, camera pose,
// renderer internals) private; the harness opens them for inspection only.
#define private public
#define protected public
#include "application/application.h"
#include "pathtracer/bidirection.h"
#include "pathtracer/raytraced_renderer.h"
#include "scene/collada/collada.h"
#include "scene/environment_light.h"
#include "scene/light.h"
#include "scene/sphere.h"
#include "scene/triangle.h"
#undef private
#undef protected

#include "../include/bdpt_flatscene.h"

using namespace CGL;
using namespace CGL::SceneObjects;

namespace {

struct Loaded {
  Application* app = nullptr;
  RaytracedRenderer* rr = nullptr;
};

Loaded load_scene(const char* path, size_t W, size_t H, int ns, int m, int threads) {
  AppConfig config;
  config.pathtracer_ns_aa = ns;
  config.pathtracer_max_ray_depth = m;
  config.pathtracer_num_threads = threads;
  Collada::SceneInfo* info = new Collada::SceneInfo();
  if (Collada::ColladaParser::load(path, info) < 0) {
    fprintf(stderr, "cannot load %s\n", path);
    exit(2);
  }
  Loaded L;
  L.app = new Application(config, false);
  L.app->init();
  L.app->load(info);
  if (W && H) L.app->resize(W, H);
  L.app->set_up_pathtracer();
  L.rr = static_cast<RaytracedRenderer*>(L.app->renderer);
  return L;
}

void put3(double* dst, const Vector3D& v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }

int material_index(std::vector<BSDF*>& table, BSDF* b) {
  for (size_t i = 0; i < table.size(); i++) if (table[i] == b) return (int)i;
  table.push_back(b);
  return (int)table.size() - 1;
}

void collect_materials(const std::vector<Primitive*>& prims, std::vector<BSDF*>& table) {
  for (Primitive* p : prims) material_index(table, p->get_bsdf());
}

std::vector<double> read_doubles(const char* path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
  f.seekg(0, std::ios::end);
  size_t n = (size_t)f.tellg() / sizeof(double);
  f.seekg(0);
  std::vector<double> v(n);
  f.read((char*)v.data(), n * sizeof(double));
  return v;
}

void write_doubles(const char* path, const std::vector<double>& v) {
  std::ofstream f(path, std::ios::binary);
  f.write((const char*)v.data(), v.size() * sizeof(double));
}

int cmd_dump(int argc, char** argv) {
  Loaded L = load_scene(argv[2], 0, 0, 1, 1, 1);
  const std::vector<Primitive*>& prims = L.rr->bvh->primitives;
  std::vector<BSDF*> mats;
  collect_materials(prims, mats);

  BdptSceneHeader hdr;
  memset(&hdr, 0, sizeof hdr);
  hdr.magic = BDPT_FSCENE_MAGIC;
  hdr.n_prims = (uint32_t)prims.size();
  hdr.n_materials = (uint32_t)mats.size();
  hdr.n_lights = (uint32_t)L.rr->scene->lights.size();

  BdptCamera cam;
  memset(&cam, 0, sizeof cam);
  Camera* c = L.rr->camera;
  put3(cam.pos, c->pos);
  put3(cam.c2w + 0, c->c2w[0]);
  put3(cam.c2w + 3, c->c2w[1]);
  put3(cam.c2w + 6, c->c2w[2]);
  cam.hfov_deg = c->hFov; cam.vfov_deg = c->vFov;
  cam.nclip = c->nClip; cam.fclip = c->fClip;
  cam.screen_dist = c->screenDist;
  cam.screen_w = (double)c->screenW; cam.screen_h = (double)c->screenH;
  put3(cam.target, c->targetPos);

  std::vector<BdptMaterial> fm(mats.size());
  for (size_t i = 0; i < mats.size(); i++) {
    BdptMaterial& o = fm[i];
    memset(&o, 0, sizeof o);
    BSDF* b = mats[i];
    if (auto* d = dynamic_cast<DiffuseBSDF*>(b)) { o.type = BDPT_MAT_DIFFUSE; put3(o.reflectance, d->reflectance); }
    else if (auto* e = dynamic_cast<EmissionBSDF*>(b)) { o.type = BDPT_MAT_EMISSION; put3(o.radiance, e->radiance); }
    else if (auto* mi = dynamic_cast<MirrorBSDF*>(b)) { o.type = BDPT_MAT_MIRROR; put3(o.reflectance, mi->reflectance); /* MirrorBSDF::roughness is never initialised (bsdf.h:214): not recorded */ }
    else if (auto* g = dynamic_cast<GlassBSDF*>(b)) { o.type = BDPT_MAT_GLASS; put3(o.reflectance, g->reflectance); put3(o.transmittance, g->transmittance); o.ior = g->ior; o.roughness = g->roughness; }
    else if (auto* r = dynamic_cast<RefractionBSDF*>(b)) { o.type = BDPT_MAT_REFRACTION; put3(o.transmittance, r->transmittance); o.ior = r->ior; o.roughness = r->roughness; }
    else if (auto* mf = dynamic_cast<MicrofacetBSDF*>(b)) { o.type = BDPT_MAT_MICROFACET; put3(o.eta, mf->eta); put3(o.k, mf->k); o.alpha = mf->alpha; }
    else { fprintf(stderr, "unknown BSDF class at material %zu\n", i); return 3; }
  }

  std::vector<BdptLight> fl(hdr.n_lights);
  for (size_t i = 0; i < fl.size(); i++) {
    BdptLight& o = fl[i];
    memset(&o, 0, sizeof o);
    SceneLight* l = L.rr->scene->lights[i];
    if (auto* a = dynamic_cast<AreaLight*>(l)) {
      o.type = BDPT_LIGHT_AREA; put3(o.radiance, a->radiance); put3(o.position, a->position);
      put3(o.direction, a->direction); put3(o.dim_x, a->dim_x); put3(o.dim_y, a->dim_y);
    } else if (auto* p = dynamic_cast<PointLight*>(l)) {
      o.type = BDPT_LIGHT_POINT; put3(o.radiance, p->radiance); put3(o.position, p->position);
    } else { fprintf(stderr, "light %zu is of a class BDPT aborts on (light.cpp)\n", i); o.type = -1; }
  }

  std::vector<BdptPrim> fp(prims.size());
  for (size_t i = 0; i < prims.size(); i++) {
    BdptPrim& o = fp[i];
    memset(&o, 0, sizeof o);
    o.material = material_index(mats, prims[i]->get_bsdf());
    if (auto* t = dynamic_cast<Triangle*>(prims[i])) {
      o.type = BDPT_PRIM_TRIANGLE;
      put3(o.data + 0, t->p1); put3(o.data + 3, t->p2); put3(o.data + 6, t->p3);
      put3(o.data + 9, t->n1); put3(o.data + 12, t->n2); put3(o.data + 15, t->n3);
    } else if (auto* s = dynamic_cast<Sphere*>(prims[i])) {
      o.type = BDPT_PRIM_SPHERE;
      put3(o.data, s->o); o.data[3] = s->r;
    } else { fprintf(stderr, "unknown primitive class\n"); return 3; }
  }

  std::ofstream f(argv[3], std::ios::binary);
  f.write((const char*)&hdr, sizeof hdr);
  f.write((const char*)&cam, sizeof cam);
  f.write((const char*)fm.data(), fm.size() * sizeof(BdptMaterial));
  f.write((const char*)fl.data(), fl.size() * sizeof(BdptLight));
  f.write((const char*)fp.data(), fp.size() * sizeof(BdptPrim));
  fprintf(stderr, "dumped %zu prims, %zu materials, %zu lights\n", fp.size(), fm.size(), fl.size());
  return 0;
}

int cmd_intersect(int argc, char** argv) {
  Loaded L = load_scene(argv[2], 0, 0, 1, 1, 1);
  const std::vector<Primitive*>& prims = L.rr->bvh->primitives;
  std::unordered_map<const Primitive*, long> index;
  for (size_t i = 0; i < prims.size(); i++) index[prims[i]] = (long)i;
  std::vector<BSDF*> mats;
  collect_materials(prims, mats);
  std::vector<double> rays = read_doubles(argv[3]);
  size_t n = rays.size() / 8;
  std::vector<double> out(n * 6);
  for (size_t i = 0; i < n; i++) {
    const double* r = &rays[i * 8];
    Ray ray(Vector3D(r[0], r[1], r[2]), Vector3D(r[3], r[4], r[5]));
    ray.min_t = r[6]; ray.max_t = r[7];
    Intersection isect;
    bool hit = L.rr->bvh->intersect(ray, &isect);
    double* o = &out[i * 6];
    if (hit) {
      o[0] = (double)index[isect.primitive]; o[1] = isect.t; put3(o + 2, isect.n);
      o[5] = (double)material_index(mats, isect.bsdf);
    } else { o[0] = -1; o[1] = INF_D; o[2] = o[3] = o[4] = 0; o[5] = -1; }
  }
  write_doubles(argv[4], out);
  return 0;
}

int cmd_bsdf(int argc, char** argv) {
  Loaded L = load_scene(argv[2], 0, 0, 1, 1, 1);
  std::vector<BSDF*> mats;
  collect_materials(L.rr->bvh->primitives, mats);
  std::vector<double> in = read_doubles(argv[3]);
  size_t n = in.size() / 7;
  std::vector<double> out(n * 8);
  for (size_t i = 0; i < n; i++) {
    const double* q = &in[i * 7];
    BSDF* b = mats[(size_t)q[0]];
    Vector3D wo(q[1], q[2], q[3]), wi(q[4], q[5], q[6]);
    double* o = &out[i * 8];
    put3(o, b->f(wo, wi));
    // MicrofacetBSDF::sample_pdf aborts (advanced_bsdf.cpp:144-148); report NaN instead of dying
    o[3] = dynamic_cast<MicrofacetBSDF*>(b) ? NAN : b->sample_pdf(Vector3D(), wi);
    o[4] = b->is_delta() ? 1.0 : 0.0;
    put3(o + 5, b->get_emission());
  }
  write_doubles(argv[4], out);
  return 0;
}

int cmd_bsdf_sample(int argc, char** argv) {
  Loaded L = load_scene(argv[2], 0, 0, 1, 1, 1);
  std::vector<BSDF*> mats;
  collect_materials(L.rr->bvh->primitives, mats);
  std::vector<double> in = read_doubles(argv[3]);
  size_t n = in.size() / 4;
  std::vector<double> out(n * 7);
  for (size_t i = 0; i < n; i++) {
    const double* q = &in[i * 4];
    BSDF* b = mats[(size_t)q[0]];
    Vector3D wi(NAN, NAN, NAN);
    double pdf = NAN;
    Vector3D f = b->sample_f(Vector3D(q[1], q[2], q[3]), &wi, &pdf);
    double* o = &out[i * 7];
    put3(o, wi); o[3] = pdf; put3(o + 4, f);
  }
  write_doubles(argv[4], out);
  return 0;
}

int cmd_envdir(int argc, char** argv) {
  size_t W = atoi(argv[2]), H = atoi(argv[3]);
  std::vector<double> m = read_doubles(argv[4]);
  HDRImageBuffer* img = new HDRImageBuffer(W, H);
  for (size_t i = 0; i < W * H; i++) img->data[i] = Vector3D(m[3 * i], m[3 * i + 1], m[3 * i + 2]);
  EnvironmentLight* env = new EnvironmentLight(img);
  std::vector<double> in = read_doubles(argv[5]);
  size_t n = in.size() / 3;
  std::vector<double> out(n * 3);
  for (size_t i = 0; i < n; i++) {
    Ray r(Vector3D(0, 0, 0), Vector3D(in[3 * i], in[3 * i + 1], in[3 * i + 2]));
    put3(&out[3 * i], env->sample_dir(r));
  }
  write_doubles(argv[6], out);
  return 0;
}

int cmd_exr(int argc, char** argv) {
  const char* err = nullptr;
  EXRImage exr;
  InitEXRImage(&exr);
  if (ParseMultiChannelEXRHeaderFromFile(&exr, argv[2], &err) != 0) { fprintf(stderr, "tinyexr: %s\n", err ? err : "?"); return 2; }
  for (int i = 0; i < exr.num_channels; i++)
    if (exr.pixel_types[i] == TINYEXR_PIXELTYPE_HALF) exr.requested_pixel_types[i] = TINYEXR_PIXELTYPE_FLOAT;
  if (LoadMultiChannelEXRFromFile(&exr, argv[2], &err) != 0) { fprintf(stderr, "tinyexr: %s\n", err ? err : "?"); return 2; }
  int ci[3] = {-1, -1, -1};
  for (int i = 0; i < exr.num_channels; i++) {
    if (!strcmp(exr.channel_names[i], "R")) ci[0] = i;
    if (!strcmp(exr.channel_names[i], "G")) ci[1] = i;
    if (!strcmp(exr.channel_names[i], "B")) ci[2] = i;
  }
  if (ci[0] < 0 || ci[1] < 0 || ci[2] < 0) { fprintf(stderr, "no R G B channels\n"); return 2; }
  std::vector<double> out;
  out.push_back(exr.width); out.push_back(exr.height);
  for (size_t i = 0; i < (size_t)exr.width * exr.height; i++)
    for (int c = 0; c < 3; c++) out.push_back(((float*)exr.images[ci[c]])[i]);
  write_doubles(argv[3], out);
  return 0;
}

int cmd_lensray(int argc, char** argv) {
  size_t W = atoi(argv[3]), H = atoi(argv[4]);
  Loaded L = load_scene(argv[2], W, H, 1, 1, 1);
  Camera* c = L.rr->camera;
  c->lensRadius = atof(argv[5]); c->focalDistance = atof(argv[6]);
  std::vector<double> in = read_doubles(argv[7]);
  size_t n = in.size() / 4;
  std::vector<double> out(n * 6);
  for (size_t i = 0; i < n; i++) {
    Ray r = c->generate_ray_for_thin_lens(in[4 * i], in[4 * i + 1], in[4 * i + 2], in[4 * i + 3]);
    put3(&out[6 * i], r.o); put3(&out[6 * i + 3], r.d);
  }
  write_doubles(argv[8], out);
  return 0;
}

int cmd_light(int argc, char** argv) {
  Loaded L = load_scene(argv[2], 0, 0, 1, 1, 1);
  std::vector<double> in = read_doubles(argv[3]);
  size_t n = in.size() / 7;
  std::vector<double> out(n * 6);
  for (size_t i = 0; i < n; i++) {
    const double* q = &in[i * 7];
    SceneLight* l = L.rr->scene->lights[(size_t)q[0]];
    Vector3D p(q[1], q[2], q[3]), wi(q[4], q[5], q[6]);
    double pp = 0, dp = 0;
    double* o = &out[i * 6];
    o[0] = l->contain_point(p) ? 1.0 : 0.0;
    put3(o + 1, l->sample_pdf(p, wi, &pp, &dp));
    o[4] = pp; o[5] = dp;
  }
  write_doubles(argv[4], out);
  return 0;
}

int cmd_camera(int argc, char** argv) {
  size_t W = atoi(argv[3]), H = atoi(argv[4]);
  Loaded L = load_scene(argv[2], W, H, 1, 1, 1);
  Camera* c = L.rr->camera;
  std::vector<double> in = read_doubles(argv[5]);
  size_t n = in.size() / 5;
  std::vector<double> out(n * 11);
  for (size_t i = 0; i < n; i++) {
    const double* q = &in[i * 5];
    Ray r = c->generate_ray(q[0], q[1]);
    double* o = &out[i * 11];
    put3(o, r.o); put3(o + 3, r.d);
    Vector3D wi, ep, nrm; double dist, pp, dp; int x, y;
    Vector3D We = c->sample_ray_pdf(Vector3D(q[2], q[3], q[4]), &wi, &ep, &dist, &pp, &dp, &nrm, &x, &y, (int)W, (int)H);
    o[6] = We.x; o[7] = dp; o[8] = x; o[9] = y; o[10] = dist;
  }
  write_doubles(argv[6], out);
  return 0;
}

// The loop of bidirection.cpp:472-500 with the (i_eye, i_light) filter the
// author left commented out at :492-493.
class StrategyFilterBDPT : public BidirectionalPathTracer {
 public:
  int want_t = -1, want_s = -1;
  Vector3D est_radiance_global_illumination(const Ray& r) override {
    Vector3D L_out;
    std::vector<PathVertex> eye_path, light_path;
    prepare_bidirectional_subpath(r, 1., 1., eye_path, Vector3D(1., 1., 1.), r.d, false, nullptr);
    Ray light_init_ray;
    Vector3D light_init_radiance, light_init_normal;
    double light_point_pdf, light_dir_pdf;
    SceneLight* light = nullptr;
    light_init_ray = sample_light_ray(light_point_pdf, light_dir_pdf, light_init_radiance, light_init_normal, light);
    prepare_bidirectional_subpath(light_init_ray, light_point_pdf, light_dir_pdf, light_path,
                                  light_init_radiance, light_init_normal, true, light);
    for (int i = 1; i < (int)eye_path.size(); i++)
      for (int j = 0; j < (int)light_path.size(); j++) {
        if (i != want_t || j != want_s) continue;
        L_out += estimate_bidirection_radiance(i, j, eye_path, light_path);
      }
    return L_out;
  }
};

void dump_buffer(const std::string& path, const HDRImageBuffer& b) {
  std::vector<double> v(b.w * b.h * 3);
  for (size_t i = 0; i < b.w * b.h; i++) { v[3 * i] = b.data[i].x; v[3 * i + 1] = b.data[i].y; v[3 * i + 2] = b.data[i].z; }
  write_doubles(path.c_str(), v);
}

int cmd_render(int argc, char** argv) {
  size_t W = atoi(argv[3]), H = atoi(argv[4]);
  int ns = atoi(argv[5]), m = atoi(argv[6]), threads = atoi(argv[7]);
  std::string prefix = argv[8];
  Loaded L = load_scene(argv[2], W, H, ns, m, threads);
  RaytracedRenderer* rr = L.rr;
  if (argc >= 11) {
    StrategyFilterBDPT* f = new StrategyFilterBDPT();
    PathTracer* old = rr->pt;
    f->ns_aa = old->ns_aa; f->max_ray_depth = old->max_ray_depth;
    f->ns_area_light = old->ns_area_light; f->ns_diff = old->ns_diff; f->ns_glsy = old->ns_glsy;
    f->ns_refr = old->ns_refr; f->samplesPerBatch = old->samplesPerBatch; f->maxTolerance = old->maxTolerance;
    f->direct_hemisphere_sample = old->direct_hemisphere_sample; f->envLight = old->envLight;
    f->want_t = atoi(argv[9]); f->want_s = atoi(argv[10]);
    rr->pt = f;
    rr->set_frame_size(W, H);
  }
  auto t0 = std::chrono::steady_clock::now();
  {
    std::unique_lock<std::mutex> lk(rr->m_done);
    rr->start_raytracing();
    rr->cv_done.wait(lk, [rr] { return rr->state == RaytracedRenderer::DONE; });
  }
  double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  BidirectionalPathTracer* pt = static_cast<BidirectionalPathTracer*>(rr->pt);
  dump_buffer(prefix + ".sum.f64", pt->sampleBuffer);
  dump_buffer(prefix + ".eye.f64", pt->eyeBuffer);
  dump_buffer(prefix + ".light.f64", pt->lightBuffer);
  printf("HARNESS render wall_s=%.6f rays=%llu isects=%llu samples=%zu threads=%d\n", wall,
         rr->bvh->total_rays, rr->bvh->total_isects, W * H * (size_t)ns, threads);
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: ref_harness dump|intersect|bsdf|bsdf_sample|light|camera|render ...\n"); return 1; }
  std::string c = argv[1];
  if (c == "dump" && argc >= 4) return cmd_dump(argc, argv);
  if (c == "intersect" && argc >= 5) return cmd_intersect(argc, argv);
  if (c == "bsdf" && argc >= 5) return cmd_bsdf(argc, argv);
  if (c == "bsdf_sample" && argc >= 5) return cmd_bsdf_sample(argc, argv);
  if (c == "envdir" && argc >= 7) return cmd_envdir(argc, argv);
  if (c == "exr" && argc >= 4) return cmd_exr(argc, argv);
  if (c == "lensray" && argc >= 9) return cmd_lensray(argc, argv);
  if (c == "light" && argc >= 5) return cmd_light(argc, argv);
  if (c == "camera" && argc >= 7) return cmd_camera(argc, argv);
  if (c == "render" && argc >= 9) return cmd_render(argc, argv);
  fprintf(stderr, "bad arguments for %s\n", c.c_str());
  return 1;
}
