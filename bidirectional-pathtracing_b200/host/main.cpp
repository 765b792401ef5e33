// main.cpp — the `pathtracer` command line of the B200 build.
//
// Keeps the reference's windowless surface (reference src/application/main.cpp:79-182):
//     pathtracer [-s spp] [-m max_ray_depth] [-t threads] [-r W H] -f out.png scene.dae
// -s / -m / -r / -f mean what they mean there; -t is accepted and ignored (the GPU path has no host
// worker threads); -l -a -b -d -H are accepted and ignored exactly as the reference's BDPT ignores them;
// -c <file> loads a camera settings file (Camera::load_settings, camera.cpp:169-186; applied after -r like
// main.cpp:174-178); -p x y dx dy renders a cell and writes a dx x dy image (render_to_file's cell path,
// raytraced_renderer.cpp:338-346); -e <file> adds an environment light with a lat-long map (.exr / .pfm, or "r,g,b"
// for a constant one) to the scene's lights like main.cpp:115-118 + raytraced_renderer.cpp:117-119 do — the
// reference's BDPT then aborts (environment_light.cpp:182-208); what this build renders instead is defined in
// include/bdpt_flatscene.h (BDPT_LIGHT_ENV).
// Extra options of this build: -g N (use N GPUs of this box: samples sharded, scene ingested per GPU in parallel
// host threads, ONE ncclReduce of the accumulators to GPU 0), -S seed, -R (enable the roulette rule the reference ships
// commented out), -B (balance heuristic), -o scene.fscene (dump the flat scene and exit).
// Output: out.png (toColor: exposure sqrt(2), gamma 2.2) and out_rate.png like the reference.
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bdpt_b200.h"
#include "dae_scene.h"
#include "envmap_read.h"
#include "png_write.h"

static void usage(const char* bin) {
  printf("Usage: %s [options] <scenefile>\n", bin);
  printf("Program Options:\n");
  printf("  -s  <INT>        Number of camera rays per pixel\n");
  printf("  -l  <INT>        Number of samples per area light (ignored by BDPT)\n");
  printf("  -t  <INT>        Number of render threads (accepted, unused on the GPU)\n");
  printf("  -m  <INT>        Maximum ray depth\n");
  printf("  -e  <PATH>       Path to environment map (.exr, .pfm, or r,g,b for a constant environment)\n");
  printf("  -f  <FILENAME>   Image (.png) file to save output to\n");
  printf("  -r  <INT> <INT>  Width and height of output image\n");
  printf("  -p  <x> <y> <dx> <dy>  Used with -f and -r to render a cell\n");
  printf("  -c  <FILENAME>   Load camera settings file (mainly to set camera position when windowless)\n");
  printf("  -g  <INT>        Number of GPUs to shard the samples over (default 1)\n");
  printf("  -S  <INT>        Random seed (default 0)\n");
  printf("  -R               Russian roulette (off in the reference)\n");
  printf("  -b  <FLOAT>      The size of the aperture  / -d <FLOAT> the focal distance: like the reference's BDPT, ignored unless\n");
  printf("  -L               ... thin-lens camera for BDPT (extension)\n");
  printf("  -B               Balance heuristic for MIS (the reference uses the power heuristic, beta = 2)\n");
  printf("  -o  <FILENAME>   Write the flat scene (.fscene) and exit\n");
  printf("  -h               Print this help message\n\n");
}

#define MSG(...) do { fprintf(stderr, "[PathTracer] "); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } while (0)

int main(int argc, char** argv) {
  int ns = 1, m = 1, threads = 1, W = 0, H = 0, gpus = 1, roulette = 0, balance = 0;
  unsigned long long seed = 0;
  std::string out_png, out_fscene, cam_settings, env_spec;
  int cell_x = -1, cell_y = 0, cell_dx = 0, cell_dy = 0;
  double lens_radius = 0.0, focal_distance = 4.7;   // -b / -d (application.h:43-87 defaults: 0 / 4.7)
  int use_lens = 0;
  int opt;
  while ((opt = getopt(argc, argv, "s:l:t:m:e:hH:f:r:c:b:d:a:p:g:S:RBLo:")) != -1) {
    switch (opt) {
      case 'f': out_png = optarg; break;
      case 'r': W = atoi(argv[optind - 1]); H = optind < argc ? atoi(argv[optind]) : 0; optind++; break;
      case 's': ns = atoi(optarg); break;
      case 'm': m = atoi(optarg); break;
      case 't': threads = atoi(optarg); break;
      case 'g': gpus = atoi(optarg); break;
      case 'S': seed = strtoull(optarg, nullptr, 10); break;
      case 'R': roulette = 1; break;
      case 'B': balance = 1; break;
      case 'o': out_fscene = optarg; break;
      case 'c': cam_settings = optarg; break;
      case 'p':   // four arguments, parsed like main.cpp:97-103
        if (optind + 2 >= argc) { usage(argv[0]); return 1; }
        cell_x = atoi(argv[optind - 1]); cell_y = atoi(argv[optind]); cell_dx = atoi(argv[optind + 1]); cell_dy = atoi(argv[optind + 2]);
        optind += 3;
        break;
      case 'b': lens_radius = atof(optarg); break;          // kept, but like the reference's BDPT only honoured ...
      case 'd': focal_distance = atof(optarg); break;
      case 'L': use_lens = 1; break;                         // ... when -L asks for the thin-lens extension
      case 'l': break;                                       // no effect on the BDPT image
      case 'a': optind++; break;                              // two arguments, ignored
      case 'H': optind--; break;                              // flag in the reference (main.cpp:131-134)
      case 'e': env_spec = optarg; break;
      default: usage(argv[0]); return 1;
    }
  }
  (void)threads;
  if (optind >= argc) { usage(argv[0]); return 1; }
  std::string scene_path = argv[optind];
  MSG("Input scene file: %s", scene_path.c_str());

  bdpt_host::FlatSceneData scene;
  std::string err;
  if (!bdpt_host::load_dae(scene_path, &scene, &err)) { MSG("Error: %s", err.c_str()); return 1; }
  if (W <= 0 || H <= 0) { W = 800; H = 600; }
  bdpt_host::EnvMap envmap;
  if (!env_spec.empty()) {
    printf("[PathTracer] Loading environment map %s\n", env_spec.c_str());
    if (!bdpt_host::read_envmap(env_spec, &envmap, &err)) { MSG("Error: %s", err.c_str()); return 1; }
    BdptLight e; memset(&e, 0, sizeof e);
    e.type = BDPT_LIGHT_ENV; e.radiance[0] = e.radiance[1] = e.radiance[2] = 1.0;
    scene.lights.push_back(e);    // RaytracedRenderer::set_scene appends envLight to the scene's lights (:117-119)
  }
  if (!cam_settings.empty()) {
    bdpt_host::apply_screen_size(&scene.camera, W, H);    // -r first, then the file (main.cpp:174-178)
    if (!bdpt_host::apply_camera_settings(&scene.camera, cam_settings, W, H, &err)) { MSG("Error: %s", err.c_str()); return 1; }
    printf("[Camera] Loaded settings from %s\n", cam_settings.c_str());
  }
  const bool cell = cell_x >= 0;
  if (cell && (cell_dx <= 0 || cell_dy <= 0 || cell_y < 0 || cell_x + cell_dx > W || cell_y + cell_dy > H)) { MSG("Error: cell outside the frame"); return 1; }
  if (!out_fscene.empty()) {
    bdpt_host::apply_screen_size(&scene.camera, W, H);
    std::vector<unsigned char> b = bdpt_host::serialize(scene);
    FILE* f = fopen(out_fscene.c_str(), "wb");
    if (!f || fwrite(b.data(), 1, b.size(), f) != b.size()) { MSG("cannot write %s", out_fscene.c_str()); return 1; }
    fclose(f);
    MSG("Wrote flat scene (%zu primitives) to %s", scene.prims.size(), out_fscene.c_str());
    return 0;
  }
  if (out_png.empty()) { MSG("this build is windowless: give -f <file.png>"); return 1; }
  if (ns < 1) ns = 1;
  if (gpus < 1) gpus = 1;
  MSG("Rendering on %d GPU%s (-t %d accepted and unused)", gpus, gpus > 1 ? "s" : "", threads);

  // one context per GPU; the scene is replicated, every GPU ingests its copy in its own host thread
  std::vector<BdptContext*> ctx(gpus, nullptr);
  std::vector<int> rcs(gpus, 0);
  for (int g = 0; g < gpus; g++)
    if (bdpt_create(g, &ctx[g]) != BDPT_OK) { MSG("Error: %s", bdpt_last_error(nullptr)); return 1; }
  {
    std::vector<std::thread> th;
    for (int g = 0; g < gpus; g++)
      th.emplace_back([&, g]() {
        int rc = bdpt_upload_scene(ctx[g], &scene.camera, scene.materials.data(), (uint32_t)scene.materials.size(), scene.lights.data(),
                                   (uint32_t)scene.lights.size(), scene.prims.data(), (uint32_t)scene.prims.size());
        if (rc == BDPT_OK) rc = bdpt_set_frame(ctx[g], W, H);
        if (rc == BDPT_OK && use_lens) rc = bdpt_set_lens(ctx[g], lens_radius, focal_distance);
        if (rc == BDPT_OK && envmap.w > 0) rc = bdpt_set_environment_map(ctx[g], envmap.rgb.data(), envmap.w, envmap.h);
        if (rc == BDPT_OK && cell) rc = bdpt_set_cell(ctx[g], cell_x, cell_y, cell_dx, cell_dy);
        rcs[g] = rc;
      });
    for (auto& t : th) t.join();
  }
  for (int g = 0; g < gpus; g++) if (rcs[g] != BDPT_OK) { MSG("Error: %s", bdpt_last_error(ctx[g])); return 1; }
  if (gpus > 1 && bdpt_comm_init_all(ctx.data(), gpus) != BDPT_OK) { MSG("Error: %s", bdpt_last_error(ctx[0])); return 1; }
  BdptStats st0;
  bdpt_get_stats(ctx[0], &st0);
  fprintf(stdout, "[PathTracer] Building BVH from %zu primitives... Done! (%.4f sec)\n", scene.prims.size(), st0.ms_bvh_build / 1e3);
  fprintf(stdout, "[PathTracer] Rendering... "); fflush(stdout);

  auto t0 = std::chrono::steady_clock::now();
  {
    std::vector<std::thread> th;
    for (int g = 0; g < gpus; g++) {
      // contiguous blocks of sample indices, remainder to the low ranks (same rule as bdpt/shard.py)
      int base = ns / gpus, rem = ns % gpus;
      BdptRenderParams p; memset(&p, 0, sizeof p);
      p.ns_total = ns; p.max_ray_depth = m; p.n_samples = base + (g < rem ? 1 : 0); p.first_sample = g * base + std::min(g, rem);
      p.flags = (roulette ? BDPT_FLAG_ROULETTE : 0) | (balance ? BDPT_FLAG_BALANCE : 0); p.only_t = p.only_s = -1; p.seed = seed;
      th.emplace_back([&, g, p]() { rcs[g] = bdpt_render(ctx[g], &p); });
    }
    for (auto& t : th) t.join();
  }
  for (int g = 0; g < gpus; g++) if (rcs[g] != BDPT_OK) { MSG("Error: %s", bdpt_last_error(ctx[g])); return 1; }
  // several GPUs: ONE ncclReduce of the accumulators to GPU 0 over NVLink, then GPU 0 tone-maps like a single GPU would
  if (gpus > 1 && bdpt_reduce_all(ctx.data(), gpus, 0) != BDPT_OK) { MSG("Error: %s", bdpt_last_error(ctx[0])); return 1; }
  std::vector<uint8_t> rgba((size_t)W * H * 4);
  if (bdpt_tonemap_rgba8(ctx[0], BDPT_BUFFER_SUM, rgba.data()) != BDPT_OK) { MSG("Error: %s", bdpt_last_error(ctx[0])); return 1; }
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  unsigned long long rays = 0;
  for (int g = 0; g < gpus; g++) { BdptStats s; bdpt_get_stats(ctx[g], &s); rays += s.rays_closest + s.rays_shadow; }
  fprintf(stdout, "\r[PathTracer] Rendering... 100%%! (%.4fs)\n", secs);
  fprintf(stdout, "[PathTracer] BVH traced %llu rays.\n", rays);
  fprintf(stdout, "[PathTracer] Average speed %.4f million rays per second.\n", (double)rays / secs * 1e-6);
  fprintf(stdout, "[PathTracer] %.4f million BDPT samples per second.\n", (double)(cell ? cell_dx * cell_dy : W * H) * ns / secs * 1e-6);
  int out_w = W, out_h = H;
  if (cell) {
    // raytrace_cell copies the cell out of the frame (raytraced_renderer.cpp:643-647) and save_image flips it: rows
    // [H - y - dy, H - y) of the top-down image, columns [x, x + dx)
    std::vector<uint8_t> crop((size_t)cell_dx * cell_dy * 4);
    for (int r = 0; r < cell_dy; r++)
      memcpy(&crop[(size_t)r * cell_dx * 4], &rgba[((size_t)(H - cell_y - cell_dy + r) * W + cell_x) * 4], (size_t)cell_dx * 4);
    rgba.swap(crop); out_w = cell_dx; out_h = cell_dy;
  }
  fprintf(stderr, "[PathTracer] Saving to file: %s... ", out_png.c_str());
  if (!bdpt_host::write_png_rgba8(out_png, rgba.data(), out_w, out_h)) { fprintf(stderr, "failed\n"); return 1; }
  fprintf(stderr, "Done!\n");
  // save_sampling_rate_image (raytraced_renderer.cpp:730-761): every pixel received ns_aa samples -> rate 1 -> red
  if (out_png.size() > 4) {
    std::vector<uint8_t> rate((size_t)W * H * 4);
    for (size_t i = 0; i < (size_t)W * H; i++) { rate[4 * i] = 255; rate[4 * i + 1] = 0; rate[4 * i + 2] = 0; rate[4 * i + 3] = 255; }
    if (cell)   // pixels outside the cell were not sampled: rate 0 -> blue
      for (int yy = 0; yy < H; yy++)
        for (int xx = 0; xx < W; xx++) {
          const int fy = H - 1 - yy;   // frame row (0 = bottom) of image row yy
          if (xx >= cell_x && xx < cell_x + cell_dx && fy >= cell_y && fy < cell_y + cell_dy) continue;
          uint8_t* o = &rate[4 * ((size_t)yy * W + xx)]; o[0] = 0; o[1] = 0; o[2] = 255;
        }
    bdpt_host::write_png_rgba8(out_png.substr(0, out_png.size() - 4) + "_rate.png", rate.data(), W, H);   // always frame-sized (:737-760)
  }
  for (BdptContext* c : ctx) bdpt_destroy(c);
  fprintf(stdout, cell ? "[PathTracer] Cell job completed.\n" : "[PathTracer] Job completed.\n");
  return 0;
}
