// host_api.cpp — C entry points of libbdpt_host.so (include/bdpt_host.h): the .dae loader as a library,
// so that the ingest path can be tested against the reference's own loader without a GPU.
#include "../../include/bdpt_host.h"

#include <cstdlib>
#include <cstring>
#include <string>

#include "dae_scene.h"
#include "envmap_read.h"

namespace { thread_local std::string g_err; }

extern "C" {

int bdpt_host_load_dae(const char* path, int width, int height, unsigned char** blob, size_t* size) {
  if (!path || !blob || !size) return -2;
  bdpt_host::FlatSceneData s;
  if (!bdpt_host::load_dae(path, &s, &g_err)) return -3;
  if (width > 0 && height > 0) bdpt_host::apply_screen_size(&s.camera, width, height);
  std::vector<unsigned char> b = bdpt_host::serialize(s);
  *blob = (unsigned char*)malloc(b.size());
  if (!*blob) { g_err = "out of memory"; return -2; }
  memcpy(*blob, b.data(), b.size());
  *size = b.size();
  return 0;
}

int bdpt_host_apply_camera_settings(void* camera, const char* path, int width, int height) {
  if (!camera || !path) return -2;
  return bdpt_host::apply_camera_settings((BdptCamera*)camera, path, width, height, &g_err) ? 0 : -3;
}

int bdpt_host_load_envmap(const char* spec, float** rgb, int* width, int* height) {
  if (!spec || !rgb || !width || !height) return -2;
  bdpt_host::EnvMap m;
  if (!bdpt_host::read_envmap(spec, &m, &g_err)) return -3;
  *rgb = (float*)malloc(m.rgb.size() * sizeof(float));
  if (!*rgb) { g_err = "out of memory"; return -2; }
  memcpy(*rgb, m.rgb.data(), m.rgb.size() * sizeof(float));
  *width = m.w; *height = m.h;
  return 0;
}

void bdpt_host_free(void* p) { free(p); }
const char* bdpt_host_last_error(void) { return g_err.c_str(); }

}  // extern "C"
