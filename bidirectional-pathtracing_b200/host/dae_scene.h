// dae_scene.h — host-side scene ingest: COLLADA (.dae + CGL profile) -> flat POD scene.
#pragma once
#include <string>
#include <vector>

#include "../../include/bdpt_flatscene.h"

namespace bdpt_host {

struct FlatSceneData {
  BdptCamera camera;
  std::vector<BdptMaterial> materials;
  std::vector<BdptLight> lights;
  std::vector<BdptPrim> prims;
};

// Loads a scene the way the reference's `pathtracer` does before `-r` is applied (camera at 800x600).
bool load_dae(const std::string& path, FlatSceneData* out, std::string* err);
// Camera::set_screen_size (reference camera.cpp:83-89): what `-r W H` does to the camera.
void apply_screen_size(BdptCamera* cam, int w, int h);
// Camera::load_settings (reference camera.cpp:169-186), the `-c <file>` option: the 30 numbers Camera::dump_settings
// writes (fovs, aspect, clips, position, target, spherical pose, c2w row by row, screen size / distance, lens).  The
// reference applies it AFTER `-r` (main.cpp:174-178), so the file's fovs win over Camera::set_screen_size: the camera
// comes back with screen_w / screen_h set to the frame (w, h) and the file's fovs.
bool apply_camera_settings(BdptCamera* cam, const std::string& path, int w, int h, std::string* err);
// The ".fscene" byte layout of include/bdpt_flatscene.h.
std::vector<unsigned char> serialize(const FlatSceneData& s);

}  // namespace bdpt_host
