/* bdpt_host.h — C ABI of libbdpt_host.so: the host-side scene ingest (no GPU needed).
 *
 * bdpt_host_load_dae stands where the reference's `ColladaParser::load` + `Application::load` +
 * `Scene::get_static_scene` + `RaytracedRenderer::build_accel`'s primitive collection do
 * (reference src/scene/collada/collada.cpp:133-227, src/application/application.cpp:228-304,
 * src/scene/gl_scene/scene.cpp:158-170, src/pathtracer/raytraced_renderer.cpp:355-360): it turns a
 * COLLADA file into the flat scene of bdpt_flatscene.h (".fscene" byte layout), ready for
 * bdpt_upload_scene.  width/height > 0 additionally apply `-r W H` (Camera::set_screen_size).
 */
#ifndef BDPT_HOST_H
#define BDPT_HOST_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
/* 0 = ok; the blob is malloc'ed and must be released with bdpt_host_free. <0 = error (bdpt_host_last_error). */
int bdpt_host_load_dae(const char* path, int width, int height, unsigned char** blob, size_t* size);
/* Camera::load_settings (reference src/pathtracer/camera.cpp:169-186), the `-c <file>` option, applied to a BdptCamera
 * (bdpt_flatscene.h) in place.  The reference loads the file AFTER `-r` (src/application/main.cpp:174-178): the file's
 * fovs replace the ones Camera::set_screen_size derived; width / height are the frame the camera will render. */
int bdpt_host_apply_camera_settings(void* camera, const char* path, int width, int height);
/* The `-e <path>` option (reference src/application/main.cpp:40-78,115-118: OpenEXR through tinyexr): a lat-long
 * environment map as width x height RGB floats, rows top first (malloc'ed, release with bdpt_host_free), ready for
 * bdpt_set_environment_map.  spec: an .exr file (single-part scan-line, HALF / FLOAT R G B, compression NONE / RLE /
 * ZIPS / ZIP), a .pfm file, or "r,g,b" for a constant environment. */
int bdpt_host_load_envmap(const char* spec, float** rgb, int* width, int* height);
void bdpt_host_free(void* p);
const char* bdpt_host_last_error(void);
#ifdef __cplusplus
}
#endif
#endif
