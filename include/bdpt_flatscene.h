/* bdpt_flatscene.h — the flat, POD scene exchanged across the C-ABI.
 *
 * One self-contained description of what the reference hands its integrator:
 *   - the primitive list in the order RaytracedRenderer::build_accel collects
 *     it (reference src/pathtracer/raytraced_renderer.cpp:355-360: objects in
 *     node order, triangles in face order) — primitive index == hit id;
 *   - the BSDF table (reference src/pathtracer/bsdf.h:132-300);
 *   - the light table (AreaLight / PointLight, src/scene/light.h:63-139);
 *   - the placed camera (src/pathtracer/camera.h:104-128) at its 800x600
 *     configuration; `screen_dist` lets a caller redo Camera::set_screen_size
 *     (camera.cpp:83-89) for any -r W H.
 *
 * All geometry is fp64 so that nothing is lost between the loader and either
 * consumer (the fp64 CPU oracle, the fp32 CUDA path which narrows on upload).
 *
 * On disk (".fscene"): BdptSceneHeader, BdptCamera, materials[], lights[],
 * prims[] back to back, little-endian, no padding beyond the structs' own.
 */
#ifndef BDPT_FLATSCENE_H
#define BDPT_FLATSCENE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BDPT_FSCENE_MAGIC 0x3130534654504442ull /* "BDPTFS01" little-endian */

enum BdptMaterialType {
  BDPT_MAT_DIFFUSE = 0,    /* DiffuseBSDF     bsdf.cpp:52-85        */
  BDPT_MAT_EMISSION = 1,   /* EmissionBSDF    bsdf.cpp:99-118       */
  BDPT_MAT_MIRROR = 2,     /* MirrorBSDF      advanced_bsdf.cpp:17-35   */
  BDPT_MAT_GLASS = 3,      /* GlassBSDF       advanced_bsdf.cpp:198-259 */
  BDPT_MAT_REFRACTION = 4, /* RefractionBSDF  advanced_bsdf.cpp:163-184 */
  BDPT_MAT_MICROFACET = 5  /* MicrofacetBSDF  advanced_bsdf.cpp:48-148  */
};

enum BdptLightType {
  BDPT_LIGHT_AREA = 0,  /* AreaLight  light.cpp:199-284 */
  BDPT_LIGHT_POINT = 1, /* PointLight light.cpp:102-153 */
  /* EnvironmentLight (src/scene/environment_light.cpp), which the reference's BDPT aborts on (:182-208).
   * Extension defined here: a far sphere that emits inward, radius 1000 x the scene's bounding radius about its
   * centre — i.e. an ordinary area light for every formula of the path (bdpt_oracle.cpp "environment").  `radiance`
   * is the constant radiance, or the scale of the lat-long map given with bdpt_set_environment_map. */
  BDPT_LIGHT_ENV = 2
};

enum BdptPrimType { BDPT_PRIM_TRIANGLE = 0, BDPT_PRIM_SPHERE = 1 };

typedef struct BdptSceneHeader {
  uint64_t magic;
  uint32_t n_prims;
  uint32_t n_materials;
  uint32_t n_lights;
  uint32_t reserved;
} BdptSceneHeader;

typedef struct BdptCamera {
  double pos[3];
  double c2w[9];      /* three column vectors: c2w[0..2]=screen X dir, [3..5]=Y, [6..8]=Z (points back toward the eye) */
  double hfov_deg;    /* at (screen_w, screen_h) below */
  double vfov_deg;
  double nclip;
  double fclip;
  double screen_dist; /* pixels; fixed at configure time (camera.cpp:46) */
  double screen_w;
  double screen_h;
  double target[3];   /* look-at point (bbox centroid), informational */
} BdptCamera;

typedef struct BdptMaterial {
  int32_t type;       /* BdptMaterialType */
  int32_t pad;
  double reflectance[3];   /* diffuse / mirror / glass */
  double transmittance[3]; /* glass / refraction */
  double radiance[3];      /* emission */
  double eta[3];           /* microfacet */
  double k[3];             /* microfacet */
  double alpha;            /* microfacet roughness */
  double ior;              /* glass / refraction */
  double roughness;        /* carried, unused by the reference */
} BdptMaterial;

typedef struct BdptLight {
  int32_t type;       /* BdptLightType */
  int32_t pad;
  double radiance[3];
  double position[3];
  double direction[3]; /* area: emitting side */
  double dim_x[3];     /* area: full edge vectors (area = |dim_x|*|dim_y|) */
  double dim_y[3];
} BdptLight;

typedef struct BdptPrim {
  int32_t type;       /* BdptPrimType */
  int32_t material;   /* index into materials[] */
  /* triangle: p1 p2 p3 n1 n2 n3 (18 doubles); sphere: centre[3], radius, rest 0 */
  double data[18];
} BdptPrim;

#ifdef __cplusplus
}
#endif
#endif /* BDPT_FLATSCENE_H */
