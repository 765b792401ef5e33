// shading.cuh — BSDF, light and camera device functions (fp32).
// Each function states the reference function whose behaviour it reproduces
// (paths relative to the reference tree).  Local directions are in the frame of
// make_frame() (z = shading normal).
#pragma once
#include "common.cuh"
#include "rng.cuh"
#include "../../include/bdpt_flatscene.h"

namespace bdpt {

struct Mat {
  int type;
  f3 reflectance, transmittance, radiance, eta, k;
  float alpha, ior;
};

template <int TR = 0>
__device__ inline int mat_type(const DScene& sc, int m) {
  if (TR & BDPT_TRAIT_LAMBERT) return BDPT_MAT_DIFFUSE;   // callers only ask whether the type is a delta type
  return __float_as_int(ldg4(sc.materials + 5 * m).w);
}

__device__ inline bool type_is_delta(int type) {  // BSDF::is_delta of each class, bsdf.h
  return type == BDPT_MAT_MIRROR || type == BDPT_MAT_GLASS || type == BDPT_MAT_REFRACTION;
}

template <int TR = 0>
__device__ inline Mat load_mat(const DScene& sc, int m) {
  Mat r;
  const float4* p = sc.materials + 5 * m;
  float4 a = ldg4(p), b = ldg4(p + 1), c = ldg4(p + 2);
  r.reflectance = mk3(a.x, a.y, a.z); r.type = __float_as_int(a.w);
  if (TR & BDPT_TRAIT_LAMBERT) {   // lets the compiler drop every other BSDF from the caller
    if (r.type != BDPT_MAT_EMISSION) r.type = BDPT_MAT_DIFFUSE;
    __builtin_assume(r.type == BDPT_MAT_DIFFUSE || r.type == BDPT_MAT_EMISSION);
  } else if (TR & BDPT_TRAIT_NOMF) {
    if (r.type == BDPT_MAT_MICROFACET) r.type = BDPT_MAT_DIFFUSE;
    __builtin_assume(r.type != BDPT_MAT_MICROFACET);
  }
  r.transmittance = mk3(b.x, b.y, b.z); r.ior = b.w;
  r.radiance = mk3(c.x, c.y, c.z); r.alpha = c.w;
  if (r.type == BDPT_MAT_MICROFACET) {
    float4 d = ldg4(p + 3), e = ldg4(p + 4);
    r.eta = mk3(d.x, d.y, d.z); r.k = mk3(e.x, e.y, e.z);
  } else { r.eta = mk3(0.f); r.k = mk3(0.f); }
  return r;
}

// sampler.cpp:91-94  CosineWeightedHemisphereSampler3D::pdf
__device__ inline float cosine_pdf(float z) { return z > 0.f ? z * BDPT_INV_PI : 0.f; }

// sin and cos of 2 pi u, u in [0, 1): the SFU forms evaluated at 2 pi u - pi, |argument| <= pi, where their
// absolute error is ~5e-7 (one fp32 ulp of a unit vector's component); sincospif is ~100 instructions
__device__ inline void sincos_2pi(float u, float* s, float* c) {
  const float a = (2.0f * u - 1.0f) * BDPT_PI;
  *s = -__sinf(a); *c = -__cosf(a);
}

// sampler.cpp:76-85  CosineWeightedHemisphereSampler3D::get_sample(pdf)
__device__ inline f3 cosine_hemisphere(float xi1, float xi2, float* pdf) {
  float r = sqrtf(xi1);
  float s, c;
  sincos_2pi(xi2, &s, &c);
  float z = sqrtf(1.0f - xi1);
  *pdf = z * BDPT_INV_PI;
  return mk3(r * c, r * s, z);
}

// sampler.cpp:17-24  UniformSphereSampler3D::get_sample
__device__ inline f3 uniform_sphere(float u1, float u2) {
  float z = u1 * 2.f - 1.f;
  float st = sqrtf(fmaxf(0.f, 1.f - z * z));
  float s, c;
  sincos_2pi(u2, &s, &c);
  return mk3(c * st, s * st, z);
}

// advanced_bsdf.cpp:280-303  BSDF::refract
__device__ inline bool refract_local(f3 wo, f3* wi, float ior) {
  bool enter = wo.z > 0.f;
  float eta = enter ? 1.0f / ior : ior;
  float z_sq = 1.f - eta * eta * (1.f - wo.z * wo.z);
  if (z_sq < 0.f) return false;
  *wi = mk3(-eta * wo.x, -eta * wo.y, enter ? -sqrtf(z_sq) : sqrtf(z_sq));
  return true;
}

__device__ inline float pow5(float x) { float x2 = x * x; return x2 * x2 * x; }

// Schlick term of GlassBSDF (advanced_bsdf.cpp:214-222 and :247-253)
__device__ inline float glass_R(float ior, f3 refr, bool wo_z_positive) {
  float c = fabsf(refr.z) / norm(refr);
  float eta = wo_z_positive ? 1.0f / ior : ior;
  float r0 = (1.f - eta) / (1.f + eta);
  r0 *= r0;
  return r0 + (1.f - r0) * pow5(1.f - c);
}

// Microfacet helpers: bsdf.h:176-184 (getTheta, Lambda), advanced_bsdf.cpp:48-82 (G, D, F)
__device__ inline float mf_lambda(float alpha, float wz) {
  float cz = fminf(fmaxf(wz, -1.0f + 1e-5f), 1.0f - 1e-5f);
  float tan_t = sqrtf(fmaxf(0.f, 1.f - cz * cz)) / cz;
  float a = 1.0f / (alpha * tan_t);
  return 0.5f * (erff(a) - 1.0f + expf(-a * a) / (a * BDPT_PI));
}
__device__ inline f3 mf_F(const Mat& m, float c) {  // c = |cos theta_i|
  float c2 = c * c;
  f3 e2k2 = m.eta * m.eta + m.k * m.k;
  f3 tec = 2.f * c * m.eta;
  f3 rs, rp;
  rs.x = (e2k2.x - tec.x + c2) / (e2k2.x + tec.x + c2);
  rs.y = (e2k2.y - tec.y + c2) / (e2k2.y + tec.y + c2);
  rs.z = (e2k2.z - tec.z + c2) / (e2k2.z + tec.z + c2);
  rp.x = (e2k2.x * c2 - tec.x + 1.f) / (e2k2.x * c2 + tec.x + 1.f);
  rp.y = (e2k2.y * c2 - tec.y + 1.f) / (e2k2.y * c2 + tec.y + 1.f);
  rp.z = (e2k2.z * c2 - tec.z + 1.f) / (e2k2.z * c2 + tec.z + 1.f);
  return (rs + rp) * 0.5f;
}

// The BSDF evaluations below only ever need the z components of the two local directions and, for the microfacet
// lobe, the squared length of the tangential part of wo + wi (every lobe here is isotropic about the shading normal),
// so they are written in that form: k_connect then needs no tangent frame at all — woz = dot(n, wo_world),
// wiz = dot(n, wi_world), hxy2 = |(wo_world + wi_world) x n|^2 — and the vector forms used by the walk and the test
// hooks forward to them.  Directions are unit length.  hxy2 is the half vector's sin^2 before normalisation: taking
// it from a cross product (or from local x, y) instead of 1 - cos^2 keeps tan^2(theta_h) accurate to fp32 rounding for
// the alpha = 0.05 lobes of the reference's metals, where exp(-tan^2 / alpha^2) amplifies an absolute error of 1e-7
// in cos^2 to 1e-4 relative in f.  Only MicrofacetBSDF reads it.
__device__ inline float half_xy2(f3 wo, f3 wi) { float x = wo.x + wi.x, y = wo.y + wi.y; return x * x + y * y; }
__device__ inline float half_xy2_world(f3 wo, f3 wi, f3 n) { return norm2(cross(wo + wi, n)); }

// Schlick term from the cosine of the refracted direction
__device__ inline float glass_R_cos(float ior, float c, bool wo_z_positive) {
  float eta = wo_z_positive ? 1.0f / ior : ior;
  float r0 = (1.f - eta) / (1.f + eta);
  r0 *= r0;
  return r0 + (1.f - r0) * pow5(1.f - c);
}

// Beckmann D of the half vector given s_z = woz + wiz and hxy2 (advanced_bsdf.cpp:52-63): tan^2 = hxy2 / s_z^2,
// cos^2 = s_z^2 / (hxy2 + s_z^2)
__device__ inline float mf_D_h(float alpha, float sz, float hxy2) {
  const float sz2 = sz * sz, ia2 = 1.0f / (alpha * alpha);
  const float c2 = sz2 / (hxy2 + sz2);
  return expf(-hxy2 / sz2 * ia2) * ia2 / (BDPT_PI * c2 * c2);
}

// BSDF::f(wo, wi): bsdf.cpp:52-60 diffuse; :99-101 emission; advanced_bsdf.cpp:17-19,
// :159-161,:194-196 mirror/refraction/glass (all zero); :84-95 microfacet.
__device__ inline f3 bsdf_f_z(const Mat& m, float woz, float wiz, float hxy2) {
  if (m.type == BDPT_MAT_DIFFUSE) {
    if (woz < 0.f || wiz < 0.f) return mk3(0.f);
    return m.reflectance * BDPT_INV_PI;
  }
  if (m.type == BDPT_MAT_MICROFACET) {
    if (woz <= BDPT_EPS_F || wiz <= BDPT_EPS_F) return mk3(0.f);
    float G = 1.0f / (1.0f + mf_lambda(m.alpha, wiz) + mf_lambda(m.alpha, woz));
    return mf_F(m, fabsf(wiz)) * (G * mf_D_h(m.alpha, woz + wiz, hxy2) / (4.f * woz * wiz));
  }
  return mk3(0.f);
}
__device__ inline f3 bsdf_f(const Mat& m, f3 wo, f3 wi) { return bsdf_f_z(m, wo.z, wi.z, half_xy2(wo, wi)); }

// BSDF::sample_pdf(wo, wi): bsdf.cpp:76-81,112-117; advanced_bsdf.cpp:31-34,180-183,236-259.
// Microfacet: the reference aborts (advanced_bsdf.cpp:144-148); here it returns the half-vector density
// sample_f uses (needs the real wo).
__device__ inline float bsdf_pdf_z(const Mat& m, float woz, float wiz, float hxy2) {
  switch (m.type) {
    case BDPT_MAT_DIFFUSE:
    case BDPT_MAT_EMISSION:
      return cosine_pdf(wiz);
    case BDPT_MAT_MIRROR:
    case BDPT_MAT_REFRACTION:
      return 1.f;
    case BDPT_MAT_GLASS: {
      // refract(wi) then Schlick on the refracted cosine: only |refr.z| / |refr| is needed
      float eta = wiz > 0.f ? 1.0f / m.ior : m.ior;
      float s2 = eta * eta * (1.f - wiz * wiz);
      float z_sq = 1.f - s2;
      if (z_sq < 0.f) return 1.f;
      float R = glass_R_cos(m.ior, sqrtf(z_sq) * rsqrtf(s2 + z_sq), woz > 0.f);
      return wiz > 0.f ? R : 1.f - R;
    }
    case BDPT_MAT_MICROFACET: {
      if (woz <= BDPT_EPS_F || wiz <= BDPT_EPS_F) return 0.f;
      // p_theta * p_phi / sin(theta) of advanced_bsdf.cpp:126-130 = exp(-tan^2/alpha^2) / (pi alpha^2 cos^3), then
      // / (4 dot(wi, h)) with dot(wi, h) = |wo + wi| / 2
      const float sz = woz + wiz, sz2 = sz * sz, ia2 = 1.0f / (m.alpha * m.alpha);
      const float s2 = hxy2 + sz2, il = rsqrtf(s2);
      const float hz = sz * il;
      float pdf_h = expf(-hxy2 / sz2 * ia2) * ia2 * BDPT_INV_PI / (hz * hz * hz);
      return pdf_h * 0.5f * il;
    }
  }
  return 0.f;
}
__device__ inline float bsdf_pdf(const Mat& m, f3 wo, f3 wi) { return bsdf_pdf_z(m, wo.z, wi.z, half_xy2(wo, wi)); }

// The density the MIS chains use for "scatter toward wi at this vertex".  The reference always passes
// wo = (0,0,0) (bidirection.cpp:150,241), which is replicated for every BSDF it can evaluate that
// way.  Its MicrofacetBSDF::sample_pdf aborts instead (advanced_bsdf.cpp:144-148), so for microfacet — an
// extension with no reference behaviour to match — the real arrival direction wo is used.
__device__ inline float bsdf_pdf_mis_z(const Mat& m, float woz_real, float wiz, float hxy2) {
  const bool mf = m.type == BDPT_MAT_MICROFACET;
  return bsdf_pdf_z(m, mf ? woz_real : 0.f, wiz, hxy2);
}
__device__ inline float bsdf_pdf_mis(const Mat& m, f3 wo_real, f3 wi) {
  return bsdf_pdf_mis_z(m, wo_real.z, wi.z, half_xy2(wo_real, wi));
}

// BSDF::sample_f(wo, &wi, &pdf): bsdf.cpp:65-74 diffuse; :106-110 emission;
// advanced_bsdf.cpp:21-29 mirror; :163-175 refraction; :198-234 glass; :97-141 microfacet.
// Draw order follows the reference: two uniforms for a hemisphere sample (Xi1 then Xi2), one coin
// for glass only when refraction is possible.
template <class RNG>
__device__ inline f3 bsdf_sample_f(const Mat& m, f3 wo, f3* wi, float* pdf, RNG& rng) {
  switch (m.type) {
    case BDPT_MAT_DIFFUSE: {
      float u1 = rng.uniform(), u2 = rng.uniform();
      *wi = cosine_hemisphere(u1, u2, pdf);
      return bsdf_f(m, wo, *wi);
    }
    case BDPT_MAT_EMISSION: {
      float u1 = rng.uniform(), u2 = rng.uniform();
      *wi = cosine_hemisphere(u1, u2, pdf);
      return mk3(0.f);
    }
    case BDPT_MAT_MIRROR: {
      *wi = mk3(-wo.x, -wo.y, wo.z);
      *pdf = 1.f;
      float c = fabsf(wi->z) / norm(*wi);
      return m.reflectance * (1.0f / c);
    }
    case BDPT_MAT_REFRACTION: {
      if (!refract_local(wo, wi, m.ior)) {  // reference leaves wi/pdf undefined on TIR; zero throughput either way
        *wi = mk3(-wo.x, -wo.y, wo.z); *pdf = 1.f;
        return mk3(0.f);
      }
      *pdf = 1.f;
      float eta = wo.z > 0.f ? 1.0f / m.ior : m.ior;
      float c = fabsf(wi->z) / norm(*wi);
      return m.transmittance * (1.0f / (c * eta * eta));
    }
    case BDPT_MAT_GLASS: {
      f3 wr = mk3(-wo.x, -wo.y, wo.z), wt;
      if (!refract_local(wo, &wt, m.ior)) {
        *pdf = 1.f; *wi = wr;
        float c = fabsf(wr.z) / norm(wo);
        return m.reflectance * (1.0f / c);
      }
      float eta = wo.z > 0.f ? 1.0f / m.ior : m.ior;
      float R = glass_R(m.ior, wt, wo.z > 0.f);
      if (rng.uniform() < R) {
        *wi = wr; *pdf = R;
        float c = fabsf(wr.z) / norm(wr);
        return m.reflectance * (R / c);
      } else {
        *wi = wt; *pdf = 1.f - R;
        float c = fabsf(wt.z) / norm(wt);
        return m.transmittance * ((1.f - R) / (c * eta * eta));
      }
    }
    case BDPT_MAT_MICROFACET: {
      float rx = rng.uniform(), ry = rng.uniform();
      float tan2 = -m.alpha * m.alpha * logf(1.f - rx);
      float ct = rsqrtf(1.f + tan2), st = sqrtf(tan2) * ct;
      float s, c;
      sincos_2pi(ry, &s, &c);
      f3 h = mk3(st * c, st * s, ct);
      float nwo = norm(wo);
      float cth = dot(wo, h) / nwo;
      f3 hp = h * (cth * nwo);
      *wi = normalized(hp - (wo - hp));
      if (wo.z <= BDPT_EPS_F || wi->z <= BDPT_EPS_F) { *pdf = 1.f; *wi = mk3(0.f, 0.f, 1.f); return mk3(0.f); }
      float pdf_h = 2.f * expf(-tan2 / (m.alpha * m.alpha)) / (m.alpha * m.alpha * ct * ct * ct) * (0.5f * BDPT_INV_PI);
      *pdf = pdf_h / (4.f * dot(*wi, h));
      return bsdf_f(m, wo, *wi);
    }
  }
  *pdf = 1.f; *wi = mk3(0.f, 0.f, 1.f);
  return mk3(0.f);
}

// ------------------------------------------------------------------ lights
struct Light {
  int type;
  f3 radiance, position, direction, dim_x, dim_y;   // environment: position = centre, dim_x = {R, scene radius, -}
  float area;
};

// ---- environment light: an extension (the reference's EnvironmentLight aborts in every method BDPT calls,
// environment_light.cpp:182-208).  Defined in oracle/bdpt_oracle.cpp ("environment"): a sphere of radius R = 1000 x
// the scene's bounding radius around the scene's centre that emits inward — an area light like any other for every
// formula of the path.  Radiance toward the scene from sphere point c + R u = radiance * map(u), the map looked up
// like EnvironmentLight::sample_dir does (dir -> theta, phi -> x, y -> bilerp, :83-136,157-165).
__device__ inline f3 env_map_lookup(const DScene& sc, f3 dir) {
  if (!sc.envmap) return mk3(1.f);
  const int w = sc.env_w, h = sc.env_h;
  const f3 u = normalized(dir);
  const float theta = acosf(fminf(fmaxf(u.y, -1.f), 1.f));
  const float phi = atan2f(-u.z, u.x) + BDPT_PI;
  const float x = phi * (0.5f * BDPT_INV_PI) * (float)w, y = theta * BDPT_INV_PI * (float)h;
  int right = (int)lroundf(x), left, v = (int)lroundf(y);
  const float u1 = (float)right - x + .5f;
  float v1;
  if (right == 0 || right == w) { left = w - 1; right = 0; } else left = right - 1;
  if (v == 0) { v = 1; v1 = 1.f; } else if (v == h) { v = h - 1; v1 = 0.f; } else v1 = (float)v - y + .5f;
  const int bottom = w * v, top = bottom - w;
  const float u0 = 1.f - u1;
  const float4 tl = __ldg(sc.envmap + top + left), tr = __ldg(sc.envmap + top + right);
  const float4 bl = __ldg(sc.envmap + bottom + left), br = __ldg(sc.envmap + bottom + right);
  return (mk3(tl.x, tl.y, tl.z) * u1 + mk3(tr.x, tr.y, tr.z) * u0) * v1 + (mk3(bl.x, bl.y, bl.z) * u1 + mk3(br.x, br.y, br.z) * u0) * (1.f - v1);
}
__device__ inline f3 env_radiance(const DScene& sc, const Light& l, f3 dir_from_centre) { return l.radiance * env_map_lookup(sc, dir_from_centre); }
// emission direction density at a sphere point (solid angle): the directions that reach the scene's bounding disk
__device__ inline float env_dir_pdf(const Light& l, float cos_at_sphere) {
  const float R = l.dim_x.x, rs = l.dim_x.y;
  return cos_at_sphere > 0.f ? R * R * cos_at_sphere / (BDPT_PI * rs * rs) : 0.f;
}

__device__ inline Light load_light(const DScene& sc, int i) {
  const float4* p = sc.lights + 5 * i;
  float4 a = ldg4(p), b = ldg4(p + 1), c = ldg4(p + 2), d = ldg4(p + 3), e = ldg4(p + 4);
  Light l;
  l.radiance = mk3(a.x, a.y, a.z); l.type = __float_as_int(a.w);
  l.position = mk3(b.x, b.y, b.z); l.area = b.w;
  l.direction = mk3(c.x, c.y, c.z);
  l.dim_x = mk3(d.x, d.y, d.z); l.dim_y = mk3(e.x, e.y, e.z);
  return l;
}

// AreaLight::contain_point light.cpp:257-262 (a plane test only); PointLight :136-139
__device__ inline bool light_contain_point(const Light& l, f3 p) {
  if (l.type == BDPT_LIGHT_ENV) return fabsf(norm(p - l.position) - l.dim_x.x) < 1e-3f * l.dim_x.x;
  if (l.type == BDPT_LIGHT_AREA) {
    f3 d = normalized(l.position - p);
    return fabsf(dot(d, l.direction)) < BDPT_EPS_F;
  }
  return norm(p - l.position) < BDPT_EPS_F;
}

// AreaLight::sample_pdf light.cpp:264-284; PointLight::sample_pdf :141-151.
// wi points TOWARD the light (world space).  Returns Le.
__device__ inline f3 light_sample_pdf(const DScene& sc, const Light& l, f3 p, f3 wi, float* point_pdf, float* dir_pdf) {
  if (!light_contain_point(l, p)) { *point_pdf = 0.f; *dir_pdf = 0.f; return mk3(0.f); }
  if (l.type == BDPT_LIGHT_POINT) { *point_pdf = 1.f; *dir_pdf = 0.25f * BDPT_INV_PI; return l.radiance; }
  if (l.type == BDPT_LIGHT_ENV) {
    const f3 u = normalized(p - l.position);       // outward; the sphere emits along -u, wi points TOWARD the light
    *point_pdf = 1.0f / l.area;
    *dir_pdf = env_dir_pdf(l, is_zero(wi) ? 1.f : dot(u, normalized(wi)));
    return *dir_pdf > 0.f ? env_radiance(sc, l, u) : mk3(0.f);
  }
  *point_pdf = 1.0f / l.area;
  // local z of normalize(w2o * (-wi)); w2o's z row is the unit light direction
  f3 zdir = normalized(l.direction);
  float z = dot(zdir, -wi) / norm(wi);
  *dir_pdf = cosine_pdf(z);
  return *dir_pdf > 0.f ? l.radiance : mk3(0.f);
}

// AreaLight::sample_Le light.cpp:219-232; PointLight::sample_Le :115-123
__device__ inline f3 light_sample_Le(const DScene& sc, const Light& l, Rng& rng, f3* o, f3* d, float* point_pdf, float* dir_pdf, f3* normal) {
  if (l.type == BDPT_LIGHT_POINT) {
    float u1 = rng.uniform(), u2 = rng.uniform();
    *o = l.position; *d = uniform_sphere(u1, u2);
    *point_pdf = 1.f; *dir_pdf = 0.25f * BDPT_INV_PI; *normal = *d;
    return l.radiance;
  }
  if (l.type == BDPT_LIGHT_ENV) {
    // a uniform direction of travel, a uniform point of the scene's bounding disk perpendicular to it, carried back
    // onto the sphere
    const float u1 = rng.uniform(), u2 = rng.uniform(), u3 = rng.uniform(), u4 = rng.uniform();
    const f3 dd = uniform_sphere(u1, u2);
    const Frame fr = make_frame(dd);
    float sn, cs;
    sincos_2pi(u4, &sn, &cs);
    const float rr = l.dim_x.y * sqrtf(u3);
    const f3 p0 = (l.position - dd * l.dim_x.x) + fr.x * (rr * cs) + fr.y * (rr * sn);
    const f3 u = normalized(p0 - l.position);
    *o = l.position + u * l.dim_x.x; *d = dd; *normal = -u;
    *point_pdf = 1.0f / l.area;
    *dir_pdf = env_dir_pdf(l, -dot(u, dd));
    return env_radiance(sc, l, u);
  }
  float sx = rng.uniform() - 0.5f, sy = rng.uniform() - 0.5f;
  *o = l.position + sx * l.dim_x + sy * l.dim_y;
  float u1 = rng.uniform(), u2 = rng.uniform();
  f3 dl = cosine_hemisphere(u1, u2, dir_pdf);
  Frame fr = make_frame(l.direction);
  *d = fr.to_world(dl);
  *point_pdf = 1.0f / l.area;
  *normal = l.direction;
  return l.radiance;
}

// AreaLight::sample_Le_point light.cpp:234-255; PointLight::sample_Le_point :125-134
__device__ inline f3 light_sample_Le_point(const DScene& sc, const Light& l, f3 p, Rng& rng, f3* point, float* point_pdf, float* dir_pdf,
                                           f3* normal) {
  if (l.type == BDPT_LIGHT_POINT) {
    f3 wi = normalized(l.position - p);
    *point = l.position; *point_pdf = 1.f; *dir_pdf = 0.25f * BDPT_INV_PI; *normal = -wi;
    return l.radiance;
  }
  if (l.type == BDPT_LIGHT_ENV) {
    const float u1 = rng.uniform(), u2 = rng.uniform();
    const f3 u = uniform_sphere(u1, u2);           // a uniform point of the sphere
    *point = l.position + u * l.dim_x.x;
    const f3 wi = normalized(*point - p);
    *point_pdf = 1.0f / l.area; *normal = -u;
    *dir_pdf = env_dir_pdf(l, dot(u, wi));
    return *dir_pdf > 0.f ? env_radiance(sc, l, u) : mk3(0.f);
  }
  float sx = rng.uniform() - 0.5f, sy = rng.uniform() - 0.5f;
  *point = l.position + sx * l.dim_x + sy * l.dim_y;
  f3 d = *point - p;
  float cos_theta = dot(d, l.direction);
  f3 wi = d * (1.0f / norm(d));
  *point_pdf = 1.0f / l.area;
  *normal = l.direction;
  *dir_pdf = cosine_pdf(dot(normalized(l.direction), -wi));
  return cos_theta < 0.f ? l.radiance : mk3(0.f);
}

// ------------------------------------------------------------------ camera
// Camera::generate_ray camera.cpp:191-212 (x, y in [0,1])
__device__ inline void camera_generate_ray(const DCamera& c, float x, float y, f3* o, f3* d) {
  f3 dir = mk3((2.f * x - 1.f) * c.tan_h, (2.f * y - 1.f) * c.tan_v, -1.f);
  f3 w = dir.x * c.c2w_x + dir.y * c.c2w_y + dir.z * c.c2w_z;
  *o = c.pos;
  *d = normalized(w);
}

// ---- thin lens: an extension.  The reference has Camera::generate_ray_for_thin_lens (camera_lens.cpp:22-43) but its
// BDPT never calls it.  With lens_radius > 0 the eye subpath starts at a uniformly sampled point of the
// lens (that function, verbatim) and a light-tracing connection (t = 1) goes through a fresh point of the lens: the
// formulas of sample_ray_pdf with the lens point in place of the pinhole and the angle of the actual ray; the raster
// position is that of the sensor sample whose rays converge where this ray meets the plane of focus.  t = 0 (a light
// subpath that crosses the lens) is kernels.cuh's lens_strategy.
__device__ inline f3 lens_point(const DCamera& c, float rnd_r, float rnd_u) {   // camera space, z = 0
  float sn, cs;
  sincos_2pi(rnd_u, &sn, &cs);
  const float r = c.lens_radius * sqrtf(rnd_r);
  return mk3(r * cs, r * sn, 0.f);
}
// Camera::generate_ray_for_thin_lens camera_lens.cpp:22-43 (rndTheta = 2 pi rnd_u)
__device__ inline void camera_generate_ray_lens(const DCamera& c, float x, float y, float rnd_r, float rnd_u, f3* o, f3* d) {
  const f3 pl = lens_point(c, rnd_r, rnd_u);
  const f3 pf = mk3((2.f * x - 1.f) * c.tan_h, (2.f * y - 1.f) * c.tan_v, -1.f) * c.focal_dist;
  const f3 dir = pf - pl;
  *o = c.pos + pl.x * c.c2w_x + pl.y * c.c2w_y;
  *d = normalized(dir.x * c.c2w_x + dir.y * c.c2w_y + dir.z * c.c2w_z);
}
// returns We; *eye = the lens point (world), wi from p toward it
__device__ inline float camera_sample_ray_pdf_lens(const DCamera& c, f3 p, f3 lens_c, f3* eye, f3* wi, float* dist, float* dir_pdf, int* px, int* py) {
  *eye = c.pos + lens_c.x * c.c2w_x + lens_c.y * c.c2w_y;
  const f3 v = *eye - p;
  *dist = norm(v);
  *wi = v * (1.0f / *dist);
  const f3 mw = -(*wi);
  const f3 wc = mk3(dot(c.w2c_r0, mw), dot(c.w2c_r1, mw), dot(c.w2c_r2, mw));   // lens -> p, camera space
  const float cos_t = -wc.z;
  const float c2 = cos_t * cos_t;
  const float We = c2 * c2 / (4.f * c.tan_h * c.tan_v);
  *dir_pdf = (*dist * *dist) / cos_t;
  *px = -1; *py = -1;
  if (cos_t > 0.f) {
    const float sf = c.focal_dist / cos_t;
    const float qx = (lens_c.x + wc.x * sf) / c.focal_dist, qy = (lens_c.y + wc.y * sf) / c.focal_dist;
    const float fx = ((qx / c.tan_h + 1.f) * 0.5f) * (float)c.W, fy = ((qy / c.tan_v + 1.f) * 0.5f) * (float)c.H;
    *px = (fx > -2.0e9f && fx < 2.0e9f) ? __float2int_rz(fx) : -1;
    *py = (fy > -2.0e9f && fy < 2.0e9f) ? __float2int_rz(fy) : -1;
  }
  return We;
}

// Camera::sample_ray_pdf camera.cpp:214-248.  Returns We (a scalar; the reference returns it in all
// three channels); wi points from p to the pinhole.
__device__ inline float camera_sample_ray_pdf(const DCamera& c, f3 p, f3* wi, float* dist, float* dir_pdf, int* px, int* py) {
  f3 v = c.pos - p;
  *dist = norm(v);
  *wi = v * (1.0f / *dist);
  f3 mw = -(*wi);
  f3 wc = mk3(dot(c.w2c_r0, mw), dot(c.w2c_r1, mw), -dot(c.w2c_r2, mw));
  float cos_t = wc.z;  // cos(acos(wc.z))
  float c2 = cos_t * cos_t;
  float We = c2 * c2 / (4.f * c.tan_h * c.tan_v);
  *dir_pdf = (*dist * *dist) / cos_t;
  float inv = 1.0f / wc.z;
  float fx = ((wc.x * inv / c.tan_h + 1.f) * 0.5f) * (float)c.W;
  float fy = ((wc.y * inv / c.tan_v + 1.f) * 0.5f) * (float)c.H;
  // C truncation toward zero, as the reference's double -> int conversion (camera.cpp:242-243)
  *px = (fx > -2.0e9f && fx < 2.0e9f) ? __float2int_rz(fx) : -1;
  *py = (fy > -2.0e9f && fy < 2.0e9f) ? __float2int_rz(fy) : -1;
  // A point BEHIND the pinhole has no raster position (the reference divides by wc.z whatever its sign and would
  // splat its mirror image; nothing in its scenes lies behind the camera, the environment's sphere does)
  if (!(wc.z > 0.f)) { *px = -1; *py = -1; }
  return We;
}

}  // namespace bdpt
