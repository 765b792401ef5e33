// dae_scene.cpp — COLLADA (.dae, with the CGL <extra> profile) -> flat scene.
//
// Stands where the reference's ingest path does, end to end:
//   ColladaParser::load / parse_node / parse_camera / parse_light / parse_sphere / parse_polymesh /
//   parse_material                         (reference src/scene/collada/collada.cpp:133-938)
//   Application::load, init_sphere, init_light, camera placement   (src/application/application.cpp:228-348)
//   GLScene::Mesh / Sphere / AreaLight / PointLight constructors   (src/scene/gl_scene/*.h, mesh.cpp:22-43)
//   HalfedgeMesh::build + Vertex::computeNormal                     (src/util/halfEdgeMesh.cpp:29-404, halfEdgeMesh.h:492-515)
//   SceneObjects::Mesh::Mesh / get_primitives, RaytracedRenderer::build_accel order
//                                          (src/scene/object.cpp:16-58, src/pathtracer/raytraced_renderer.cpp:355-360)
//   Camera::configure / place / compute_position                    (src/pathtracer/camera.cpp:29-65,123-147)
// and produces exactly what the reference hands its integrator, as the POD scene of
// include/bdpt_flatscene.h.  tests/test_host_loader.py compares the result with the flat scenes the
// reference's own loader produced (tests/golden/scenes/).
//
// Behaviours kept on purpose (they decide the numbers): vertices and radii are parsed as fp32 then widened;
// a face contributes only the triangle (v[n-1], v[0], v[1]) (object.cpp:36-41 starts at the face's LAST halfedge);
// vertex normals are the reference's area-weighted sums INCLUDING its walk for boundary vertices
// (h = h->next()->twin(), halfEdgeMesh.h:498-504, which leaves the vertex); the 4x4 product is the AVX
// form C(i,j) = dot(A[i], B column j) with A[i] a COLUMN of A (CGL/src/matrix4x4.cpp:131-132), i.e. A^T B;
// the camera is placed, not read, at 800x600 (F9).
// Defined here, undefined in the reference: <rotate>/<translate>/<scale> start from the identity (the
// reference uses an uninitialised matrix, collada.cpp:276-330; only <matrix> is reliable there).
#include "dae_scene.h"

#include <algorithm>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <map>
#include <set>

#include "xml.h"

namespace bdpt_host {
namespace {

const double PI = 3.14159265358979323;  // CGL/include/CGL/misc.h:11
const double EPS_F = (double)0.00001f;
const double INF_D = std::numeric_limits<double>::infinity();

struct V3 {
  double x = 0, y = 0, z = 0;
  V3() {}
  V3(double a, double b, double c) : x(a), y(b), z(c) {}
};
inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, double c) { return V3(a.x * c, a.y * c, a.z * c); }
inline double dot(V3 u, V3 v) { return (u.x * v.x + u.y * v.y) + u.z * v.z; }
inline double norm(V3 v) { return std::sqrt(dot(v, v)); }
inline V3 unit(V3 v) { return v * (1. / norm(v)); }
inline V3 cross(V3 u, V3 v) { return V3(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x); }
inline V3 div(V3 a, double c) { const double rc = 1.0 / c; return V3(rc * a.x, rc * a.y, rc * a.z); }

struct V4 { double x, y, z, w; };
struct M4 {
  double e[4][4];  // e[j][i] = row i, column j (column-major like CGL)
  static M4 identity() { M4 m; for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) m.e[j][i] = i == j ? 1. : 0.; return m; }
  double& at(int i, int j) { return e[j][i]; }
  double at(int i, int j) const { return e[j][i]; }
  V4 mul(V4 v) const {  // matrix4x4.cpp:146-149
    V4 r;
    r.x = ((v.x * e[0][0] + v.y * e[1][0]) + v.z * e[2][0]) + v.w * e[3][0];
    r.y = ((v.x * e[0][1] + v.y * e[1][1]) + v.z * e[2][1]) + v.w * e[3][1];
    r.z = ((v.x * e[0][2] + v.y * e[1][2]) + v.z * e[2][2]) + v.w * e[3][2];
    r.w = ((v.x * e[0][3] + v.y * e[1][3]) + v.z * e[2][3]) + v.w * e[3][3];
    return r;
  }
};
// Matrix4x4::operator* as compiled with AVX (matrix4x4.cpp:125-141): C(i,j) = dot(A[i], column j of B),
// where A[i] is column i of A.
M4 mat_mul(const M4& A, const M4& B) {
  M4 C;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      C.at(i, j) = ((A.e[i][0] * B.at(0, j) + A.e[i][1] * B.at(1, j)) + A.e[i][2] * B.at(2, j)) + A.e[i][3] * B.at(3, j);
  return C;
}
inline V3 to3D(V4 v) { return V3(v.x, v.y, v.z); }
inline V3 project3D(V4 v) { double iw = 1.0 / v.w; return V3(v.x * iw, v.y * iw, v.z * iw); }
inline V4 v4(V3 v, double w) { return V4{v.x, v.y, v.z, w}; }

enum class Kind { NONE, CAMERA, LIGHT, SPHERE, POLYMESH };
enum class LightKind { NONE, AMBIENT, DIRECTIONAL, AREA, POINT, SPOT };

struct NodeDesc {
  Kind kind = Kind::NONE;
  M4 transform = M4::identity();
  // camera
  float hFov = 50.f, vFov = 35.f, nClip = 0.001f, fClip = 1000.f;
  // light
  LightKind light = LightKind::NONE;
  V3 spectrum = V3(1, 1, 1);
  // sphere
  float radius = 0.f;
  // mesh
  std::vector<V3> vertices;
  std::vector<std::vector<size_t>> polygons;
  bool has_material = false;
  BdptMaterial material;
};

// iostream-free number parsing (this library is dlopen'ed by Python; it keeps to the C runtime).
// Like a failed `stream >> x`, values that cannot be read are left as they were.
static void parse_doubles(const char* p, double** dst, int n) {
  char* end = nullptr;
  for (int i = 0; i < n; i++) {
    double v = strtod(p, &end);
    if (end == p) return;
    *dst[i] = v; p = end;
  }
}

struct Loader {
  std::map<std::string, XmlNode*> sources;
  M4 transform = M4::identity();
  V3 up = V3(0, 1, 0);
  std::vector<NodeDesc> nodes;
  std::string err;

  void fail(const std::string& m) { throw std::runtime_error(m); }

  void uri_load(XmlNode* x) {  // collada.cpp:56-69
    if (const char* id = x->attr("id")) sources[id] = x;
    for (auto& c : x->children) uri_load(c.get());
  }
  XmlNode* uri_find(const std::string& id) { auto it = sources.find(id); return it == sources.end() ? nullptr : it->second; }
  XmlNode* get_element(XmlNode* x, const std::string& query) {  // collada.cpp:78-99
    XmlNode* e = x;
    size_t b = 0;
    while (e && b <= query.size()) {
      size_t q = query.find('/', b);
      std::string tok = query.substr(b, q == std::string::npos ? std::string::npos : q - b);
      e = e->first_child(tok.c_str());
      if (q == std::string::npos) break;
      b = q + 1;
    }
    if (e) if (const char* url = e->attr("url")) e = uri_find(std::string(url + 1));
    return e;
  }
  XmlNode* technique_common(XmlNode* x) {  // collada.cpp:102-117
    if (XmlNode* prof = x->first_child("profile_COMMON"))
      for (XmlNode* t = prof->first_child("technique"); t; t = t->next_sibling("technique")) {
        const char* sid = t->attr("sid");
        if (sid && std::string(sid) == "common") return t;
      }
    return x->first_child("technique_common");
  }
  XmlNode* technique_CGL(XmlNode* x) {  // collada.cpp:120-131
    for (XmlNode* t = get_element(x, "extra/technique"); t; t = t->next_sibling("technique")) {
      const char* p = t->attr("profile");
      if (p && std::string(p) == "CGL") return t;
    }
    return nullptr;
  }
  static const char* text(XmlNode* e) { return e && e->has_text ? e->text.c_str() : ""; }
  static V3 spectrum_from_string(const char* s) {  // collada.cpp:29-40
    V3 v; double* d[3] = {&v.x, &v.y, &v.z}; parse_doubles(s, d, 3); return v;
  }
  static void set3(double* d, V3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }

  void parse_material(XmlNode* xml, BdptMaterial* out) {  // collada.cpp:854-938
    memset(out, 0, sizeof *out);
    out->type = BDPT_MAT_DIFFUSE;
    XmlNode* e_effect = get_element(xml, "instance_effect");
    if (!e_effect) fail("no target effects found for material");
    XmlNode* tc = technique_common(e_effect);
    XmlNode* tg = technique_CGL(e_effect);
    const V3 grey((double).5f, (double).5f, (double).5f);
    if (tg) {
      bool any = false;
      for (XmlNode* b = tg->first_child(); b; b = b->next_sibling()) {
        const std::string& type = b->name;
        BdptMaterial m; memset(&m, 0, sizeof m);
        if (type == "emission") {
          m.type = BDPT_MAT_EMISSION; set3(m.radiance, spectrum_from_string(text(get_element(b, "radiance"))));
        } else if (type == "mirror") {
          m.type = BDPT_MAT_MIRROR; set3(m.reflectance, spectrum_from_string(text(get_element(b, "reflectance"))));
        } else if (type == "microfacet") {
          m.type = BDPT_MAT_MICROFACET;
          m.alpha = (double)(float)atof(text(get_element(b, "alpha")));
          set3(m.eta, spectrum_from_string(text(get_element(b, "eta"))));
          set3(m.k, spectrum_from_string(text(get_element(b, "k"))));
        } else if (type == "refraction") {
          m.type = BDPT_MAT_REFRACTION;
          set3(m.transmittance, spectrum_from_string(text(get_element(b, "transmittance"))));
          m.roughness = (double)(float)atof(text(get_element(b, "roughness")));
          m.ior = (double)(float)atof(text(get_element(b, "ior")));
        } else if (type == "glass") {
          m.type = BDPT_MAT_GLASS;
          set3(m.transmittance, spectrum_from_string(text(get_element(b, "transmittance"))));
          set3(m.reflectance, spectrum_from_string(text(get_element(b, "reflectance"))));
          m.roughness = (double)(float)atof(text(get_element(b, "roughness")));
          m.ior = (double)(float)atof(text(get_element(b, "ior")));
        } else continue;
        *out = m; any = true;  // the last recognised BSDF wins
      }
      if (!any) fail("CGL material profile without a recognised BSDF");  // the reference would leave a null BSDF and crash
    } else if (tc) {
      XmlNode* d = get_element(tc, "phong/diffuse/color");
      set3(out->reflectance, d ? spectrum_from_string(text(d)) : grey);
    } else {
      set3(out->reflectance, grey);
    }
  }

  void parse_camera(XmlNode* xml, NodeDesc& n) {  // collada.cpp:434-475
    n.kind = Kind::CAMERA;
    XmlNode* p = get_element(xml, "optics/technique_common/perspective");
    if (!p) fail("no perspective defined in camera");
    XmlNode *xf = p->first_child("xfov"), *yf = p->first_child("yfov"), *zn = p->first_child("znear"), *zf = p->first_child("zfar");
    n.hFov = xf ? (float)atof(text(xf)) : 50.0f;
    n.vFov = yf ? (float)atof(text(yf)) : 35.0f;
    n.nClip = zn ? (float)atof(text(zn)) : 0.001f;
    n.fClip = zf ? (float)atof(text(zf)) : 1000.0f;
    if (!yf) {
      XmlNode* ar = get_element(p, "aspect_ratio");
      if (!ar) fail("incomplete perspective definition");
      float aspect = (float)atof(text(ar));
      n.vFov = (float)(2 * ((std::atan(std::tan((0.5 * n.hFov) * (PI / 180)) / aspect)) * (180 / PI)));
    }
  }

  void parse_light(XmlNode* xml, NodeDesc& n) {  // collada.cpp:477-578
    n.kind = Kind::LIGHT;
    XmlNode* tg = technique_CGL(xml);
    XmlNode* t = tg ? tg : technique_common(xml);
    if (!t) fail("no supported profile defined in light");
    XmlNode* l = t->first_child();
    if (!l) return;
    const std::string& type = l->name;
    if (type == "ambient") n.light = LightKind::AMBIENT;
    else if (type == "directional") n.light = LightKind::DIRECTIONAL;
    else if (type == "area") n.light = LightKind::AREA;
    else if (type == "point") n.light = LightKind::POINT;
    else if (type == "spot") n.light = LightKind::SPOT;
    else fail("light type " + type + " is not supported");
    XmlNode* c = get_element(l, "color");
    if (!c) fail("no color definition in light");
    n.spectrum = spectrum_from_string(text(c));
  }

  void parse_sphere(XmlNode* xml, NodeDesc& n) {  // collada.cpp:580-603
    n.kind = Kind::SPHERE;
    XmlNode* t = technique_CGL(xml);
    if (!t) fail("no CGL profile technique in sphere geometry");
    XmlNode* r = get_element(t, "sphere/radius");
    if (!r) fail("invalid sphere definition");
    n.radius = (float)atof(text(r));
  }

  void parse_polymesh(XmlNode* xml, NodeDesc& n) {  // collada.cpp:606-852 (positions and vertex indices only)
    n.kind = Kind::POLYMESH;
    XmlNode* mesh = xml->first_child("mesh");
    if (!mesh) fail("no mesh data defined in geometry");
    std::map<std::string, std::vector<float>> arr;
    for (XmlNode* s = mesh->first_child("source"); s; s = s->next_sibling("source")) {
      XmlNode* fa = s->first_child("float_array");
      const char* id = s->attr("id");
      if (!fa || !id) continue;
      size_t count = (size_t)atoll(fa->attr("count") ? fa->attr("count") : "0");
      std::vector<float> f; f.reserve(count);
      const char* p = text(fa);
      char* end = nullptr;
      for (size_t i = 0; i < count; i++) {
        float v = strtof(p, &end);
        if (end == p) v = f.empty() ? 0.f : f.back();   // a failed stream extraction leaves the last value
        f.push_back(v); p = end;
      }
      arr[id] = std::move(f);
    }
    XmlNode* ev = mesh->first_child("vertices");
    if (!ev) fail("no vertices defined in geometry");
    std::string vertices_id = ev->attr("id") ? ev->attr("id") : "";
    std::vector<V3> vertices;
    for (XmlNode* in = ev->first_child("input"); in; in = in->next_sibling("input")) {
      const char* sem = in->attr("semantic");
      if (!sem || std::string(sem) != "POSITION") continue;
      std::string src = in->attr("source") ? in->attr("source") + 1 : "";
      auto it = arr.find(src);
      if (it == arr.end()) fail("undefined input source " + src);
      const std::vector<float>& f = it->second;
      for (size_t i = 0; i + 2 < f.size(); i += 3) vertices.push_back(V3(f[i], f[i + 1], f[i + 2]));
    }
    XmlNode* pl = mesh->first_child("polylist");
    if (!pl) return;
    bool hv = false, hn = false, ht = false;
    size_t vo = 0;
    for (XmlNode* in = pl->first_child("input"); in; in = in->next_sibling("input")) {
      std::string sem = in->attr("semantic") ? in->attr("semantic") : "";
      std::string src = in->attr("source") ? in->attr("source") + 1 : "";
      size_t off = (size_t)atoll(in->attr("offset") ? in->attr("offset") : "0");
      if (sem == "VERTEX") {
        hv = true; vo = off;
        if (src != vertices_id) fail("undefined source for VERTEX semantic: " + src);
        n.vertices = vertices;
      } else if (sem == "NORMAL") hn = true;
      else if (sem == "TEXCOORD") ht = true;
    }
    size_t npoly = (size_t)atoll(pl->attr("count") ? pl->attr("count") : "0");
    size_t stride = (hv ? 1 : 0) + (hn ? 1 : 0) + (ht ? 1 : 0);
    XmlNode* evc = pl->first_child("vcount");
    if (!evc) fail("polygon sizes undefined in geometry");
    std::vector<size_t> sizes; size_t nidx = 0;
    {
      const char* p = text(evc); char* end = nullptr;
      for (size_t i = 0; i < npoly; i++) { size_t s = (size_t)strtoull(p, &end, 10); p = end; sizes.push_back(s); nidx += s * stride; }
    }
    XmlNode* ep = pl->first_child("p");
    if (!ep) fail("no index array defined in geometry");
    std::vector<size_t> idx; idx.reserve(nidx);
    {
      const char* p = text(ep); char* end = nullptr;
      for (size_t i = 0; i < nidx; i++) { idx.push_back((size_t)strtoull(p, &end, 10)); p = end; }
    }
    n.polygons.resize(npoly);
    if (hv) {
      size_t k = 0;
      for (size_t i = 0; i < npoly; i++)
        for (size_t j = 0; j < sizes[i]; j++) { n.polygons[i].push_back(idx[k * stride + vo]); k++; }
    }
  }

  void parse_node(XmlNode* xml) {  // collada.cpp:236-432
    NodeDesc node;
    for (XmlNode* e = xml->first_child(); e; e = e->next_sibling()) {
      const std::string& name = e->name;
      if (name == "matrix") {
        M4 m = M4::identity();
        double* d[16];
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) d[4 * i + j] = &m.at(i, j);
        parse_doubles(text(e), d, 16);
        node.transform = m;
        break;  // the reference stops reading transforms at the first <matrix>
      }
      if (name == "rotate") {
        M4 m = M4::identity();
        std::string sid = e->attr("sid") ? e->attr("sid") : "";
        double* d[4] = {nullptr, nullptr, nullptr, nullptr};
        switch (sid.empty() ? 0 : sid.back()) {
          case 'X': d[0] = &m.at(1, 1); d[1] = &m.at(1, 2); d[2] = &m.at(2, 1); d[3] = &m.at(2, 2); break;
          case 'Y': d[0] = &m.at(0, 0); d[1] = &m.at(2, 0); d[2] = &m.at(0, 2); d[3] = &m.at(2, 2); break;
          case 'Z': d[0] = &m.at(0, 0); d[1] = &m.at(0, 1); d[2] = &m.at(1, 0); d[3] = &m.at(1, 1); break;
          default: break;
        }
        if (d[0]) parse_doubles(text(e), d, 4);
        node.transform = mat_mul(m, node.transform);
      }
      if (name == "translate") {
        M4 m = M4::identity();
        double* d[3] = {&m.at(0, 3), &m.at(1, 3), &m.at(2, 3)};
        parse_doubles(text(e), d, 3);
        node.transform = mat_mul(m, node.transform);
      }
      if (name == "scale") {
        M4 m = M4::identity();
        double* d[3] = {&m.at(0, 0), &m.at(1, 1), &m.at(1, 1)};  // sic: collada.cpp:321 reads (1,1) twice
        parse_doubles(text(e), d, 3);
        node.transform = mat_mul(m, node.transform);
      }
    }
    M4 save = transform;
    node.transform = mat_mul(transform, node.transform);
    transform = node.transform;
    for (XmlNode* c = get_element(xml, "node"); c; c = c->next_sibling("node")) parse_node(c);
    transform = save;

    XmlNode* e_camera = get_element(xml, "instance_camera");
    XmlNode* e_light = get_element(xml, "instance_light");
    XmlNode* e_geom = get_element(xml, "instance_geometry");
    if (e_camera) parse_camera(e_camera, node);
    else if (e_light) parse_light(e_light, node);
    else if (e_geom) {
      bool is_mesh = get_element(e_geom, "mesh") != nullptr;
      bool is_sphere = !is_mesh && get_element(e_geom, "extra") != nullptr;
      if (is_mesh) parse_polymesh(e_geom, node);
      else if (is_sphere) parse_sphere(e_geom, node);
      if (is_mesh || is_sphere) {
        XmlNode* im = get_element(xml, "instance_geometry/bind_material/technique_common/instance_material");
        if (im) {
          const char* target = im->attr("target");
          if (!target) fail("no target material in instance");
          XmlNode* em = uri_find(std::string(target + 1));
          if (!em) fail(std::string("invalid target material id: ") + (target + 1));
          parse_material(em, &node.material);
          node.has_material = true;
        }
      }
    }
    nodes.push_back(std::move(node));  // post-order: children first (collada.cpp:431)
  }

  void load_file(const std::string& path) {  // collada.cpp:133-227
    FILE* in = fopen(path.c_str(), "rb");
    if (!in) fail("cannot open " + path);
    std::string data;
    char chunk[1 << 16];
    size_t got;
    while ((got = fread(chunk, 1, sizeof chunk, in)) > 0) data.append(chunk, got);
    fclose(in);
    XmlParser parser(data);
    std::unique_ptr<XmlNode> doc = parser.parse();
    if (doc->name != "COLLADA") fail("not a COLLADA file");
    XmlNode* root = doc.get();
    uri_load(root);
    if (XmlNode* asset = get_element(root, "asset")) {
      XmlNode* ua = get_element(asset, "up_axis");
      if (!ua) fail("no up direction defined in COLLADA file");
      std::string up_dir = text(ua);
      transform = M4::identity();
      if (up_dir == "X_UP") {
        transform.at(0, 0) = 0; transform.at(0, 1) = 1; transform.at(1, 0) = 1; transform.at(1, 1) = 0; transform.at(2, 2) = -1;
        up = V3(1, 0, 0);
      } else if (up_dir == "Z_UP") {
        transform.at(1, 1) = 0; transform.at(1, 2) = 1; transform.at(2, 1) = 1; transform.at(2, 2) = 0; transform.at(0, 0) = -1;
        up = V3(0, 0, 1);
      } else if (up_dir == "Y_UP") {
        up = V3(0, 1, 0);
      } else fail("invalid up direction in COLLADA file");
    }
    XmlNode* scene = get_element(root, "scene/instance_visual_scene");
    if (!scene) fail("no scene description found in file");
    for (XmlNode* n = get_element(scene, "node"); n; n = n->next_sibling("node")) parse_node(n);
    keep_alive = std::move(doc);
  }
  std::unique_ptr<XmlNode> keep_alive;
};

// ---------------------------------------------------------------------------- half-edge mesh
// HalfedgeMesh::build (halfEdgeMesh.cpp:29-404) with index-based links, same construction order.
struct HE { int twin = -1, next = -1, vertex = -1, face = -1; };
struct HalfedgeMesh {
  std::vector<HE> he;
  std::vector<int> face_he;       // per real face: its halfedge (the LAST one created for the polygon)
  std::vector<char> face_boundary;  // per face (real faces first, then boundary loops)
  std::vector<int> vert_he;       // per vertex (creation order)
  std::vector<V3> pos, nrm;       // per vertex (creation order)
  size_t n_real_faces = 0;

  bool he_is_boundary(int h) const { return face_boundary[he[h].face] != 0; }

  void build(const std::vector<std::vector<size_t>>& polygons, const std::vector<V3>& positions) {
    std::map<size_t, int> index_to_vertex;
    std::map<int, size_t> degree;
    for (const auto& p : polygons) {
      if (p.size() < 3) throw std::runtime_error("each polygon must have at least three vertices");
      std::set<size_t> distinct;
      for (size_t i : p) {
        distinct.insert(i);
        auto it = index_to_vertex.find(i);
        if (it == index_to_vertex.end()) { int v = (int)vert_he.size(); vert_he.push_back(-1); index_to_vertex[i] = v; degree[v] = 1; }
        else degree[it->second]++;
      }
      if (distinct.size() < p.size()) throw std::runtime_error("a polygon does not have distinct vertices");
    }
    n_real_faces = polygons.size();
    face_he.assign(n_real_faces, -1);
    face_boundary.assign(n_real_faces, 0);
    std::map<std::pair<size_t, size_t>, int> pair_to_he;
    for (size_t f = 0; f < polygons.size(); f++) {
      const auto& p = polygons[f];
      std::vector<int> fh;
      size_t deg = p.size();
      for (size_t i = 0; i < deg; i++) {
        size_t a = p[i], b = p[(i + 1) % deg];
        if (pair_to_he.count({a, b})) throw std::runtime_error("mesh is non-manifold or inconsistently oriented");
        int hab = (int)he.size(); he.push_back(HE());
        pair_to_he[{a, b}] = hab;
        he[hab].face = (int)f; face_he[f] = hab;
        he[hab].vertex = index_to_vertex[a]; vert_he[he[hab].vertex] = hab;
        fh.push_back(hab);
        auto iba = pair_to_he.find({b, a});
        if (iba != pair_to_he.end()) { he[hab].twin = iba->second; he[iba->second].twin = hab; }
        else he[hab].twin = -1;
      }
      for (size_t i = 0; i < deg; i++) he[fh[i]].next = fh[(i + 1) % deg];
    }
    // boundary vertices point at their boundary halfedge
    for (size_t v = 0; v < vert_he.size(); v++) {
      int h = vert_he[v];
      do {
        if (he[h].twin < 0) { vert_he[v] = h; break; }
        h = he[he[h].twin].next;
      } while (h != vert_he[v]);
    }
    // virtual faces along boundary loops (the halfedge list grows while it is being walked)
    for (size_t h0 = 0; h0 < he.size(); h0++) {
      if (he[h0].twin >= 0) continue;
      int b = (int)face_boundary.size(); face_boundary.push_back(1);
      std::vector<int> bh;
      int i = (int)h0;
      do {
        int t = (int)he.size(); he.push_back(HE());
        bh.push_back(t);
        he[i].twin = t; he[t].twin = i; he[t].face = b;
        he[t].vertex = he[he[i].next].vertex;
        i = he[i].next;
        while (i != (int)h0 && he[i].twin >= 0) { i = he[i].twin; i = he[i].next; }
      } while (i != (int)h0);
      size_t deg = bh.size();
      for (size_t p = 0; p < deg; p++) he[bh[p]].next = bh[(p + deg - 1) % deg];
    }
    for (size_t v = 0; v < vert_he.size(); v++) vert_he[v] = he[he[vert_he[v]].twin].next;
    for (size_t v = 0; v < vert_he.size(); v++) {
      size_t count = 0; int h = vert_he[v];
      do { if (!he_is_boundary(h)) count++; h = he[he[h].twin].next; } while (h != vert_he[v]);
      if (count != degree[(int)v]) throw std::runtime_error("at least one of the vertices is nonmanifold");
    }
    if (positions.size() != vert_he.size()) throw std::runtime_error("number of vertex positions differs from the number of distinct vertices");
    pos.assign(vert_he.size(), V3()); nrm.assign(vert_he.size(), V3());
    size_t k = 0;
    for (const auto& e : index_to_vertex) pos[e.second] = positions[k++];
    for (size_t v = 0; v < vert_he.size(); v++) compute_normal((int)v);
  }

  bool vertex_is_boundary(int v) const {  // halfEdgeMesh.h:532-546
    int h = vert_he[v];
    do { if (he_is_boundary(h)) return true; h = he[he[h].twin].next; } while (h != vert_he[v]);
    return false;
  }
  void compute_normal(int v) {  // halfEdgeMesh.h:492-515
    V3 n(0., 0., 0.), pi = pos[v];
    int h = vert_he[v];
    const bool boundary = vertex_is_boundary(v);
    do {
      V3 pj = pos[he[he[h].next].vertex];
      V3 pk = pos[he[he[he[h].next].next].vertex];
      n = n + cross(pj - pi, pk - pi);
      h = boundary ? he[he[h].next].twin : he[he[h].twin].next;
    } while (h != vert_he[v]);
    nrm[v] = n * (1. / norm(n));
  }
};

}  // namespace

bool load_dae(const std::string& path, FlatSceneData* out, std::string* err) {
  try {
    Loader L;
    L.load_file(path);
    out->materials.clear(); out->lights.clear(); out->prims.clear();
    memset(&out->camera, 0, sizeof out->camera);

    // Application::load (application.cpp:228-304)
    const int screenW = 800, screenH = 600;
    V3 c_dir(0, 0, 0);
    bool have_camera = false;
    double hFov = 50, vFov = 35, nClip = 0.01, fClip = 100, screenDist = 0;
    V3 bmin(INF_D, INF_D, INF_D), bmax(-INF_D, -INF_D, -INF_D);
    auto expand = [&](V3 p) {
      bmin = V3(std::min(bmin.x, p.x), std::min(bmin.y, p.y), std::min(bmin.z, p.z));
      bmax = V3(std::max(bmax.x, p.x), std::max(bmax.y, p.y), std::max(bmax.z, p.z));
    };
    {  // Application::init: the dummy camera (application.cpp:97-103)
      double ar1 = std::tan((hFov * (PI / 180)) / 2) / std::tan((vFov * (PI / 180)) / 2), ar = (double)screenW / screenH;
      if (ar1 < ar) hFov = 2 * (std::atan(std::tan((vFov * (PI / 180)) / 2) * ar) * (180 / PI));
      else if (ar1 > ar) vFov = 2 * (std::atan(std::tan((hFov * (PI / 180)) / 2) / ar) * (180 / PI));
      screenDist = ((double)screenH) / (2.0 * std::tan((vFov * (PI / 180)) / 2));
    }
    for (NodeDesc& n : L.nodes) {
      const M4& T = n.transform;
      switch (n.kind) {
        case Kind::CAMERA: {
          c_dir = unit(to3D(T.mul(v4(V3(0, 0, -1), 1))));
          // Camera::configure (camera.cpp:29-47)
          nClip = n.nClip; fClip = n.fClip; hFov = n.hFov; vFov = n.vFov;
          double ar1 = std::tan((hFov * (PI / 180)) / 2) / std::tan((vFov * (PI / 180)) / 2), ar = (double)screenW / screenH;
          if (ar1 < ar) hFov = 2 * (std::atan(std::tan((vFov * (PI / 180)) / 2) * ar) * (180 / PI));
          else if (ar1 > ar) vFov = 2 * (std::atan(std::tan((hFov * (PI / 180)) / 2) / ar) * (180 / PI));
          screenDist = ((double)screenH) / (2.0 * std::tan((vFov * (PI / 180)) / 2));
          have_camera = true;
          break;
        }
        case Kind::LIGHT: {
          BdptLight l; memset(&l, 0, sizeof l);
          Loader::set3(l.radiance, n.spectrum);
          const V3 lpos(0, 0, 0), ldir(0, 0, -1), lup(0, 1, 0);  // LightInfo defaults, light_info.cpp:7-24
          if (n.light == LightKind::AREA) {  // gl_scene/area_light.h:17-33
            l.type = BDPT_LIGHT_AREA;
            V3 position = to3D(T.mul(v4(lpos, 1)));
            V3 direction = to3D(T.mul(v4(ldir, 1))) - position;
            direction = direction * (1. / norm(direction));
            V3 dim_y = lup, dim_x = cross(lup, ldir);
            Loader::set3(l.position, position); Loader::set3(l.direction, direction);
            Loader::set3(l.dim_x, to3D(T.mul(v4(dim_x, 1))) - position);
            Loader::set3(l.dim_y, to3D(T.mul(v4(dim_y, 1))) - position);
          } else if (n.light == LightKind::POINT) {  // gl_scene/point_light.h:17-22
            l.type = BDPT_LIGHT_POINT;
            Loader::set3(l.position, to3D(T.mul(v4(lpos, 1))));
          } else {
            throw std::runtime_error("light type not supported by the BDPT integrator (only area and point lights implement sample_Le; "
                                     "the reference aborts on the others, src/scene/light.cpp)");
          }
          out->lights.push_back(l);
          break;
        }
        case Kind::SPHERE: {  // application.cpp:344-348, gl_scene/sphere.cpp:12-20
          V3 position = project3D(T.mul(V4{0, 0, 0, 1}));
          double scale = norm(to3D(T.mul(V4{1, 0, 0, 0})));
          double r = (double)n.radius * scale;
          BdptMaterial m;
          if (n.has_material) m = n.material; else { memset(&m, 0, sizeof m); m.type = BDPT_MAT_DIFFUSE; m.reflectance[0] = m.reflectance[1] = m.reflectance[2] = (double)0.5f; }
          int mi = (int)out->materials.size(); out->materials.push_back(m);
          BdptPrim p; memset(&p, 0, sizeof p);
          p.type = BDPT_PRIM_SPHERE; p.material = mi;
          p.data[0] = position.x; p.data[1] = position.y; p.data[2] = position.z; p.data[3] = r;
          out->prims.push_back(p);
          expand(V3(position.x - r, position.y - r, position.z - r)); expand(V3(position.x + r, position.y + r, position.z + r));
          break;
        }
        case Kind::POLYMESH: {  // gl_scene/mesh.cpp:22-43, object.cpp:16-58
          std::vector<V3> verts = n.vertices;
          for (V3& v : verts) v = project3D(T.mul(v4(v, 1)));
          HalfedgeMesh hm;
          hm.build(n.polygons, verts);
          for (const V3& p : hm.pos) expand(p);
          if (hm.n_real_faces == 0) break;
          BdptMaterial m;
          if (n.has_material) m = n.material; else { memset(&m, 0, sizeof m); m.type = BDPT_MAT_DIFFUSE; m.reflectance[0] = m.reflectance[1] = m.reflectance[2] = (double)0.5f; }
          int mi = (int)out->materials.size(); out->materials.push_back(m);
          for (size_t f = 0; f < hm.n_real_faces; f++) {
            int h = hm.face_he[f];
            int v[3] = {hm.he[h].vertex, hm.he[hm.he[h].next].vertex, hm.he[hm.he[hm.he[h].next].next].vertex};
            BdptPrim p; memset(&p, 0, sizeof p);
            p.type = BDPT_PRIM_TRIANGLE; p.material = mi;
            for (int k = 0; k < 3; k++) {
              p.data[3 * k + 0] = hm.pos[v[k]].x; p.data[3 * k + 1] = hm.pos[v[k]].y; p.data[3 * k + 2] = hm.pos[v[k]].z;
              p.data[9 + 3 * k + 0] = hm.nrm[v[k]].x; p.data[9 + 3 * k + 1] = hm.nrm[v[k]].y; p.data[9 + 3 * k + 2] = hm.nrm[v[k]].z;
            }
            out->prims.push_back(p);
          }
          break;
        }
        default: break;
      }
    }
    // camera placement (application.cpp:272-296, camera.cpp:53-65,123-147)
    BdptCamera& cam = out->camera;
    cam.hfov_deg = hFov; cam.vfov_deg = vFov; cam.nclip = nClip; cam.fclip = fClip;
    cam.screen_dist = screenDist; cam.screen_w = screenW; cam.screen_h = screenH;
    if (bmin.x <= bmax.x && have_camera) {
      V3 target = div(bmin + bmax, 2);
      V3 extent = bmax - bmin;
      double canonical = norm(extent) / 2 * 1.5;
      double r = canonical * 2, minR = canonical / 10.0, maxR = canonical * 20.0;
      double phi = std::acos(c_dir.y), theta = std::atan2(c_dir.x, c_dir.z);
      r = std::min(std::max(r, minR), maxR);
      if (std::sin(phi) == 0) phi = phi + EPS_F;
      double sinPhi = std::sin(phi);
      if (sinPhi == 0) { phi += EPS_F; sinPhi = std::sin(phi); }
      V3 dirToCamera(r * sinPhi * std::sin(theta), r * std::cos(phi), r * sinPhi * std::cos(theta));
      V3 pos = target + dirToCamera;
      V3 upVec(0, sinPhi > 0 ? 1 : -1, 0);
      V3 sx = cross(upVec, dirToCamera); sx = sx * (1. / norm(sx));
      V3 sy = cross(dirToCamera, sx); sy = sy * (1. / norm(sy));
      V3 sz = unit(dirToCamera);
      Loader::set3(cam.pos, pos); Loader::set3(cam.target, target);
      Loader::set3(cam.c2w + 0, sx); Loader::set3(cam.c2w + 3, sy); Loader::set3(cam.c2w + 6, sz);
    } else {
      throw std::runtime_error(have_camera ? "scene has no geometry" : "scene has no camera");
    }
    return true;
  } catch (const std::exception& e) {
    if (err) *err = e.what();
    return false;
  }
}

void apply_screen_size(BdptCamera* cam, int w, int h) {  // Camera::set_screen_size, camera.cpp:83-89
  cam->screen_w = w; cam->screen_h = h;
  cam->hfov_deg = 2 * (std::atan(((double)w) / (2 * cam->screen_dist)) * (180 / PI));
  cam->vfov_deg = 2 * (std::atan(((double)h) / (2 * cam->screen_dist)) * (180 / PI));
}

bool apply_camera_settings(BdptCamera* cam, const std::string& path, int w, int h, std::string* err) {
  FILE* f = fopen(path.c_str(), "r");
  if (!f) { if (err) *err = "cannot open camera settings file " + path; return false; }
  double v[30];
  int n = 0;
  while (n < 30 && fscanf(f, "%lf", &v[n]) == 1) n++;
  fclose(f);
  if (n < 28) { if (err) *err = "camera settings file " + path + " is too short"; return false; }
  // hFov vFov ar nClip fClip | pos[3] target[3] | phi theta r minR maxR | c2w(row, col) x 9 | screenW screenH screenDist | focalDistance lensRadius
  cam->hfov_deg = v[0]; cam->vfov_deg = v[1]; cam->nclip = v[3]; cam->fclip = v[4];
  for (int i = 0; i < 3; i++) { cam->pos[i] = v[5 + i]; cam->target[i] = v[8 + i]; }
  for (int i = 0; i < 9; i++) cam->c2w[3 * (i % 3) + i / 3] = v[16 + i];   // file: row-major; BdptCamera: three column vectors
  cam->screen_dist = v[27];
  // the frame keeps its size; with screen_w/h equal to it the library does not re-derive the fovs (camera.cpp:83-89)
  cam->screen_w = w > 0 ? w : v[25];
  cam->screen_h = h > 0 ? h : v[26];
  return true;
}

std::vector<unsigned char> serialize(const FlatSceneData& s) {
  BdptSceneHeader h; memset(&h, 0, sizeof h);
  h.magic = BDPT_FSCENE_MAGIC; h.n_prims = (uint32_t)s.prims.size(); h.n_materials = (uint32_t)s.materials.size(); h.n_lights = (uint32_t)s.lights.size();
  std::vector<unsigned char> out;
  auto put = [&](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; out.insert(out.end(), b, b + n); };
  put(&h, sizeof h); put(&s.camera, sizeof s.camera);
  put(s.materials.data(), s.materials.size() * sizeof(BdptMaterial));
  put(s.lights.data(), s.lights.size() * sizeof(BdptLight));
  put(s.prims.data(), s.prims.size() * sizeof(BdptPrim));
  return out;
}

}  // namespace bdpt_host
