// flyscene.cpp -- see flyscene.hpp.  Every ray-tracing call below crosses the C ABI into CUDA.
#include "flyscene.hpp"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <stdexcept>

namespace rt {

namespace {
void check(int rc, const char *what) {
  if (rc < 0) throw std::runtime_error(std::string(what) + ": " + rt_last_error());
}
}  // namespace

// ---------------------------------------------------------------------------------------------
// BoundingBox
// ---------------------------------------------------------------------------------------------
bool BoundingBox::boxIntersect(const Vector3f &origin, const Vector3f &dest) {
  uint8_t hit = 0;
  check(rt_box_intersect_box(vmin.data(), vmax.data(), 1, origin.data(), dest.data(), &hit), "BoundingBox::boxIntersect");
  return hit != 0;
}

// ---------------------------------------------------------------------------------------------
// BoxTree
// ---------------------------------------------------------------------------------------------
BoxTree::BoxTree(const RtSceneDesc &scene, int cap) {
  check(rt_scene_create(&scene, &scene_), "BoxTree::BoxTree");
  describe(scene, cap);
}

BoxTree::BoxTree(const RtSceneDesc &scene, int cap, RtScene *borrowed) : scene_(borrowed), owns_(false) { describe(scene, cap); }

void BoxTree::describe(const RtSceneDesc &scene, int cap) {
  capacity = cap;
  float mn[3], mx[3];
  rt_scene_root_box(scene_, mn, mx);
  box = BoundingBox(Vector3f(mn), Vector3f(mx));
  oct_stats.assign(4, 0);
  rt_ref_octree_stats(&scene, cap, oct_stats.data());
  // root flags as the reference sets them (src/boxTree.cpp:23-30)
  isEmpty = scene.n_faces == 0;
  isLeaf = !isEmpty && scene.n_faces <= cap;
  if (isLeaf)
    for (int i = 0; i < scene.n_faces; ++i) faces.push_back(i);
}

BoxTree::~BoxTree() {
  if (scene_ && owns_) rt_scene_destroy(scene_);
}

BoxTree::BoxTree(BoxTree &&o) noexcept { *this = std::move(o); }

BoxTree &BoxTree::operator=(BoxTree &&o) noexcept {
  if (this != &o) {
    if (scene_ && owns_) rt_scene_destroy(scene_);
    owns_ = o.owns_;
    box = o.box; capacity = o.capacity; isLeaf = o.isLeaf; isEmpty = o.isEmpty;
    faces = std::move(o.faces);
    oct_stats = std::move(o.oct_stats);
    scene_ = o.scene_;
    o.scene_ = nullptr;
  }
  return *this;
}

std::set<int> BoxTree::intersect(const Vector3f &origin, const Vector3f &dest) {
  std::set<int> out;
  if (!scene_) return out;
  int64_t n_tris = 0;
  rt_scene_info(scene_, nullptr, nullptr, &n_tris, nullptr, nullptr);
  std::vector<int32_t> ids((size_t)std::max<int64_t>(1, n_tris));
  int n = rt_octree_candidates(scene_, origin.data(), dest.data(), ids.data(), (int32_t)ids.size());
  check(n, "BoxTree::intersect");
  out.insert(ids.begin(), ids.begin() + std::min<size_t>((size_t)n, ids.size()));
  return out;
}

// ---------------------------------------------------------------------------------------------
// arealight::getPointLights, arealight.hpp:15-25: the grid is anchored at the origin, not at the
// light: sample(i,j) = ((i+.5)*uvec.x/usteps, (j+.5)*vvec.y/vsteps, uvec.z)
// ---------------------------------------------------------------------------------------------
std::vector<Vector3f> arealight::getPointLights() {
  std::vector<Vector3f> out;
  for (int i = 0; i < usteps; ++i)
    for (int j = 0; j < vsteps; ++j)
      out.push_back(Vector3f((float)(i + 0.5) * (uvec.x / (float)usteps), (float)(j + 0.5) * (vvec.y / (float)vsteps), uvec.z));
  return out;
}

// ---------------------------------------------------------------------------------------------
// Flycamera
// ---------------------------------------------------------------------------------------------
namespace {
Vector3f rotate_about(Vector3f axis, float angle, Vector3f v) {  // Rodrigues
  const float c = std::cos(angle), s = std::sin(angle);
  return c * v + s * cross(axis, v) + ((1.f - c) * dot(axis, v)) * axis;
}
void inverse3(const float m[9], float out[9]) {  // cofactor inverse, like Eigen's fixed-size 3x3 path
  const float c00 = m[4] * m[8] - m[5] * m[7], c10 = m[5] * m[6] - m[3] * m[8], c20 = m[3] * m[7] - m[4] * m[6];
  const float det = m[0] * c00 + m[1] * c10 + m[2] * c20;
  const float id = 1.f / det;
  out[0] = c00 * id; out[1] = (m[2] * m[7] - m[1] * m[8]) * id; out[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  out[3] = c10 * id; out[4] = (m[0] * m[8] - m[2] * m[6]) * id; out[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  out[6] = c20 * id; out[7] = (m[1] * m[6] - m[0] * m[7]) * id; out[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}
}  // namespace

void Flycamera::reset() {
  translation = Vector3f(0, 0, 0);
  default_translation = Vector3f(0.f, 0.f, -2.f);  // flycamera.hpp:81
  rot_x = rot_y = 0.f;
  viewport[0] = viewport[1] = 0.f;
  viewport[2] = viewport[3] = 1.f;
  updateViewMatrix();
}

void Flycamera::setPerspectiveMatrix(float fy, float a, float, float) { fovy = fy; aspect = a; }
void Flycamera::setViewport(float w, float h) { viewport[0] = 0.f; viewport[1] = 0.f; viewport[2] = w; viewport[3] = h; }
void Flycamera::translate(float dx, float dy, float dz) { translation = translation + Vector3f(dx, dy, dz); updateViewMatrix(); }
void Flycamera::setRotation(float rx, float ry) { rot_x = rx; rot_y = ry; updateViewMatrix(); }

void Flycamera::updateViewMatrix() {
  // flycamera.hpp:166-191: rows of the rotation are the yawed/pitched axes; view = R * T(default) * T(translation)
  const Vector3f ux(1, 0, 0), uy(0, 1, 0), uz(0, 0, 1);
  const Vector3f rx = normalized(rotate_about(uy, rot_y, ux));
  const Vector3f rz = normalized(rotate_about(rx, rot_x, rotate_about(uy, rot_y, uz)));
  const Vector3f ry = normalized(rotate_about(rx, rot_x, uy));
  const float R[9] = {rx.x, rx.y, rx.z, ry.x, ry.y, ry.z, rz.x, rz.y, rz.z};
  const Vector3f t = default_translation + translation;
  for (int r = 0; r < 3; ++r) {
    view[4 * r] = R[3 * r]; view[4 * r + 1] = R[3 * r + 1]; view[4 * r + 2] = R[3 * r + 2];
    view[4 * r + 3] = R[3 * r] * t.x + R[3 * r + 1] * t.y + R[3 * r + 2] * t.z;
  }
}

Vector3f Flycamera::getCenter() const {
  const RtCamera c = abi();
  return Vector3f(c.eye);
}

RtCamera Flycamera::abi() const {
  RtCamera c{};
  const float L[9] = {view[0], view[1], view[2], view[4], view[5], view[6], view[8], view[9], view[10]};
  float Li[9];
  inverse3(L, Li);
  const float tr[3] = {view[3], view[7], view[11]};
  for (int r = 0; r < 3; ++r) {
    const float it = -(Li[3 * r] * tr[0] + Li[3 * r + 1] * tr[1] + Li[3 * r + 2] * tr[2]);
    c.view_inv[4 * r] = Li[3 * r]; c.view_inv[4 * r + 1] = Li[3 * r + 1]; c.view_inv[4 * r + 2] = Li[3 * r + 2];
    c.view_inv[4 * r + 3] = it;
    c.eye[r] = it;  // linear^-1 * (-translation), camera.hpp:115-118
  }
  memcpy(c.viewport, viewport, sizeof(viewport));
  c.fovy = fovy;
  c.aspect = aspect;
  return c;
}

Vector3f Flycamera::screenToWorld(float i, float j) const {
  const RtCamera c = abi();
  const float px[2] = {i, j};
  float out[3];
  check(rt_screen_to_world(&c, 1, px, out), "Flycamera::screenToWorld");
  return Vector3f(out);
}

// ---------------------------------------------------------------------------------------------
// Flyscene
// ---------------------------------------------------------------------------------------------
Flyscene::~Flyscene() {
  octree = BoxTree();  // drop the (possibly borrowed) view before the scenes it points into
  if (multi) rt_multi_destroy(multi);
  if (mesh) rt_mesh_destroy(mesh);
}

void Flyscene::initialize(int width, int height) {
  if (!mode_set) {
    // the reference's two prompts (src/flyscene.cpp:31-34)
    int a = 0, p = 1;
    std::cout << "Enter 0 if Point Lights or 1 if Area Lights : " << std::endl;
    std::cin >> a;
    std::cout << "Enter 0 if spherical or 1 if point : " << std::endl;
    std::cin >> p;
    areaLight = a != 0;
    pointLight = p != 0;
  }
  flycamera.setPerspectiveMatrix(60.0f, width / (float)height, 0.1f, 100.0f);  // :46
  flycamera.setViewport((float)width, (float)height);                           // :47
  check(rt_mesh_load_obj(model_path.c_str(), &mesh), "loadObjFile");
  check(rt_mesh_desc(mesh, &desc), "rt_mesh_desc");
  std::cout << "OBJ info:\nnumber faces : " << desc.n_faces << "\nnumber materials : " << desc.n_materials << std::endl;
  lights.push_back(Vector3f(-1.0f, 1.0f, 1.0f));  // :72
  std::cout << "Seting up acceleration data structure ..." << std::endl;
  const auto t0 = std::chrono::high_resolution_clock::now();
  if (devices.size() > 1) {
    // one scene per GPU, baked once; `octree` views the first device's copy
    check(rt_multi_create(&desc, (int)devices.size(), devices.data(), &multi), "rt_multi_create");
    RtScene *first = nullptr;
    check(rt_multi_scene(multi, 0, &first), "rt_multi_scene");
    octree = BoxTree(desc, 1000, first);
  } else {
    if (devices.size() == 1) check(rt_init(devices[0]), "rt_init");
    octree = BoxTree(desc, 1000);  // :86,93
  }
  octree_seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
  std::cout << "Seting up acceleration data structure: done!\nELAPSED TIME:" << octree_seconds << std::endl;
}

int Flyscene::getNumberOfFaces() const { return desc.n_faces; }

Face Flyscene::getFace(int i) const {
  Face f;
  if (i < 0 || i >= desc.n_faces) return f;
  f.id = i;
  f.material_id = desc.material_id[i];
  f.normal = Vector3f(desc.face_normals + 3 * (size_t)i);
  return f;
}

RtParams Flyscene::params(int w, int h) const {
  RtParams p;
  rt_default_params(&p);
  p.width = w; p.height = h;
  p.area_light = areaLight; p.point_light = pointLight;
  p.max_depth = max_depth;
  p.usteps = usteps; p.vsteps = vsteps;
  p.sphere_seed = sphere_seed;
  return p;
}

const std::vector<uint8_t> &Flyscene::render(int width, int height, RtStats *stats) {
  if (width == 0 || height == 0) { width = flycamera.viewportWidth(); height = flycamera.viewportHeight(); }  // :530-533
  const RtCamera cam = flycamera.abi();
  std::vector<float> lp;
  for (const Vector3f &l : lights) { lp.push_back(l.x); lp.push_back(l.y); lp.push_back(l.z); }
  RtLights L{};
  L.n = (int32_t)lights.size();
  L.pos = lp.data();
  memcpy(L.color, light_color, sizeof(light_color));
  const RtParams p = params(width, height);
  frame.assign((size_t)width * height * 4, 0);
  if (multi) check(rt_multi_render(multi, &cam, &L, &p, frame.data(), stats), "raytraceScene");
  else check(rt_render(octree.handle(), &cam, &L, &p, frame.data(), nullptr, nullptr, nullptr, stats), "raytraceScene");
  return frame;
}

void Flyscene::raytraceScene(int width, int height) {
  const auto start = std::chrono::high_resolution_clock::now();
  if (width == 0 || height == 0) { width = flycamera.viewportWidth(); height = flycamera.viewportHeight(); }
  std::cout << "Ray tracing ..." << std::endl;
  render(width, height, nullptr);
  std::cout << "Writting to restult.ppm ... " << std::endl;  // (sic) the reference's own message, :639
  check(rt_write_ppm("result.ppm", frame.data(), width, height, 0), "writePPMImage");
  render_seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - start).count();
  std::cout << "ray tracing done! " << std::endl;
  std::cout << "ELAPSED TIME:" << render_seconds << std::endl;
}

Vector3f Flyscene::traceRay(Vector3f &origin, Vector3f &direction, int level, std::vector<Vector3f> &lts, bool) {
  // `level` only matters together with a depth cap: a ray entering at level L may spawn max_depth-L bounces
  std::vector<float> lp;
  for (const Vector3f &l : lts) { lp.push_back(l.x); lp.push_back(l.y); lp.push_back(l.z); }
  RtLights L{};
  L.n = (int32_t)lts.size();
  L.pos = lp.data();
  memcpy(L.color, light_color, sizeof(light_color));
  RtParams p = params(1, 1);
  if (p.max_depth >= 0) p.max_depth = std::max(0, p.max_depth - level);
  float rgb[3];
  check(rt_trace_rays(octree.handle(), 1, origin.data(), direction.data(), &L, &p, rgb, nullptr, nullptr), "traceRay");
  return Vector3f(rgb);
}

float Flyscene::rayTriangleIntersection(Vector3f &rayPoint, Vector3f &rayDirection, Face &triangle) {
  float t = -72.f;
  const int32_t f = triangle.id;
  check(rt_ray_triangle(octree.handle(), 1, rayPoint.data(), rayDirection.data(), &f, &t), "rayTriangleIntersection");
  return t;
}

Vector3f Flyscene::phongShade(Vector3f &origin, Vector3f &hitPoint, Face &triangle, std::vector<Vector3f> &lts) {
  std::vector<float> lp;
  for (const Vector3f &l : lts) { lp.push_back(l.x); lp.push_back(l.y); lp.push_back(l.z); }
  RtLights L{};
  L.n = (int32_t)lts.size();
  L.pos = lp.data();
  memcpy(L.color, light_color, sizeof(light_color));
  const RtParams p = params(1, 1);
  const int32_t f = triangle.id;
  float rgb[3];
  check(rt_phong_shade(octree.handle(), 1, origin.data(), hitPoint.data(), &f, &L, &p, rgb), "phongShade");
  return Vector3f(rgb);
}

bool Flyscene::lightStrikes(Vector3f &hitPoint, std::vector<Vector3f> &lts, bool visibleLights[]) {
  if (lts.empty()) return false;
  std::vector<float> lp;
  for (const Vector3f &l : lts) { lp.push_back(l.x); lp.push_back(l.y); lp.push_back(l.z); }
  RtLights L{};
  L.n = (int32_t)lts.size();
  L.pos = lp.data();
  std::vector<uint8_t> vis(lts.size());
  check(rt_light_strikes(octree.handle(), 1, hitPoint.data(), &L, vis.data()), "lightStrikes");
  bool any = false;
  for (size_t l = 0; l < lts.size(); ++l) { visibleLights[l] = vis[l] != 0; any = any || visibleLights[l]; }
  return any;
}

std::vector<DebugRayLevel> Flyscene::debugRay(float px, float py, int maxDepth, bool print) {
  std::vector<DebugRayLevel> out;
  std::vector<float> lp;
  for (const Vector3f &l : lights) { lp.push_back(l.x); lp.push_back(l.y); lp.push_back(l.z); }
  RtLights L{};
  L.n = (int32_t)lights.size();
  L.pos = lp.data();
  memcpy(L.color, light_color, sizeof(light_color));
  const Vector3f eye = flycamera.getCenter();
  Vector3f pos = eye;
  Vector3f dir = flycamera.screenToWorld(px, py) - eye;  // src/flyscene.cpp:434-441
  for (int n = 0; n <= maxDepth; ++n) {
    DebugRayLevel lv;
    lv.level = n;
    lv.origin = pos;
    lv.direction = dir;
    RtParams p = params(1, 1);
    if (p.max_depth >= 0) p.max_depth = std::max(0, p.max_depth - n);
    float rgb[3];
    int32_t face = -1;
    float t = 0.f;
    check(rt_trace_rays(octree.handle(), 1, pos.data(), dir.data(), &L, &p, rgb, &face, &t), "debugRay");
    lv.colour = Vector3f(rgb);
    lv.face = face;
    lv.t = t;
    if (face < 0 || face >= desc.n_faces) { out.push_back(lv); break; }
    lv.hit = pos + t * dir;
    lv.normal = Vector3f(desc.face_normals + 3 * (size_t)face);
    lv.reflected = dir - (2.f * dot(dir, lv.normal)) * lv.normal;  // :349
    lv.shininess = desc.materials[desc.material_id[face]].ns;
    if (!lights.empty()) {
      std::vector<uint8_t> vis(lights.size());
      check(rt_light_strikes(octree.handle(), 1, lv.hit.data(), &L, vis.data()), "lightStrikes");
      for (size_t i = 0; i < lights.size(); ++i) {
        const Vector3f dl = normalized(lv.hit - lights[i]);                         // :384
        lv.light_visible.push_back(vis[i] != 0);
        lv.cos_theta.push_back(dot(dl, lv.normal));                                  // :386
        lv.cos_phi.push_back(dot(normalized(-1.f * (lv.hit - eye)), lv.reflected));  // :387
      }
    }
    out.push_back(lv);
    pos = lv.hit;
    dir = lv.reflected;
  }
  if (print) {
    for (const DebugRayLevel &lv : out) {
      if (lv.face < 0) break;
      std::cout << "\n-------------------------------------------------------------\n"
                << "                      DEBUG RAY INFO (level = " << lv.level << ")\n"
                << "                      ==============                       \n\n";
      auto pv = [](const char *k, const Vector3f &v) { std::cout << k << "(" << v.x << ", " << v.y << ", " << v.z << ")\n"; };
      pv(" Hitpoint = ", lv.hit);
      std::cout << " Distance = " << norm(lv.hit - eye) << "\n";
      pv(" Normal vector = ", lv.normal);
      pv(" Reflection vector = ", lv.reflected);
      pv(" Color rendered = ", lv.colour);
      std::cout << " shininess = " << lv.shininess << "\n\n LIGHTS INFO                    \n";
      for (size_t i = 0; i < lv.cos_theta.size(); ++i)
        std::cout << " ----light " << i << " ----\n visible = " << (int)lv.light_visible[i] << "\n cos (Theta) = " << lv.cos_theta[i]
                  << "\n cos (Phi)   = " << lv.cos_phi[i] << "\n";
      std::cout << "\n HIT TRIANGLE INFO                    \n";
      for (int k = 0; k < 3; ++k) pv(k == 0 ? "                   Vertex 1 = " : k == 1 ? "                   Vertex 2 = " : "                   Vertex 3 = ",
                                     Vector3f(desc.verts + 9 * (size_t)lv.face + 3 * k));
      std::cout << "-------------------------------------------------------------" << std::endl;
    }
  }
  return out;
}

arealight Flyscene::createAreaLight(Vector3f corner, float lengthX, float lengthY, int us, int vs) {
  Vector3f uvec = corner + lengthX * Vector3f(1, 0, 0);  // src/flyscene.cpp:957-958: points, not edge vectors
  Vector3f vvec = corner + lengthY * Vector3f(0, 1, 0);
  return arealight(corner, uvec, us, vvec, vs);
}

std::vector<Vector3f> Flyscene::createSpherePoint(Vector3f lightPoint) {
  if (pointLight) return {lightPoint};
  if (areaLight) return createAreaLight(lightPoint, 0.3f, 0.15f, usteps, vsteps).getPointLights();
  // spherical mode (src/flyscene.cpp:974-993) with the documented deterministic draws (RtParams.sphere_seed)
  const RtParams p = params(1, 1);
  float out[25 * 3];
  const int n = rt_light_samples(&p, lightPoint.data(), out);
  if (n < 0) throw std::runtime_error(rt_last_error());
  std::vector<Vector3f> pts;
  for (int k = 0; k < n; ++k) pts.push_back(Vector3f(out[3 * k], out[3 * k + 1], out[3 * k + 2]));
  return pts;
}

}  // namespace rt
