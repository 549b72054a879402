// flyscene.hpp -- C++ host facade with the reference's own entry points over the C ABI.
//
// The reference has no plugin/FFI layer: its render path is the public surface of four classes
// (SURVEY.md 8b).  This header keeps those class and method names, argument meaning and error
// behaviour, and forwards every compute call through include/rt_api.h into the CUDA kernels --
// there is no CPU implementation of the ray tracer behind it.
//
//   reference                                   here
//   ------------------------------------------  -----------------------------------------------------
//   class BoundingBox   src/boundingBox.hpp:21  rt::BoundingBox   (boxIntersect -> rt_box_intersect_box)
//   class BoxTree       src/boxTree.hpp:15      rt::BoxTree       (ctor -> rt_scene_create, intersect -> rt_octree_candidates)
//   class arealight     arealight.hpp:5         rt::arealight     (host-side sample grid, same expression)
//   class Flyscene      src/flyscene.hpp:29     rt::Flyscene      (raytraceScene -> rt_render + rt_write_ppm, ...)
//   Tucano::Flycamera   tucano/utils/flycamera.hpp   rt::Flycamera (the state raytraceScene reads)
//
// Vector3f here is rt::Vec3f (three packed floats, layout-compatible with Eigen::Vector3f); a
// maintainer who keeps Eigen passes `v.data()` to the same C functions.
#pragma once

#include <cstdint>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "../../include/rt_api.h"
#include "vec3.hpp"

namespace rt {

using Vector3f = Vec3f;

class Flyscene;

// ---------------------------------------------------------------------------------------------
// src/boundingBox.hpp:21-43
// ---------------------------------------------------------------------------------------------
class BoundingBox {
 public:
  BoundingBox() {}
  BoundingBox(const Vector3f &minv, const Vector3f &maxv) : vmin(minv), vmax(maxv) {}
  // slab test of the infinite forward ray origin -> dest (reference semantics incl. inf/NaN)
  bool boxIntersect(const Vector3f &origin, const Vector3f &dest);
  Vector3f getMin() { return vmin; }
  Vector3f getMax() { return vmax; }
  void setMin(Vector3f mn) { vmin = mn; }
  void setMax(Vector3f mx) { vmax = mx; }

 private:
  Vector3f vmin, vmax;
};

// ---------------------------------------------------------------------------------------------
// src/boxTree.hpp:15-62.  The acceleration structure lives on the GPU as a flattened BVH plus the
// reference-octree candidate filter; `box`, `capacity`, `isLeaf`, `isEmpty` keep their meaning for
// the root, `faces` lists the root's faces when it is a leaf.  `children` stays empty: the octants
// are not materialised on the host (their count and shape are reported by stats()).
// ---------------------------------------------------------------------------------------------
class BoxTree {
 public:
  BoundingBox box;
  int capacity = 1000;
  bool isLeaf = false;
  bool isEmpty = false;
  std::vector<BoxTree> children;
  std::vector<int> faces;

  BoxTree() {}
  // builds BVH + candidate filter from a baked scene description and uploads it (rt_scene_create)
  BoxTree(const RtSceneDesc &scene, int capacity);
  // the same view over a scene that somebody else owns (device 0 of a multi-GPU scene, rt_multi_scene)
  BoxTree(const RtSceneDesc &scene, int capacity, RtScene *borrowed);
  ~BoxTree();
  BoxTree(BoxTree &&o) noexcept;
  BoxTree &operator=(BoxTree &&o) noexcept;
  BoxTree(const BoxTree &) = delete;
  BoxTree &operator=(const BoxTree &) = delete;

  // candidate face ids for the query origin -> dest, exactly the reference's std::set
  std::set<int> intersect(const Vector3f &origin, const Vector3f &dest);
  // {reachable leaves, inner nodes, face references, largest leaf} of the reference octree
  std::vector<int64_t> stats() const { return oct_stats; }
  RtScene *handle() const { return scene_; }

 private:
  void describe(const RtSceneDesc &scene, int cap);
  RtScene *scene_ = nullptr;
  bool owns_ = true;
  std::vector<int64_t> oct_stats;
};

// ---------------------------------------------------------------------------------------------
// arealight.hpp:5-37
// ---------------------------------------------------------------------------------------------
class arealight {
 public:
  arealight(Vector3f &corner, Vector3f &uvec, int usteps, Vector3f &vvec, int vsteps)
      : corner(corner), uvec(uvec), usteps(usteps), vvec(vvec), vsteps(vsteps) {}
  std::vector<Vector3f> getPointLights();
  Vector3f corner, uvec;
  int usteps;
  Vector3f vvec;
  int vsteps;
};

// ---------------------------------------------------------------------------------------------
// The camera state the render path reads (tucano/camera.hpp:115-173, flycamera.hpp:76-191).
// ---------------------------------------------------------------------------------------------
class Flycamera {
 public:
  Flycamera() { reset(); }
  void reset();
  void setPerspectiveMatrix(float fy, float aspect, float near_plane, float far_plane);
  void setViewport(float w, float h);
  void translate(float dx, float dy, float dz);  // camera-space translation vector (flycamera.hpp translation_vector)
  void setRotation(float rot_x_axis, float rot_y_axis);
  Vector3f getCenter() const;
  Vector3f screenToWorld(float i, float j) const;  // rt_screen_to_world
  int viewportWidth() const { return (int)viewport[2]; }
  int viewportHeight() const { return (int)viewport[3]; }
  RtCamera abi() const;  // what crosses the C ABI

 private:
  void updateViewMatrix();
  float view[12];  // 3x4 row-major affine
  float viewport[4];
  float fovy = 60.f, aspect = 1.f;
  float rot_x = 0.f, rot_y = 0.f;
  Vector3f translation, default_translation;
};

// ---------------------------------------------------------------------------------------------
// src/flyscene.hpp:29-200 (render-path members only; the GL preview / debug gizmos are out of scope)
// ---------------------------------------------------------------------------------------------
struct Face {  // Tucano::Face as the ray tracer reads it
  int id = -1;
  int material_id = -1;
  Vector3f normal;
};

// One level of the debug ray (the fields of the reference's "DEBUG RAY INFO" console dump,
// src/flyscene.cpp:364-407)
struct DebugRayLevel {
  int level = 0;
  int face = -1;            // -1: this segment left the scene
  float t = 0.f;
  Vector3f origin, direction, hit, normal, reflected, colour;
  float shininess = 0.f;
  std::vector<bool> light_visible;          // lightStrikes per light
  std::vector<float> cos_theta, cos_phi;    // per light, as printed by the reference
};

class Flyscene {
 public:
  Flyscene() {}
  ~Flyscene();

  // reference: prompts on stdin for the two light-mode flags, loads resources/models/cube.obj,
  // normalises it, adds the first light (-1,1,1), builds the acceleration structure (src/flyscene.cpp:29-126)
  void initialize(int width, int height);
  // non-interactive variants of the same set-up
  void setModelPath(const std::string &obj) { model_path = obj; }
  void setLightMode(bool area, bool point) { areaLight = area; pointLight = point; mode_set = true; }
  void setMaxDepth(int d) { max_depth = d; }
  void setAreaGrid(int u, int v) { usteps = u; vsteps = v; }
  void setSphereSeed(uint32_t s) { sphere_seed = s; }  // spherical light mode: RtParams.sphere_seed
  // GPUs of this box that raytraceScene() spreads the image over (interleaved row bands, rt_multi_*): the role
  // of the reference's ThreadPool workers (src/flyscene.cpp:558,609).  Call before initialize(); default: the
  // device chosen by rt_init.
  void setDevices(const std::vector<int> &d) { devices = d; }

  Flycamera *getCamera() { return &flycamera; }
  void addLight() { lights.push_back(flycamera.getCenter()); }
  std::vector<Vector3f> &getLights() { return lights; }
  int getNumberOfFaces() const;
  Face getFace(int i) const;

  // renders the current view and writes result.ppm (P3) into the CWD (src/flyscene.cpp:519-648)
  void raytraceScene(int width = 0, int height = 0);
  // same, keeping the packed RGBA frame in memory; stats may be null
  const std::vector<uint8_t> &render(int width, int height, RtStats *stats = nullptr);

  Vector3f traceRay(Vector3f &origin, Vector3f &direction, int level, std::vector<Vector3f> &lights, bool countRay);
  float rayTriangleIntersection(Vector3f &rayPoint, Vector3f &rayDirection, Face &triangle);
  Vector3f phongShade(Vector3f &origin, Vector3f &hitPoint, Face &triangle, std::vector<Vector3f> &lights);
  bool lightStrikes(Vector3f &hitPoint, std::vector<Vector3f> &lights, bool visibleLights[]);
  // Headless counterpart of recursiveDebugRay (src/flyscene.cpp:241-430): follows the ray through pixel
  // (px, py) and its mirror reflections for up to maxDepth bounces with single-ray GPU queries and returns
  // what the reference prints per level (its cylinder / sphere gizmos are GL and out of scope).  Unlike the
  // reference -- which re-tests every level against the camera->pixel candidate set -- each level is a proper
  // traceRay from the previous hit point.  print = true writes the reference's console layout to stdout.
  std::vector<DebugRayLevel> debugRay(float px, float py, int maxDepth, bool print = false);
  arealight createAreaLight(Vector3f corner, float lengthX, float lengthY, int usteps, int vsteps);
  std::vector<Vector3f> createSpherePoint(Vector3f lightPoint);

  BoxTree octree;          // src/flyscene.hpp:176
  double octree_seconds = 0.0, render_seconds = 0.0;

 private:
  RtParams params(int w, int h) const;
  Flycamera flycamera;
  std::vector<Vector3f> lights;
  RtMesh *mesh = nullptr;
  RtMulti *multi = nullptr;  // more than one device: the scenes of all devices (octree views device 0's)
  std::vector<int> devices;
  RtSceneDesc desc{};
  std::string model_path = "resources/models/cube.obj";  // src/flyscene.cpp:51
  bool areaLight = false, pointLight = true, mode_set = false;
  int max_depth = -1, usteps = 5, vsteps = 5;
  uint32_t sphere_seed = 1;
  float light_color[3] = {1.f, 1.f, 0.f};  // lightrep.setColor, src/flyscene.cpp:68
  std::vector<uint8_t> frame;
};

}  // namespace rt
