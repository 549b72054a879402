// obj_loader.hpp -- host-side OBJ/MTL load + "bake" for the B200 render path.
//
// Produces the flat per-face arrays the device path consumes (RtSceneDesc, include/rt_api.h),
// following the scene conventions of the reference so that the same file gives the same scene:
//   loader           dependencies/tucano/tucano/utils/objimporter.hpp:83-284, mtlIO.hpp:45-125
//   face normals     dependencies/tucano/tucano/mesh.hpp:441-468
//   vertex normals   objimporter.hpp:50-74 (accumulated by vertex id on top of the file's vn list)
//   normalisation    mesh.hpp:621-642, model.hpp:169-173 (scale 1/radius about the centroid)
//   material params  materials/mtl.hpp:16-116
// All citations relative to /root/reference.  This is a fresh implementation (the memory-mapped text is parsed in slices by all
// host cores, no iostreams in the vertex/face loops), not a copy of those files.
#pragma once

#include <cstdint>
#include <memory>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "../../include/rt_api.h"
#include "vec3.hpp"

namespace rt {

// std::vector whose resize() leaves new elements uninitialised: the loader sizes the per-face arrays once and fills
// them from all host cores; value-initialising ~100 MB first would touch every page from one thread.
template <class T>
struct default_init_allocator : std::allocator<T> {
  template <class U> struct rebind { using other = default_init_allocator<U>; };
  using std::allocator<T>::allocator;
  template <class U> void construct(U *p) { ::new (static_cast<void *>(p)) U; }
  template <class U, class... Args> void construct(U *p, Args &&...args) { ::new (static_cast<void *>(p)) U(std::forward<Args>(args)...); }
};
template <class T> using RawVector = std::vector<T, default_init_allocator<T>>;

struct BakedMesh {
  // per face (T entries)
  RawVector<float> verts;      // [T][3][3] world space
  RawVector<float> fnormals;   // [T][3]
  RawVector<float> vnormals;   // [T][3][3]
  RawVector<int32_t> mat_id;   // [T]
  RawVector<int32_t> vertex_ids;  // [T][3]
  std::vector<RtMaterial> materials;
  std::vector<std::string> material_names;
  // object-space data
  std::vector<float> obj_verts;  // [NV][3]
  Vec3f centroid;
  float radius = 1.f;
  float norm_scale = 1.f;
  int32_t n_faces() const { return (int32_t)mat_id.size(); }
  int32_t n_vertices() const { return (int32_t)(obj_verts.size() / 3); }
};

// Throws std::runtime_error when the OBJ cannot be opened (the reference exits the process,
// objimporter.hpp:103-106).  A missing MTL is reported on stderr and a default material is used
// (mtlIO.hpp:49-53,111-115).
BakedMesh load_obj(const std::string &obj_path, bool normalize = true);

// Parse one MTL file, appending to `out` (names to `names`).  Returns false if it cannot be opened.
bool load_mtl(const std::string &path, std::vector<RtMaterial> &out, std::vector<std::string> &names);

RtMaterial default_material();

}  // namespace rt
