// ref_octree.cpp -- see ref_octree.hpp.  Re-derivation of the reference BoxTree build
// (src/boxTree.cpp:11-31 constructor, :88-147 split, :203-336 clasifyFace, :338-456 helpers;
// root box src/boundingBox.cpp:14-43), parallel over octants, flattened to parent links + a
// face -> leaves CSR.  Float expressions follow the reference's evaluation order (vec3.hpp).
#include "ref_octree.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <future>
#include <memory>

#include "vec3.hpp"

namespace rt {

namespace {

struct ONode {
  Vec3f mn, mx;
  bool leaf = false, empty = false;
  std::vector<int32_t> faces;
  std::vector<std::unique_ptr<ONode>> kids;  // 0 or 8
};

struct Ctx {
  const float *verts;
  int capacity;
};

inline bool inside(const ONode &n, Vec3f v) {
  return n.mn.x <= v.x && n.mx.x >= v.x && n.mn.y <= v.y && n.mx.y >= v.y && n.mn.z <= v.z && n.mx.z >= v.z;
}

// One separating-axis probe: projections p0, p1 of two (normalised) vertices against radius rad.
inline bool axis_ok(float p0, float p1, float rad) {
  const float lo = std::min(p1, p0), hi = std::max(p1, p0);
  return !(lo > rad || hi < -rad);
}
// the reference's three probe shapes (X: y/z, Y: x/z, Z: x/y components)
inline bool probe_x(float a, float b, float fa, float fb, Vec3f u, Vec3f v, Vec3f h) {
  return axis_ok(a * u.y - b * u.z, a * v.y - b * v.z, fa * h.y + fb * h.z);
}
inline bool probe_y(float a, float b, float fa, float fb, Vec3f u, Vec3f v, Vec3f h) {
  return axis_ok(-a * u.x + b * u.z, -a * v.x + b * v.z, fa * h.x + fb * h.z);
}
inline bool probe_z(float a, float b, float fa, float fb, Vec3f u, Vec3f v, Vec3f h) {
  return axis_ok(a * u.x - b * u.y, a * v.x - b * v.y, fa * h.x + fb * h.y);
}

inline bool plane_overlaps(Vec3f n, Vec3f p, Vec3f h) {
  Vec3f lo, hi;
  for (int i = 0; i < 3; ++i) {
    if (n[i] > 0.0f) { lo[i] = -h[i] - p[i]; hi[i] = h[i] - p[i]; }
    else { lo[i] = h[i] - p[i]; hi[i] = -h[i] - p[i]; }
  }
  if (dot(n, lo) > 0.0f) return false;
  return dot(n, hi) >= 0.0f;
}

// Does the reference put `face` into octant `nd`?  (any vertex inside, else its SAT variant)
bool belongs(const Ctx &cx, const ONode &nd, int32_t face) {
  const float *p = cx.verts + (size_t)face * 9;
  const Vec3f v0(p), v1(p + 3), v2(p + 6);
  if (inside(nd, v0) || inside(nd, v1) || inside(nd, v2)) return true;

  const Vec3f c(nd.mn.x + (nd.mx.x - nd.mn.x) / 2.f, nd.mn.y + (nd.mx.y - nd.mn.y) / 2.f,
                nd.mn.z + (nd.mx.z - nd.mn.z) / 2.f);
  const Vec3f h = normalized(nd.mx - c);
  const Vec3f a = normalized(v0 - c), b = normalized(v1 - c), cc = normalized(v2 - c);
  const Vec3f e0 = b - a, e1 = cc - b, e2 = a - cc;

  float fx = std::fabs(e0.x), fy = std::fabs(e0.y), fz = std::fabs(e0.z);
  if (!probe_x(e0.z, e0.y, fz, fy, a, cc, h)) return false;
  if (!probe_y(e0.z, e0.x, fz, fx, a, cc, h)) return false;
  if (!probe_z(e0.y, e0.x, fy, fx, b, cc, h)) return false;
  fx = std::fabs(e1.x); fy = std::fabs(e1.y); fz = std::fabs(e1.z);
  if (!probe_x(e1.z, e1.y, fz, fy, a, cc, h)) return false;
  if (!probe_y(e1.z, e1.x, fz, fx, a, cc, h)) return false;
  if (!probe_z(e1.y, e1.x, fy, fx, a, b, h)) return false;
  fx = std::fabs(e2.x); fy = std::fabs(e2.y); fz = std::fabs(e2.z);
  if (!probe_x(e2.z, e2.y, fz, fy, a, b, h)) return false;
  if (!probe_y(e2.z, e2.x, fz, fx, a, b, h)) return false;
  if (!probe_z(e2.y, e2.x, fy, fx, b, cc, h)) return false;

  for (int k = 0; k < 3; ++k) {
    const float lo = std::min(std::min(a[k], b[k]), cc[k]), hi = std::max(std::max(a[k], b[k]), cc[k]);
    if (lo > h[k] || hi < -h[k]) return false;
  }
  return plane_overlaps(normalized(cross(a - b, a - cc)), a, h);
}

void subdivide(const Ctx &cx, ONode &n, int depth) {
  n.leaf = false;
  const float dx = (n.mx.x - n.mn.x) / 2, dy = (n.mx.y - n.mn.y) / 2, dz = (n.mx.z - n.mn.z) / 2;
  const Vec3f vx(dx, 0, 0), vy(0, dy, 0), vz(0, 0, dz), lo = n.mn, hi = n.mx;
  // octant corners with the reference's own expression trees (left-to-right sums)
  const Vec3f cmin[8] = {lo, lo + vz, lo + vy, lo + vy + vz, lo + vx, lo + vx + vz, lo + vx + vy, lo + vx + vy + vz};
  const Vec3f cmax[8] = {lo + vx + vy + vz, lo + vx + vy + 2 * vz, lo + vx + 2 * vy + vz, lo + vx + 2 * vy + 2 * vz,
                         lo + 2 * vx + vy + vz, hi - vy, hi - vz, hi};
  n.kids.resize(8);
  for (int k = 0; k < 8; ++k) {
    n.kids[k].reset(new ONode());
    n.kids[k]->mn = cmin[k];
    n.kids[k]->mx = cmax[k];
  }
  const bool parallel = n.faces.size() > 40000;
  auto fill = [&](int k) {
    ONode &ch = *n.kids[k];
    for (int32_t f : n.faces)
      if (belongs(cx, ch, f)) ch.faces.push_back(f);
  };
  if (parallel) {
    std::vector<std::future<void>> fut;
    for (int k = 0; k < 8; ++k) fut.push_back(std::async(std::launch::async, fill, k));
    for (auto &f : fut) f.get();
  } else {
    for (int k = 0; k < 8; ++k) fill(k);
  }
  std::vector<int32_t>().swap(n.faces);

  auto finish = [&](int k) {
    ONode &ch = *n.kids[k];
    if (ch.faces.empty() && ch.kids.empty()) ch.empty = true;
    if ((int)ch.faces.size() < cx.capacity || depth <= 0) ch.leaf = true;
    if ((int)ch.faces.size() > cx.capacity && depth > 0) subdivide(cx, ch, depth - 1);
  };
  if (parallel) {
    std::vector<std::future<void>> fut;
    for (int k = 0; k < 8; ++k) fut.push_back(std::async(std::launch::async, finish, k));
    for (auto &f : fut) f.get();
  } else {
    for (int k = 0; k < 8; ++k) finish(k);
  }
}

}  // namespace

RefOctree build_ref_octree(const float *verts, int32_t T, int capacity, int max_depth, int /*threads*/) {
  RefOctree out;
  ONode root;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {FLT_MIN, FLT_MIN, FLT_MIN};
  for (size_t v = 0; v < (size_t)T * 3; ++v)
    for (int a = 0; a < 3; ++a) {
      const float x = verts[3 * v + a];
      mn[a] = std::min(mn[a], x);
      mx[a] = std::max(mx[a], x);
    }
  root.mn = Vec3f(mn);
  root.mx = Vec3f(mx);
  root.faces.resize((size_t)T);
  for (int32_t i = 0; i < T; ++i) root.faces[i] = i;
  Ctx cx{verts, capacity};
  if (T > capacity) subdivide(cx, root, max_depth);
  else if (T == 0) root.empty = true;
  else root.leaf = true;
  out.root_is_leaf = root.leaf;

  // ---- flatten (pre-order) and collect (face, leaf) pairs of the reachable leaves ----
  std::vector<std::pair<int32_t, int32_t>> refs;
  struct Item { const ONode *n; int32_t parent; };
  std::vector<Item> stack;
  stack.push_back({&root, -1});
  while (!stack.empty()) {
    const Item it = stack.back();
    stack.pop_back();
    const ONode &n = *it.n;
    const int32_t id = (int32_t)out.parent.size();
    out.parent.push_back(it.parent);
    const float b[6] = {n.mn.x, n.mn.y, n.mn.z, n.mx.x, n.mx.y, n.mx.z};
    out.box.insert(out.box.end(), b, b + 6);
    const bool live_leaf = n.leaf && !n.empty;
    out.is_leaf.push_back(live_leaf ? 1 : 0);
    if (live_leaf) {
      out.n_leaves++;
      out.n_refs += (int64_t)n.faces.size();
      out.max_leaf = std::max<int64_t>(out.max_leaf, (int64_t)n.faces.size());
      for (int32_t f : n.faces) refs.emplace_back(f, id);
    } else if (!n.empty) {
      // inner node (a non-leaf without children -- exactly `capacity` faces -- hides its faces)
      out.n_inner++;
      for (int k = (int)n.kids.size() - 1; k >= 0; --k)
        if (!n.kids[k]->empty) stack.push_back({n.kids[k].get(), id});
    }
  }
  out.face_off.assign((size_t)T + 1, 0);
  for (auto &r : refs) out.face_off[(size_t)r.first + 1]++;
  for (int32_t i = 0; i < T; ++i) out.face_off[(size_t)i + 1] += out.face_off[i];
  out.face_leaf.resize(refs.size());
  std::vector<int32_t> cursor(out.face_off.begin(), out.face_off.end() - 1);
  for (auto &r : refs) out.face_leaf[(size_t)cursor[r.first]++] = r.second;
  return out;
}

}  // namespace rt
