// obj_loader.cpp -- see obj_loader.hpp.
#include "obj_loader.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace rt {

namespace {

bool read_file(const std::string &path, std::string &out) {
  FILE *f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  out.resize((size_t)std::max(0L, n));
  size_t got = n > 0 ? std::fread(&out[0], 1, (size_t)n, f) : 0;
  std::fclose(f);
  out.resize(got);
  return true;
}

std::string dir_of(const std::string &p) {
  size_t k = p.find_last_of("/\\");
  return k == std::string::npos ? std::string() : p.substr(0, k + 1);
}

// Iterate over '\n'-separated lines of a buffer without copying.
struct LineReader {
  const char *p, *end;
  explicit LineReader(const std::string &s) : p(s.data()), end(s.data() + s.size()) {}
  bool next(const char *&b, const char *&e) {
    if (p >= end) return false;
    b = p;
    const char *nl = (const char *)std::memchr(p, '\n', (size_t)(end - p));
    e = nl ? nl : end;
    p = nl ? nl + 1 : end;
    return true;
  }
};

inline bool starts(const char *b, const char *e, const char *lit) {
  size_t n = std::strlen(lit);
  return (size_t)(e - b) >= n && std::memcmp(b, lit, n) == 0;
}

// normalisation with the squared norm reduced left to right (see load_obj, face normals)
inline Vec3f normalized_ltr(Vec3f a) {
  const float z = (a.x * a.x + a.y * a.y) + a.z * a.z;
  if (z > 0.f) return a / std::sqrt(z);
  return a;
}

// Three whitespace-separated floats, each rounded correctly from decimal (what operator>> does).
inline void parse3(const char *b, float out[3]) {
  char *q = const_cast<char *>(b);
  for (int k = 0; k < 3; ++k) out[k] = std::strtof(q, &q);
}

}  // namespace

RtMaterial default_material() {
  // Tucano::Material::Mtl defaults (materials/mtl.hpp:21-39)
  RtMaterial m;
  m.kd[0] = m.kd[1] = m.kd[2] = 0.5f;
  m.ks[0] = m.ks[1] = m.ks[2] = 1.0f;
  m.ns = 10.f;
  m.ni = 0.f;
  m.illum = 0;
  return m;
}

bool load_mtl(const std::string &path, std::vector<RtMaterial> &out, std::vector<std::string> &names) {
  std::string buf;
  if (!read_file(path, buf)) {
    std::fprintf(stderr, "Cannot open %s\n", path.c_str());
    return false;
  }
  LineReader lr(buf);
  const char *b, *e;
  std::vector<std::string> tok;
  while (lr.next(b, e)) {
    if (b == e) continue;
    // the reference splits on single blanks: consecutive blanks yield empty tokens
    tok.clear();
    const char *s = b;
    for (const char *c = b; c <= e; ++c)
      if (c == e || *c == ' ') { tok.emplace_back(s, c); s = c + 1; }
    const std::string &k = tok[0];
    auto f = [&](size_t i) -> float { return i < tok.size() ? (float)std::atof(tok[i].c_str()) : 0.f; };
    if (k == "#") continue;
    if (k == "newmtl") {
      out.push_back(default_material());
      names.push_back(tok.size() > 1 ? tok[1] : std::string());
    } else if (out.empty()) {
      continue;  // parameter before any newmtl: the reference would touch an empty vector
    } else if (k == "Ns") out.back().ns = f(1);
    else if (k == "Kd") { out.back().kd[0] = f(1); out.back().kd[1] = f(2); out.back().kd[2] = f(3); }
    else if (k == "Ks") { out.back().ks[0] = f(1); out.back().ks[1] = f(2); out.back().ks[2] = f(3); }
    else if (k == "Ni") out.back().ni = f(1);
    else if (k == "illum") out.back().illum = tok.size() > 1 ? std::atoi(tok[1].c_str()) : 0;
    // Ka, d, Tf, map_* do not reach the ray tracer
  }
  return true;
}

BakedMesh load_obj(const std::string &obj_path, bool normalize) {
  std::string buf;
  if (!read_file(obj_path, buf)) throw std::runtime_error("Cannot open " + obj_path);
  const std::string base = dir_of(obj_path);

  BakedMesh m;
  std::vector<float> file_vn;                       // vn lines, in file order
  std::vector<std::vector<uint32_t>> groups(1);     // flat vertex-id lists, one per usemtl run
  std::vector<int> group_mat(1, -1);
  int current_mat = -1;

  LineReader lr(buf);
  const char *b, *e;
  while (lr.next(b, e)) {
    if (starts(b, e, "mtllib")) {
      std::string fn(b + std::min<ptrdiff_t>(7, e - b), e);
      fn.erase(std::remove(fn.begin(), fn.end(), '\r'), fn.end());
      load_mtl(base + fn, m.materials, m.material_names);
      if (m.materials.empty()) { m.materials.push_back(default_material()); m.material_names.emplace_back(); }
    } else if (starts(b, e, "usemtl")) {
      if (!groups.back().empty()) { groups.emplace_back(); group_mat.push_back(-1); }
      std::string name(b + std::min<ptrdiff_t>(7, e - b), e);
      for (size_t i = 0; i < m.material_names.size(); ++i)
        if (m.material_names[i] == name) current_mat = (int)i;
      group_mat.back() = current_mat;
    } else if (starts(b, e, "v ")) {
      float v[3];
      parse3(b + 2, v);
      m.obj_verts.insert(m.obj_verts.end(), v, v + 3);
    } else if (starts(b, e, "vn")) {
      float v[3] = {0, 0, 0};
      if (e - b > 3) parse3(b + 3, v);
      file_vn.insert(file_vn.end(), v, v + 3);
    } else if (starts(b, e, "f ")) {
      // "f v[/vt[/vn]] ..." -- only the vertex id of each element is used by the ray tracer
      const char *c = b + 2;
      while (c < e) {
        while (c < e && (*c == ' ' || *c == '\t' || *c == '\r')) ++c;
        if (c >= e) break;
        char *q;
        long id = std::strtol(c, &q, 10);
        if (q == c) break;
        // OBJ ids are 1-based; a negative id counts back from the vertices read so far.  Anything that does
        // not name an existing vertex would index past obj_verts / normals below: reject the file instead
        // (the reference reads out of bounds there, tucano/utils/objimporter.hpp:196-214).
        const long nv_here = (long)(m.obj_verts.size() / 3);
        const long zero_based = id < 0 ? nv_here + id : id - 1;
        if (id == 0 || zero_based < 0) throw std::runtime_error("face index " + std::to_string(id) + " out of range in " + obj_path);
        groups.back().push_back((uint32_t)zero_based);
        c = q;
        while (c < e && *c != ' ' && *c != '\t' && *c != '\r') ++c;  // skip /vt/vn
      }
    }
  }
  const size_t NV = m.obj_verts.size() / 3;
  for (const auto &g : groups)
    for (uint32_t id : g)
      if ((size_t)id >= NV)
        throw std::runtime_error("face index " + std::to_string((unsigned long)id + 1) + " exceeds the " + std::to_string(NV) +
                                 " vertices of " + obj_path);
  auto V = [&](uint32_t i) { return Vec3f(&m.obj_verts[3 * (size_t)i]); };

  // ---- vertex normals: list = file vn's followed by NV zero vectors; face normals are
  // accumulated at the *vertex* index, then everything is normalised (objimporter.hpp:50-74) ----
  std::vector<Vec3f> normals(file_vn.size() / 3 + NV);
  for (size_t i = 0; i < file_vn.size() / 3; ++i) normals[i] = Vec3f(&file_vn[3 * i]);
  for (const auto &g : groups)
    for (size_t i = 0; i + 2 < g.size(); i += 3) {
      Vec3f v1 = normalized(V(g[i + 2]) - V(g[i]));
      Vec3f v0 = normalized(V(g[i + 1]) - V(g[i]));
      Vec3f n = normalized(cross(v0, v1));
      normals[g[i]] += n; normals[g[i + 1]] += n; normals[g[i + 2]] += n;
    }
  for (auto &n : normals) n = normalized(n);

  // ---- centroid / bounding-sphere radius / normalisation (mesh.hpp:621-642) ----
  Vec3f c(0, 0, 0);
  for (size_t i = 0; i < NV; ++i) c = c + V((uint32_t)i);
  if (NV) c = c / (float)(unsigned)NV;
  float radius = 0.f;
  for (size_t i = 0; i < NV; ++i) radius = std::max(radius, norm(V((uint32_t)i) - c));
  m.centroid = c;
  m.radius = radius;
  m.norm_scale = (float)(1.0 / (double)radius);
  // shape matrix = scale(s) then translate(-centroid): x -> s*x + s*(-c)
  const float s = normalize ? m.norm_scale : 1.f;
  const Vec3f tr = normalize ? Vec3f(s * -c.x, s * -c.y, s * -c.z) : Vec3f(0, 0, 0);
  auto W = [&](uint32_t i) {
    Vec3f p = V(i);
    return Vec3f(s * p.x + tr.x, s * p.y + tr.y, s * p.z + tr.z);
  };

  // ---- faces, in group order (mesh.hpp:441-468) ----
  size_t T = 0;
  for (const auto &g : groups) T += g.size() / 3;
  m.verts.reserve(T * 9); m.fnormals.reserve(T * 3); m.vnormals.reserve(T * 9);
  m.mat_id.reserve(T); m.vertex_ids.reserve(T * 3);
  if (m.materials.empty()) { m.materials.push_back(default_material()); m.material_names.emplace_back(); }
  for (size_t gi = 0; gi < groups.size(); ++gi) {
    const auto &g = groups[gi];
    for (size_t i = 0; i + 2 < g.size(); i += 3) {
      const uint32_t ids[3] = {g[i], g[i + 1], g[i + 2]};
      for (int k = 0; k < 3; ++k) {
        Vec3f w = W(ids[k]);
        m.verts.insert(m.verts.end(), w.data(), w.data() + 3);
        const Vec3f &n = normals[ids[k]];
        m.vnormals.insert(m.vnormals.end(), n.data(), n.data() + 3);
        m.vertex_ids.push_back((int32_t)ids[k]);
      }
      // face normal from *object-space* normalised edges.  The reference normalises these two
      // edges through a dynamic-size Eigen expression (Vector4f::head(3) differences), whose
      // squared norm is reduced left to right, (x*x + y*y) + z*z -- unlike every fixed-size
      // 3-vector elsewhere (mesh.hpp:461-462 vs Redux.h); the cross product is fixed-size again.
      Vec3f e1 = normalized_ltr(V(ids[2]) - V(ids[0]));
      Vec3f e0 = normalized_ltr(V(ids[1]) - V(ids[0]));
      Vec3f fn = normalized(cross(e0, e1));
      m.fnormals.insert(m.fnormals.end(), fn.data(), fn.data() + 3);
      // OBJ files without mtllib/usemtl give material -1 in the reference (UB there); map to 0
      m.mat_id.push_back(group_mat[gi] < 0 ? 0 : group_mat[gi]);
    }
  }
  return m;
}

}  // namespace rt
