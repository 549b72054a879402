// obj_loader.cpp -- see obj_loader.hpp.
#include "obj_loader.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <thread>
#include <chrono>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

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

// the OBJ text, memory-mapped (read into a buffer when it cannot be mapped)
struct TextFile {
  const char *data = nullptr;
  size_t size = 0;
  void *mapped = nullptr;
  std::string fallback;
  bool open(const std::string &path) {
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd >= 0) {
      struct stat st;
      if (::fstat(fd, &st) == 0 && st.st_size > 0) {
        void *p = ::mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (p != MAP_FAILED) {
          ::madvise(p, (size_t)st.st_size, MADV_SEQUENTIAL);
          mapped = p; data = (const char *)p; size = (size_t)st.st_size;
        }
      }
      ::close(fd);
      if (mapped) return true;
    }
    if (!read_file(path, fallback)) return false;
    data = fallback.data(); size = fallback.size();
    return true;
  }
  ~TextFile() { if (mapped) ::munmap(mapped, size); }
};

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

namespace {

// What one thread found in its slice of the OBJ text (whole lines).  A face id counted back from "the vertices read
// so far" (a negative id) needs the number of vertices in the slices before this one: it is stored relative to the
// slice's first vertex, its position noted in `relative`, and the merge adds the slice's vertex offset.
struct ObjSlice {
  std::vector<float> v, vn;
  std::vector<int64_t> ids;        // zero-based vertex id of every face corner, in file order
  std::vector<size_t> relative;    // positions in ids that hold slice-relative values
  struct Event { size_t at; int kind; std::string name; };  // kind 0: mtllib, 1: usemtl; `at` = position in ids
  std::vector<Event> events;
  std::string error;
};

void parse_slice(const char *p, const char *end, ObjSlice &out, const std::string &obj_path) {
  while (p < end) {
    const char *b = p;
    const char *nl = (const char *)std::memchr(p, '\n', (size_t)(end - p));
    const char *e = nl ? nl : end;
    p = nl ? nl + 1 : end;
    if (starts(b, e, "mtllib")) {
      std::string fn(b + std::min<ptrdiff_t>(7, e - b), e);
      fn.erase(std::remove(fn.begin(), fn.end(), '\r'), fn.end());
      out.events.push_back({out.ids.size(), 0, std::move(fn)});
    } else if (starts(b, e, "usemtl")) {
      out.events.push_back({out.ids.size(), 1, std::string(b + std::min<ptrdiff_t>(7, e - b), e)});
    } else if (starts(b, e, "v ")) {
      float v[3];
      parse3(b + 2, v);
      out.v.insert(out.v.end(), v, v + 3);
    } else if (starts(b, e, "vn")) {
      float v[3] = {0, 0, 0};
      if (e - b > 3) parse3(b + 3, v);
      out.vn.insert(out.vn.end(), v, v + 3);
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
        if (id == 0) {
          if (out.error.empty()) out.error = "face index 0 out of range in " + obj_path;
          return;
        }
        if (id < 0) {
          out.relative.push_back(out.ids.size());
          out.ids.push_back((int64_t)id + (int64_t)(out.v.size() / 3));
        } else {
          out.ids.push_back((int64_t)id - 1);
        }
        c = q;
        while (c < e && *c != ' ' && *c != '\t' && *c != '\r') ++c;  // skip /vt/vn
      }
    }
  }
}

// f(k) for k in [0, n) on up to `threads` host threads (strided: neighbouring k cost about the same)
template <class F>
void parallel_for(size_t n, unsigned threads, F f) {
  threads = (unsigned)std::min<size_t>(std::max(1u, threads), std::max<size_t>(n, 1));
  if (threads <= 1) {
    for (size_t k = 0; k < n; ++k) f(k);
    return;
  }
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < threads; ++t)
    pool.emplace_back([=] { for (size_t k = t; k < n; k += threads) f(k); });
  for (auto &th : pool) th.join();
}
// f(begin, end) over contiguous blocks of [0, n)
template <class F>
void parallel_blocks(size_t n, unsigned threads, F f) {
  threads = (unsigned)std::min<size_t>(std::max(1u, threads), std::max<size_t>(n / 4096, 1));
  const size_t step = (n + threads - 1) / std::max(1u, threads);
  parallel_for(threads, threads, [=](size_t t) { f(std::min(n, t * step), std::min(n, (t + 1) * step)); });
}

}  // namespace

// The OBJ text is cut into slices at line ends and parsed by all host cores (number parsing is most of the load time
// of a large file), the slices are merged in file order, and the per-face work -- edge normalisations, face normals,
// the baked arrays -- runs in parallel over faces.  Every float sum whose order the reference fixes (a vertex' normal
// over its faces in file order, the centroid over the vertices) is still taken in that order.
BakedMesh load_obj(const std::string &obj_path, bool normalize) {
  const auto t_start = std::chrono::high_resolution_clock::now();
  const bool trace = std::getenv("RT_LOAD_TRACE") != nullptr;
  auto lap = [&](const char *what) {
    if (trace) std::fprintf(stderr, "[rt load] %8.3f ms  %s\n", std::chrono::duration<float, std::milli>(std::chrono::high_resolution_clock::now() - t_start).count(), what);
  };
  TextFile buf;
  if (!buf.open(obj_path)) throw std::runtime_error("Cannot open " + obj_path);
  lap("file mapped");
  const std::string base = dir_of(obj_path);
  unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  unsigned n_slices = buf.size < (1u << 20) ? 1u : hw;
  if (const char *forced = std::getenv("RT_LOAD_THREADS")) {  // (tests: any slice count, also on small files)
    const int n = std::atoi(forced);
    if (n >= 1 && n <= 64) { hw = (unsigned)n; n_slices = (unsigned)std::min<size_t>((size_t)n, std::max<size_t>(buf.size / 64, 1)); }
  }

  // ---- parse ----
  std::vector<ObjSlice> slices(n_slices);
  {
    std::vector<size_t> cut(n_slices + 1, buf.size);
    cut[0] = 0;
    for (unsigned k = 1; k < n_slices; ++k) {
      const size_t at = std::max(cut[k - 1], buf.size / n_slices * k);
      const void *nl = at < buf.size ? std::memchr(buf.data + at, '\n', buf.size - at) : nullptr;
      cut[k] = nl ? (size_t)((const char *)nl - buf.data) + 1 : buf.size;
    }
    parallel_for(n_slices, n_slices,
                 [&](size_t k) { parse_slice(buf.data + cut[k], buf.data + cut[k + 1], slices[k], obj_path); });
  }

  lap("parsed");
  // ---- merge in file order: vertices, vn's, face ids (negative ids resolved), usemtl runs ----
  BakedMesh m;
  std::vector<float> file_vn;                // vn lines, in file order
  std::vector<uint32_t> ids;                 // vertex ids of all face corners, in file order
  std::vector<size_t> group_begin(1, 0);     // a group = the face corners between two usemtl lines
  std::vector<int> group_mat(1, -1);
  int current_mat = -1;
  {
    size_t nv = 0, nvn = 0, nid = 0;
    for (const auto &sl : slices) { nv += sl.v.size(); nvn += sl.vn.size(); nid += sl.ids.size(); }
    m.obj_verts.reserve(nv); file_vn.reserve(nvn); ids.reserve(nid);
  }
  for (auto &sl : slices) {
    if (!sl.error.empty()) throw std::runtime_error(sl.error);
    const int64_t v_base = (int64_t)(m.obj_verts.size() / 3);
    for (size_t pos : sl.relative) {
      sl.ids[pos] += v_base;
      if (sl.ids[pos] < 0) throw std::runtime_error("relative face index reaches before the first vertex in " + obj_path);
    }
    size_t done = 0;
    auto copy_ids = [&](size_t upto) {
      for (; done < upto; ++done) {
        if (sl.ids[done] > 0xfffffffell) throw std::runtime_error("face index too large in " + obj_path);
        ids.push_back((uint32_t)sl.ids[done]);
      }
    };
    for (const auto &e : sl.events) {
      copy_ids(e.at);
      if (e.kind == 0) {
        load_mtl(base + e.name, m.materials, m.material_names);
        if (m.materials.empty()) { m.materials.push_back(default_material()); m.material_names.emplace_back(); }
      } else {
        if (ids.size() > group_begin.back()) { group_begin.push_back(ids.size()); group_mat.push_back(-1); }
        for (size_t k = 0; k < m.material_names.size(); ++k)
          if (m.material_names[k] == e.name) current_mat = (int)k;
        group_mat.back() = current_mat;
      }
    }
    copy_ids(sl.ids.size());
    m.obj_verts.insert(m.obj_verts.end(), sl.v.begin(), sl.v.end());
    file_vn.insert(file_vn.end(), sl.vn.begin(), sl.vn.end());
    ObjSlice().v.swap(sl.v);
  }
  group_begin.push_back(ids.size());
  const size_t NV = m.obj_verts.size() / 3;
  for (uint32_t id : ids)
    if ((size_t)id >= NV)
      throw std::runtime_error("face index " + std::to_string((unsigned long)id + 1) + " exceeds the " + std::to_string(NV) +
                               " vertices of " + obj_path);
  auto V = [&](uint32_t i) { return Vec3f(&m.obj_verts[3 * (size_t)i]); };
  lap("merged");

  // ---- faces: corner triples inside each group (a group's one or two left-over corners are dropped) ----
  std::vector<size_t> face_corner;           // position in ids of every face's first corner
  std::vector<int32_t> face_mat;
  for (size_t g = 0; g + 1 < group_begin.size(); ++g)
    for (size_t i = group_begin[g]; i + 2 < group_begin[g + 1]; i += 3) {
      face_corner.push_back(i);
      // OBJ files without mtllib/usemtl give material -1 in the reference (UB there); map to 0
      face_mat.push_back(group_mat[g] < 0 ? 0 : group_mat[g]);
    }
  const size_t T = face_corner.size();

  // ---- vertex normals: list = file vn's followed by NV zero vectors; face normals are
  // accumulated at the *vertex* index, then everything is normalised (objimporter.hpp:50-74) ----
  std::vector<Vec3f> normals(file_vn.size() / 3 + NV);
  for (size_t i = 0; i < file_vn.size() / 3; ++i) normals[i] = Vec3f(&file_vn[3 * i]);
  {
    std::vector<Vec3f> acc_n(T);
    parallel_blocks(T, hw, [&](size_t lo, size_t hi) {
      for (size_t f = lo; f < hi; ++f) {
        const uint32_t *g = &ids[face_corner[f]];
        const Vec3f v1 = normalized(V(g[2]) - V(g[0]));
        const Vec3f v0 = normalized(V(g[1]) - V(g[0]));
        acc_n[f] = normalized(cross(v0, v1));
      }
    });
    for (size_t f = 0; f < T; ++f) {  // (in file order: float sums are not associative)
      const uint32_t *g = &ids[face_corner[f]];
      normals[g[0]] += acc_n[f]; normals[g[1]] += acc_n[f]; normals[g[2]] += acc_n[f];
    }
  }
  parallel_blocks(normals.size(), hw, [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; ++i) normals[i] = normalized(normals[i]); });

  lap("vertex normals");
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

  lap("centroid, radius");
  // ---- faces, in group order (mesh.hpp:441-468) ----
  m.verts.resize(T * 9); m.fnormals.resize(T * 3); m.vnormals.resize(T * 9);
  m.mat_id.resize(T); m.vertex_ids.resize(T * 3);
  if (m.materials.empty()) { m.materials.push_back(default_material()); m.material_names.emplace_back(); }
  parallel_blocks(T, hw, [&](size_t lo, size_t hi) {
    for (size_t f = lo; f < hi; ++f) {
      const uint32_t *g = &ids[face_corner[f]];
      for (int k = 0; k < 3; ++k) {
        const Vec3f w = W(g[k]);
        const Vec3f &n = normals[g[k]];
        for (int a = 0; a < 3; ++a) {
          m.verts[9 * f + 3 * k + a] = w[a];
          m.vnormals[9 * f + 3 * k + a] = n[a];
        }
        m.vertex_ids[3 * f + k] = (int32_t)g[k];
      }
      // face normal from *object-space* normalised edges.  The reference normalises these two
      // edges through a dynamic-size Eigen expression (Vector4f::head(3) differences), whose
      // squared norm is reduced left to right, (x*x + y*y) + z*z -- unlike every fixed-size
      // 3-vector elsewhere (mesh.hpp:461-462 vs Redux.h); the cross product is fixed-size again.
      const Vec3f e1 = normalized_ltr(V(g[2]) - V(g[0]));
      const Vec3f e0 = normalized_ltr(V(g[1]) - V(g[0]));
      const Vec3f fn = normalized(cross(e0, e1));
      for (int a = 0; a < 3; ++a) m.fnormals[3 * f + a] = fn[a];
      m.mat_id[f] = face_mat[f];
    }
  });
  lap("baked");
  return m;
}

}  // namespace rt
