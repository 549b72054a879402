// rt_api.cu -- implementation of the C ABI in include/rt_api.h: scene bake upload, the per-frame
// wavefront driver (K1 nearest hit -> K2 shadow -> K3 shade/compact -> ... -> K3b fold) and the
// batched per-function entry points.  No CPU fallback: every compute entry point needs a CUDA
// device and fails with RT_ERR_NO_DEVICE otherwise.
#include "../../include/rt_api.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../host/bvh_builder.hpp"
#include "../host/obj_loader.hpp"
#include "../host/ref_octree.hpp"
#include "rt_kernels.cuh"
#include "rt_frame.cuh"
#include "rt_build.cuh"

using namespace rtd;

// ---------------------------------------------------------------------------------------------
// errors / globals
// ---------------------------------------------------------------------------------------------
namespace {

thread_local std::string g_err;
// Devices.  g_device is the device the calling thread is working on: rt_init() picks the default one, every
// entry point that takes a scene switches to the scene's own device first (use_device), so one host thread can
// drive scenes on several GPUs (rt_multi_*).
int g_default_device = -1;
thread_local int g_device = -1;
thread_local int g_sm_count = 0;
struct DeviceState {
  int sm_count = 0;
  int grids_opt = -1;  // value of persistent_ctas_per_sm the grids below were computed for
  // persistent grid sizes: [primary rays?][stats build?][plain scene?]
  int g_trace[2][2][2] = {}, g_shadow[2][2] = {}, g_frame[2][2] = {};
};
constexpr int kMaxDevices = 64;
DeviceState g_devs[kMaxDevices];
int g_opt_stats = 0;
int g_opt_leaf = 2;  // measured best on the 100 k / 1 M-triangle scenes (leaf tests are exact and expensive)
int g_opt_ctas_per_sm = 0;  // 0 = occupancy query
int g_opt_ref_candidates = 1;
int g_opt_graph_cond = 1;  // skip empty bounce levels inside the frame graph (conditional nodes)
int g_opt_fused = 2;       // frame as ONE persistent kernel (rt_frame.cuh): 0 never (wavefront pipeline of rt_kernels.cuh),
                           // 1 whenever the frame is eligible, 2 (default) when it is also small enough (below)
int g_opt_fused_max_kpix = 1200;  // auto mode: frames of at most this many thousand rays take the fused kernel
int g_opt_cont_min = 8;    // fused frame: child rays stay in the warp when at least this many lanes spawned one
int g_opt_gpu_build = 2;   // acceleration structures built on the device (rt_gpu_build.inl): 0 never, 1 whenever the scene has
                           // at least 64 primitives, 2 (default) from gpu_build_min_prims primitives on
int g_opt_gpu_build_min = 20000;
int g_opt_tile_w_log2 = 3;        // primary-ray tile of a warp: 2^k x (32 >> k) pixels (8 x 4)
int g_opt_direct_tile_w_log2 = 5; // ... when the pixels go straight to a host frame: 32 x 1, one 128-byte store per warp
int g_opt_direct_max_mb = 16;     // ... frames up to this size; larger ones are staged and copied in chunks (posted 128-byte
                                  // writes reach ~20 GB/s, a bulk copy ~55 GB/s: C5's 133 MB frame 6.6 vs 5.3 ms end to end)
thread_local int g_tile_override = 0;
int g_opt_tile_cull = 1;   // pixel tiles that cannot see the scene's bounds are background without tracing (tile_outside)
int g_opt_tile_order = 1;  // pixel tiles that can see the scene's bounds are handed out first, except when the pixels go
                           // straight to a host frame (set_tile_rect); 0 = row-major, 2 = always first
int g_opt_host_direct = 1; // rt_render: a page-locked host frame is written by the kernels themselves (no staging copy)
int g_opt_chunks = 2;      // rt_render: row chunks whose device->host copy overlaps the rendering of the next chunk
int g_opt_donate_min = 12;  // K2 on scenes with an octree filter or spheres, launches with few rounds per warp: idle lanes of a
                           // warp take over stack entries of busy lanes once at least this many lanes are idle
                           // (Trav::run_split); 0 = never; 100 + n = in every K1 / K2 launch (tests)

int fail(int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return fail(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

int use_device(int device) {
  if (device < 0 || device >= kMaxDevices) return fail(RT_ERR_INVALID, "device %d out of range", device);
  if (g_devs[device].sm_count == 0) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(RT_ERR_NO_DEVICE, "cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e));
    g_devs[device].sm_count = prop.multiProcessorCount;
  }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(RT_ERR_NO_DEVICE, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  g_device = device;
  g_sm_count = g_devs[device].sm_count;
  return RT_OK;
}

int ensure_device() {
  if (g_default_device < 0) {
    int rc = rt_init(0);
    if (rc) return rc;
  }
  return use_device(g_default_device);
}

// grow-only device buffer
std::atomic<long long> g_alloc_generation{0};  // bumped on every (re)allocation: invalidates captured graphs

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return RT_OK;
    ++g_alloc_generation;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return fail(RT_ERR_CUDA, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    cap = want;
    return RT_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T *as() const { return (T *)p; }
};

struct LevelStore {
  DevBuf ray_o, ray_d, ray_l, hit_t, hit_face, hit_list, hit_p, lit_list, vis, rec, child, type;
  int reserve(size_t n, size_t J) {
    n = std::max<size_t>(n, 1);
    int rc;
    if ((rc = ray_o.reserve(n * 16))) return rc;
    if ((rc = ray_d.reserve(n * 16))) return rc;
    if ((rc = ray_l.reserve(n * 8))) return rc;
    if ((rc = hit_t.reserve(n * 4))) return rc;
    if ((rc = hit_face.reserve(n * 4))) return rc;
    if ((rc = hit_list.reserve(n * 4))) return rc;
    if ((rc = hit_p.reserve(n * 16))) return rc;
    if ((rc = lit_list.reserve(n * 4))) return rc;
    if ((rc = vis.reserve(n * std::max<size_t>(J, 1)))) return rc;
    if ((rc = rec.reserve(n * 16))) return rc;
    if ((rc = child.reserve(n * 4))) return rc;
    if ((rc = type.reserve(n))) return rc;
    return RT_OK;
  }
  LevelBufs bufs() const {
    LevelBufs b;
    b.ray_o = ray_o.as<float4>(); b.ray_d = ray_d.as<float4>(); b.ray_l = ray_l.as<float2>();
    b.hit_t = hit_t.as<float>(); b.hit_face = hit_face.as<int32_t>(); b.hit_list = hit_list.as<int32_t>();
    b.hit_p = hit_p.as<float4>();
    b.lit_list = lit_list.as<int32_t>();
    b.vis = vis.as<uint8_t>();
    b.rec = rec.as<float4>(); b.child = child.as<int32_t>(); b.type = type.as<uint8_t>();
    return b;
  }
  void release() {
    ray_o.release(); ray_d.release(); ray_l.release(); hit_t.release(); hit_face.release(); hit_list.release(); hit_p.release(); lit_list.release();
    vis.release(); rec.release(); child.release(); type.release();
  }
};

}  // namespace

struct RtMesh {
  rt::BakedMesh mesh;
};

struct RtScene {
  int device = 0;  // CUDA device the scene lives on; every entry point switches to it first
  DevScene dev{};
  DevBuf nodes, prims, shade, mats, spheres, sphere_mat, oct_box, oct_face_off, oct_face_leaf;
  int64_t oct_stats[4] = {0, 0, 0, 0};
  float octree_ms = 0.f;
  std::vector<rt::PairNode> h_nodes;
  std::vector<int32_t> h_prim_face;
  int64_t n_leaves = 0;
  float build_ms = 0.f;
  int bvh_depth = 0;
  double sah_cost = 0.0;            // bvh_builder.cpp's definition, for either builder
  bool built_on_gpu = false;        // rt_gpu_build.inl
  DevBuf prim_order;                // GPU build: soup slot -> primitive id (host copy fetched on demand)
  float build_phase_ms[5] = {0, 0, 0, 0, 0};  // GPU build: upload, octree, sort, BVH splits, emit + bake
  int build_rounds[2] = {0, 0};     // GPU build: octree levels, BVH levels
  // per-frame workspace
  std::vector<LevelStore> levels;
  DevBuf frame_counts;    // FrameCounts
  DevBuf frame_params;    // FrameParams read by the frame kernels
  DevBuf out_rgba, out_face, out_t, out_rgbf, in_a, in_b;
  FrameCounts *h_counts = nullptr;  // pinned mirror
  cudaStream_t stream = nullptr;    // used when the caller passes the (uncapturable) legacy default stream
  // rt_render_submit / rt_render_wait: two frames in flight
  cudaStream_t copy_stream = nullptr;
  DevBuf slot_rgba[2];
  cudaEvent_t ev_rendered[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  cudaEvent_t ev_chunk[4] = {nullptr, nullptr, nullptr, nullptr};  // rt_render: chunk c rendered
  bool slot_busy[2] = {false, false};
  int submit_seq = 0;
  // fused frame kernel (rt_frame.cuh): deferred-ray slabs, chain records, counters
  DevBuf fq_o, fq_d, fq_x, frec_a, frec_b, fcounts;
  FusedCounts *h_fcounts = nullptr;  // pinned mirror
  // CUDA graph of the bounded-depth frame (memset + every kernel launch), replayed while its key matches
  cudaGraphExec_t graph_exec = nullptr;
  std::vector<long long> graph_key;
  long long ws_generation = 0;      // bumped whenever a workspace buffer is re-allocated
  size_t device_bytes() const {
    return nodes.cap + prims.cap + shade.cap + mats.cap + spheres.cap + sphere_mat.cap + oct_box.cap + oct_face_off.cap +
           oct_face_leaf.cap;
  }
};

// ---------------------------------------------------------------------------------------------
// lifecycle
// ---------------------------------------------------------------------------------------------
extern "C" int rt_api_version(void) { return RT_API_VERSION; }

extern "C" const char *rt_last_error(void) { return g_err.c_str(); }

extern "C" int rt_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(RT_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= n) return fail(RT_ERR_INVALID, "device %d out of range (0..%d)", device, n - 1);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(RT_ERR_NO_DEVICE, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(RT_ERR_NO_DEVICE, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  g_default_device = device;
  g_devs[device].sm_count = prop.multiProcessorCount;
  return use_device(device);
}

// rt_init for several GPUs of one box (rt_init(device_count, devices)): checks every device,
// makes devices[0] the default and enables peer access between all pairs that support it, so that the kernels of
// one GPU can store their pixels straight into another GPU's framebuffer over NVLink (rt_multi_*).
extern "C" int rt_init_devices(int count, const int *devices) {
  if (count <= 0 || !devices) return fail(RT_ERR_INVALID, "empty device list");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(RT_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  for (int i = 0; i < count; ++i) {
    if (devices[i] < 0 || devices[i] >= n || devices[i] >= kMaxDevices)
      return fail(RT_ERR_INVALID, "device %d out of range (0..%d)", devices[i], n - 1);
    for (int j = 0; j < i; ++j)
      if (devices[j] == devices[i]) return fail(RT_ERR_INVALID, "device %d listed twice", devices[i]);
  }
  int rc = rt_init(devices[0]);
  if (rc) return rc;
  for (int i = 0; i < count; ++i) {
    if ((rc = use_device(devices[i]))) return rc;
    for (int j = 0; j < count; ++j) {
      if (i == j) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
      if (!can) continue;
      e = cudaDeviceEnablePeerAccess(devices[j], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return fail(RT_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", devices[i], devices[j], cudaGetErrorString(e));
      cudaGetLastError();
    }
  }
  return use_device(devices[0]);
}

namespace {
// live scenes, so that rt_shutdown can release what the library allocated behind the caller's back
std::mutex g_scenes_mu;
std::vector<RtScene *> g_scenes;
void release_workspace(RtScene *sc);
}  // namespace

// Releases everything the library allocated lazily for rendering -- per-level ray queues, the fused kernel's slabs,
// captured frame graphs, staging buffers -- on every live scene, and forgets the device selection.  The scenes
// themselves (geometry on the device) stay valid and re-create their workspace on the next frame; they are freed by
// rt_scene_destroy / rt_multi_destroy.
extern "C" void rt_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_scenes_mu);
  for (RtScene *sc : g_scenes) release_workspace(sc);
  for (int d = 0; d < kMaxDevices; ++d) g_devs[d] = DeviceState();
  g_default_device = -1;
  g_device = -1;
}


extern "C" int rt_device_name(char *buf, size_t n) {
  int rc = ensure_device();
  if (rc) return rc;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, g_device));
  snprintf(buf, n, "%s (sm_%d%d, %d SMs)", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  return RT_OK;
}

extern "C" int rt_set_option(const char *key, int value) {
  if (!key) return fail(RT_ERR_INVALID, "null option key");
  if (!strcmp(key, "stats")) g_opt_stats = value ? 1 : 0;
  else if (!strcmp(key, "leaf_size")) g_opt_leaf = std::max(1, std::min(16, value));
  else if (!strcmp(key, "persistent_ctas_per_sm")) g_opt_ctas_per_sm = std::max(0, value);
  else if (!strcmp(key, "reference_candidates")) g_opt_ref_candidates = value ? 1 : 0;
  else if (!strcmp(key, "graph_conditionals")) g_opt_graph_cond = value ? 1 : 0;
  else if (!strcmp(key, "fused_frame")) g_opt_fused = std::max(0, std::min(2, value));
  else if (!strcmp(key, "fused_max_kpixels")) g_opt_fused_max_kpix = std::max(0, value);
  else if (!strcmp(key, "gpu_build")) g_opt_gpu_build = std::max(0, std::min(2, value));
  else if (!strcmp(key, "gpu_build_min_prims")) g_opt_gpu_build_min = std::max(64, value);
  else if (!strcmp(key, "host_direct")) g_opt_host_direct = std::max(0, std::min(2, value));
  else if (!strcmp(key, "host_direct_max_mb")) g_opt_direct_max_mb = std::max(0, value);
  else if (!strcmp(key, "tile_cull")) g_opt_tile_cull = value ? 1 : 0;
  else if (!strcmp(key, "tile_order")) g_opt_tile_order = std::max(0, std::min(2, value));
  else if (!strcmp(key, "tile_w_log2")) g_opt_tile_w_log2 = std::max(3, std::min(5, value));
  else if (!strcmp(key, "host_direct_tile_w_log2")) g_opt_direct_tile_w_log2 = std::max(3, std::min(5, value));
  else if (!strcmp(key, "render_chunks")) g_opt_chunks = std::max(1, std::min(4, value));
  else if (!strcmp(key, "donate_min_lanes")) g_opt_donate_min = std::max(0, std::min(132, value));
  else if (!strcmp(key, "continue_min_lanes")) g_opt_cont_min = std::max(1, std::min(33, value));
  else return fail(RT_ERR_INVALID, "unknown option '%s'", key);
  return RT_OK;
}

extern "C" void rt_default_params(RtParams *p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->width = 1000; p->height = 1000;  // src/main.cpp:8-9
  p->area_light = 0; p->point_light = 1;
  p->max_depth = -1;
  p->usteps = 5; p->vsteps = 5;
  p->area_len_x = 0.3f; p->area_len_y = 0.15f;
  p->band_rows = 8; p->band_rank = 0; p->band_world = 1;
  p->sphere_seed = 1u;
  p->sphere_radius = 1.00000012f;  // lightrep.getBoundingSphereRadius() of the reference (0x3f800001)
}

// ---------------------------------------------------------------------------------------------
// host-side scene bake
// ---------------------------------------------------------------------------------------------
extern "C" int rt_mesh_load_obj(const char *obj_path, RtMesh **out) {
  if (!obj_path || !out) return fail(RT_ERR_INVALID, "null argument");
  try {
    RtMesh *m = new RtMesh();
    m->mesh = rt::load_obj(obj_path, true);
    *out = m;
    return RT_OK;
  } catch (const std::exception &ex) {
    return fail(RT_ERR_IO, "%s", ex.what());
  }
}

extern "C" int rt_mesh_desc(const RtMesh *mesh, RtSceneDesc *desc) {
  if (!mesh || !desc) return fail(RT_ERR_INVALID, "null argument");
  memset(desc, 0, sizeof(*desc));
  const rt::BakedMesh &m = mesh->mesh;
  desc->n_faces = m.n_faces();
  desc->verts = m.verts.data();
  desc->face_normals = m.fnormals.data();
  desc->vertex_normals = m.vnormals.data();
  desc->material_id = m.mat_id.data();
  desc->n_materials = (int32_t)m.materials.size();
  desc->materials = m.materials.data();
  const float ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  memcpy(desc->model_matrix, ident, sizeof(ident));
  return RT_OK;
}

extern "C" int rt_mesh_info(const RtMesh *mesh, float centroid[3], float *radius, float *scale, int32_t *n_vertices) {
  if (!mesh) return fail(RT_ERR_INVALID, "null mesh");
  if (centroid) { centroid[0] = mesh->mesh.centroid.x; centroid[1] = mesh->mesh.centroid.y; centroid[2] = mesh->mesh.centroid.z; }
  if (radius) *radius = mesh->mesh.radius;
  if (scale) *scale = mesh->mesh.norm_scale;
  if (n_vertices) *n_vertices = mesh->mesh.n_vertices();
  return RT_OK;
}

extern "C" void rt_mesh_destroy(RtMesh *mesh) { delete mesh; }

// ---------------------------------------------------------------------------------------------
// scene upload
// ---------------------------------------------------------------------------------------------
namespace {

inline float hdot(const float *a, const float *b) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); }
inline float bits(int32_t v) { float f; memcpy(&f, &v, 4); return f; }

int upload(DevBuf &b, const void *src, size_t bytes) {
  int rc = b.reserve(std::max<size_t>(bytes, 16));
  if (rc) return rc;
  if (bytes) CUDA_TRY(cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice));
  return RT_OK;
}

}  // namespace

namespace {

// Everything rt_scene_create computes on the HOST: the flattened BVH, the primitive soup, the shading table and the
// reference-octree filter, ready to be uploaded to any number of devices (rt_multi_create bakes once).
struct HostBake {
  DevScene proto{};  // root box, model matrix, counts, oct_eps (device pointers are filled by the upload)
  std::vector<rt::PairNode> nodes;
  std::vector<float> prims, shade, mats, spheres, oct_packed;
  std::vector<int32_t> sphere_mat, oct_face_off, oct_face_leaf, prim_face;
  bool use_filter = false;
  int64_t n_leaves = 0, oct_stats[4] = {0, 0, 0, 0};
  int bvh_depth = 0;
  double sah_cost = 0.0;
  float octree_ms = 0.f, bake_ms = 0.f;
};

int validate_scene_desc(const RtSceneDesc *desc) {
  if (desc->n_faces < 0 || desc->n_spheres < 0 || desc->n_materials <= 0)
    return fail(RT_ERR_INVALID, "bad scene counts (faces %d, spheres %d, materials %d)", desc->n_faces,
                desc->n_spheres, desc->n_materials);
  if (desc->n_faces > 0 && (!desc->verts || !desc->face_normals || !desc->vertex_normals || !desc->material_id))
    return fail(RT_ERR_INVALID, "null face arrays");
  if (desc->n_spheres > 0 && (!desc->spheres || !desc->sphere_material)) return fail(RT_ERR_INVALID, "null sphere arrays");
  if (!desc->materials) return fail(RT_ERR_INVALID, "null material array");
  // geometry must be finite: the acceleration structure sorts and bins coordinates
  for (size_t k = 0; k < (size_t)desc->n_faces * 9; ++k)
    if (!std::isfinite(desc->verts[k]))
      return fail(RT_ERR_INVALID, "face %zu has a non-finite vertex coordinate", k / 9);
  for (size_t k = 0; k < (size_t)desc->n_spheres; ++k) {
    const float *sp = desc->spheres + 4 * k;
    if (!std::isfinite(sp[0]) || !std::isfinite(sp[1]) || !std::isfinite(sp[2]) || !std::isfinite(sp[3]) || sp[3] < 0.f)
      return fail(RT_ERR_INVALID, "sphere %zu is not finite or has a negative radius", k);
  }
  for (int i = 0; i < desc->n_faces; ++i)
    if (desc->material_id[i] < 0 || desc->material_id[i] >= desc->n_materials)
      return fail(RT_ERR_INVALID, "face %d has material id %d outside [0,%d)", i, desc->material_id[i], desc->n_materials);
  for (int i = 0; i < desc->n_spheres; ++i)
    if (desc->sphere_material[i] < 0 || desc->sphere_material[i] >= desc->n_materials)
      return fail(RT_ERR_INVALID, "sphere %d has a bad material id", i);
  return RT_OK;
}

void bake_scene(const RtSceneDesc *desc, HostBake &hb) {
  const auto t_start = std::chrono::high_resolution_clock::now();
  const int T = desc->n_faces, S = desc->n_spheres, N = T + S;
  // ---- reference root box: BoundingBox(Mesh&), src/boundingBox.cpp:14-43 (max starts at FLT_MIN) ----
  {
    float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
    float mx[3] = {std::numeric_limits<float>::min(), std::numeric_limits<float>::min(), std::numeric_limits<float>::min()};
    for (size_t v = 0; v < (size_t)T * 3; ++v)
      for (int a = 0; a < 3; ++a) {
        const float x = desc->verts[3 * v + a];
        mn[a] = std::min(mn[a], x);
        mx[a] = std::max(mx[a], x);
      }
    memcpy(hb.proto.root_min, mn, 12);
    memcpy(hb.proto.root_max, mx, 12);
  }
  memcpy(hb.proto.model, desc->model_matrix, sizeof(float) * 12);

  // ---- primitive boxes + BVH ----
  std::vector<rt::Aabb> boxes((size_t)N);
  std::vector<uint8_t> kind((size_t)N, 0);
  float gmin[3] = {1e30f, 1e30f, 1e30f}, gmax[3] = {-1e30f, -1e30f, -1e30f};
  for (int i = 0; i < T; ++i) {
    const float *v = desc->verts + (size_t)i * 9;
    for (int a = 0; a < 3; ++a) {
      boxes[i].mn[a] = std::min(v[a], std::min(v[3 + a], v[6 + a]));
      boxes[i].mx[a] = std::max(v[a], std::max(v[3 + a], v[6 + a]));
      gmin[a] = std::min(gmin[a], boxes[i].mn[a]); gmax[a] = std::max(gmax[a], boxes[i].mx[a]);
    }
  }
  for (int i = 0; i < S; ++i) {
    const float *s = desc->spheres + (size_t)i * 4;
    for (int a = 0; a < 3; ++a) {
      boxes[T + i].mn[a] = s[a] - s[3]; boxes[T + i].mx[a] = s[a] + s[3];
      gmin[a] = std::min(gmin[a], boxes[T + i].mn[a]); gmax[a] = std::max(gmax[a], boxes[T + i].mx[a]);
    }
    kind[T + i] = 1;
  }
  // ---- reference octree as a candidate filter (host/ref_octree.hpp) ----
  rt::RefOctree oct;
  bool &use_filter = hb.use_filter;
  use_filter = false;
  if (g_opt_ref_candidates && T > 0) {
    const auto t_oct = std::chrono::high_resolution_clock::now();
    oct = rt::build_ref_octree(desc->verts, T, 1000 /* src/flyscene.cpp:86 */, 15 /* MAX_DEPTH, src/boxTree.cpp:3 */, 0);
    hb.octree_ms = std::chrono::duration<float, std::milli>(std::chrono::high_resolution_clock::now() - t_oct).count();
    hb.oct_stats[0] = oct.n_leaves; hb.oct_stats[1] = oct.n_inner; hb.oct_stats[2] = oct.n_refs; hb.oct_stats[3] = oct.max_leaf;
    use_filter = !oct.root_is_leaf;  // a single-leaf octree offers every face whenever the root box is hit
    if (use_filter) {
      // Degenerate (sliver) faces: the reference's barycentric test is cancellation noise for them and
      // reports hits outside the triangle wherever the octree offers the face.  The
      // error of u,v is ~ eps_float * (|w|/|e|) / sin^2(theta) (theta = angle between the edges):
      //  * sin^2 <= 1e-7 (or a non-finite 1/det): phantom hits can lie anywhere in the face's octree
      //    leaves -> give the face the union of its leaf boxes as BVH bounds;
      //  * otherwise the hit point can leave the triangle by at most ~1e-6 * |e| / sin^2 -> pad the
      //    face's box by that much (only matters below sin^2 ~ 1e-2).
      // The candidate filter then decides exactly as the reference does.
      for (int i = 0; i < T; ++i) {
        const float *v = desc->verts + (size_t)i * 9;
        const float e0[3] = {v[6] - v[0], v[7] - v[1], v[8] - v[2]}, e1[3] = {v[3] - v[0], v[4] - v[1], v[5] - v[2]};
        const float d00 = hdot(e0, e0), d01 = hdot(e0, e1), d11 = hdot(e1, e1);
        const float det = d00 * d11 - d01 * d01;
        const float sin2 = det / (d00 * d11);
        if (sin2 > 1e-2f) continue;
        if (!(sin2 > 1e-7f) || !std::isfinite(1.f / det)) {
          for (int k = oct.face_off[i]; k < oct.face_off[i + 1]; ++k) {
            const float *b = &oct.box[(size_t)oct.face_leaf[k] * 6];
            for (int a = 0; a < 3; ++a) {
              boxes[i].mn[a] = std::min(boxes[i].mn[a], b[a]);
              boxes[i].mx[a] = std::max(boxes[i].mx[a], b[3 + a]);
            }
          }
        } else {
          const float extra = 1e-6f * std::sqrt(std::max(d00, d11)) / sin2;
          for (int a = 0; a < 3; ++a) { boxes[i].mn[a] -= extra; boxes[i].mx[a] += extra; }
        }
      }
    }
  }
  float diag = 1.f;
  if (N > 0) {
    const float dx = gmax[0] - gmin[0], dy = gmax[1] - gmin[1], dz = gmax[2] - gmin[2];
    diag = std::max(1.f, std::sqrt(dx * dx + dy * dy + dz * dz));
  }
  const float pad = 1e-5f * diag;  // see DESIGN.md "conservative culling"
  hb.proto.oct_eps = 1e-5f * diag;
  const int threads = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  rt::BvhBuildResult bvh = rt::build_bvh(boxes, kind, g_opt_leaf, pad, threads);
  hb.n_leaves = bvh.n_leaves;
  hb.bvh_depth = bvh.max_depth;
  hb.sah_cost = bvh.sah_cost;

  // ---- primitive soup in leaf order (80 B / primitive) ----
  std::vector<float> &prims = hb.prims;
  prims.assign((size_t)N * 20, 0.f);
  hb.prim_face.resize((size_t)N);
  for (int slot = 0; slot < N; ++slot) {
    const int p = bvh.prim_order[slot];
    float *q = &prims[(size_t)slot * 20];
    hb.prim_face[slot] = p;
    if (p < T) {
      const float *v = desc->verts + (size_t)p * 9;
      const float *n = desc->face_normals + (size_t)p * 3;
      const float *a = v, *b = v + 3, *c = v + 6;
      const float e0[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};  // v0 = vertices[2]-vertices[0], :800
      const float e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};  // v1 = vertices[1]-vertices[0], :801
      const float d00 = hdot(e0, e0), d01 = hdot(e0, e1), d11 = hdot(e1, e1);
      const float inv = 1 / (d00 * d11 - d01 * d01);                // :810
      const int illum = desc->materials[desc->material_id[p]].illum;
      q[0] = n[0]; q[1] = n[1]; q[2] = n[2]; q[3] = hdot(n, a);    // triangleNormal.dot(vertices[0]), :797
      q[4] = a[0]; q[5] = a[1]; q[6] = a[2]; q[7] = bits(p);
      q[8] = e0[0]; q[9] = e0[1]; q[10] = e0[2]; q[11] = d00;
      q[12] = e1[0]; q[13] = e1[1]; q[14] = e1[2]; q[15] = d11;
      q[16] = d01; q[17] = inv; q[18] = bits(illum == 9 ? (int32_t)PRIM_ILLUM9 : 0); q[19] = 0.f;
    } else {
      const int si = p - T;
      const float *s = desc->spheres + (size_t)si * 4;
      const int illum = desc->materials[desc->sphere_material[si]].illum;
      q[0] = s[0]; q[1] = s[1]; q[2] = s[2]; q[3] = s[3];
      q[7] = bits(p);
      q[18] = bits((int32_t)(PRIM_SPHERE | (illum == 9 ? PRIM_ILLUM9 : 0)));
    }
  }
  // ---- shading table in original face order (112 B / face) ----
  std::vector<float> &shade = hb.shade;
  shade.assign((size_t)std::max(T, 1) * 28, 0.f);
  for (int i = 0; i < T; ++i) {
    float *q = &shade[(size_t)i * 28];
    const float *v = desc->verts + (size_t)i * 9, *vn = desc->vertex_normals + (size_t)i * 9;
    const float *n = desc->face_normals + (size_t)i * 3;
    for (int k = 0; k < 3; ++k) {
      q[4 * k] = v[3 * k]; q[4 * k + 1] = v[3 * k + 1]; q[4 * k + 2] = v[3 * k + 2];
      q[12 + 4 * k] = vn[3 * k]; q[12 + 4 * k + 1] = vn[3 * k + 1]; q[12 + 4 * k + 2] = vn[3 * k + 2];
    }
    q[3] = bits(desc->material_id[i]);
    q[24] = n[0]; q[25] = n[1]; q[26] = n[2];
  }
  std::vector<float> &mats = hb.mats;
  mats.assign((size_t)desc->n_materials * 12, 0.f);
  for (int m = 0; m < desc->n_materials; ++m) {
    const RtMaterial &mt = desc->materials[m];
    float *q = &mats[(size_t)m * 12];
    q[0] = mt.kd[0]; q[1] = mt.kd[1]; q[2] = mt.kd[2]; q[3] = mt.ns;
    q[4] = mt.ks[0]; q[5] = mt.ks[1]; q[6] = mt.ks[2]; q[7] = mt.ni;
    q[8] = bits(mt.illum);
  }

  hb.nodes = std::move(bvh.nodes);
  {
    // bounds of everything the BVH holds = union of the root pair's two child boxes (segment_reaches_bvh)
    const float inf = std::numeric_limits<float>::infinity();
    float bmn[3] = {inf, inf, inf}, bmx[3] = {-inf, -inf, -inf};
    if (N > 0 && !hb.nodes.empty()) {
      const rt::PairNode &r = hb.nodes[0];
      for (int c = 0; c < 2; ++c) {
        int32_t code;
        memcpy(&code, &r.q[12 + c], 4);
        if (code == rt::kEmptyLeaf) continue;
        const float mn[3] = {r.q[4 * c + 0], r.q[4 * c + 2], r.q[8 + 2 * c]}, mx[3] = {r.q[4 * c + 1], r.q[4 * c + 3], r.q[9 + 2 * c]};
        for (int a = 0; a < 3; ++a) { bmn[a] = std::min(bmn[a], mn[a]); bmx[a] = std::max(bmx[a], mx[a]); }
      }
    }
    memcpy(hb.proto.bvh_min, bmn, 12);
    memcpy(hb.proto.bvh_max, bmx, 12);
  }
  if (use_filter) {
    const int n = oct.n_nodes();
    hb.oct_packed.assign((size_t)n * 8, 0.f);
    for (int i = 0; i < n; ++i) {
      float *q = &hb.oct_packed[(size_t)i * 8];
      const float *b = &oct.box[(size_t)i * 6];
      q[0] = b[0]; q[1] = b[1]; q[2] = b[2]; q[3] = bits(oct.parent[i]);
      q[4] = b[3]; q[5] = b[4]; q[6] = b[5];
    }
    hb.oct_face_off = std::move(oct.face_off);
    hb.oct_face_leaf = std::move(oct.face_leaf);
  }
  if (S > 0) {
    hb.spheres.assign(desc->spheres, desc->spheres + (size_t)S * 4);
    hb.sphere_mat.assign(desc->sphere_material, desc->sphere_material + S);
  }
  hb.proto.n_faces = T; hb.proto.n_spheres = S; hb.proto.n_prims = N; hb.proto.n_nodes = (int32_t)hb.nodes.size();
  hb.bake_ms = std::chrono::duration<float, std::milli>(std::chrono::high_resolution_clock::now() - t_start).count();
}

// Upload a baked scene to the CURRENT device (use_device) and create its per-frame workspace.
int upload_scene(const HostBake &hb, RtScene **out) {
  const auto t_start = std::chrono::high_resolution_clock::now();
  RtScene *sc = new RtScene();
  sc->device = g_device;
  sc->dev = hb.proto;
  sc->h_nodes = hb.nodes;
  sc->h_prim_face = hb.prim_face;
  sc->n_leaves = hb.n_leaves;
  sc->bvh_depth = hb.bvh_depth;
  sc->sah_cost = hb.sah_cost;
  sc->octree_ms = hb.octree_ms;
  memcpy(sc->oct_stats, hb.oct_stats, sizeof(hb.oct_stats));
  int rc;
  if ((rc = upload(sc->nodes, hb.nodes.data(), hb.nodes.size() * sizeof(rt::PairNode))) ||
      (rc = upload(sc->prims, hb.prims.data(), hb.prims.size() * 4)) ||
      (rc = upload(sc->shade, hb.shade.data(), hb.shade.size() * 4)) ||
      (rc = upload(sc->mats, hb.mats.data(), hb.mats.size() * 4)) ||
      (rc = upload(sc->spheres, hb.spheres.data(), hb.spheres.size() * 4)) ||
      (rc = upload(sc->sphere_mat, hb.sphere_mat.data(), hb.sphere_mat.size() * 4)) ||
      (hb.use_filter && ((rc = upload(sc->oct_box, hb.oct_packed.data(), hb.oct_packed.size() * 4)) ||
                         (rc = upload(sc->oct_face_off, hb.oct_face_off.data(), hb.oct_face_off.size() * 4)) ||
                         (rc = upload(sc->oct_face_leaf, hb.oct_face_leaf.data(), hb.oct_face_leaf.size() * 4)))) ||
      (rc = sc->frame_counts.reserve(sizeof(FrameCounts))) || (rc = sc->frame_params.reserve(sizeof(FrameParams)))) {
    rt_scene_destroy(sc);
    return rc;
  }
  if (cudaMallocHost((void **)&sc->h_counts, sizeof(FrameCounts)) != cudaSuccess ||
      cudaMallocHost((void **)&sc->h_fcounts, sizeof(FusedCounts)) != cudaSuccess ||
      cudaStreamCreateWithFlags(&sc->stream, cudaStreamNonBlocking) != cudaSuccess) {
    rt_scene_destroy(sc);
    return fail(RT_ERR_CUDA, "cudaMallocHost failed");
  }
  sc->dev.nodes = sc->nodes.as<float4>();
  sc->dev.prims = sc->prims.as<float4>();
  sc->dev.shade = sc->shade.as<float4>();
  sc->dev.mats = sc->mats.as<float4>();
  sc->dev.spheres = sc->spheres.as<float4>();
  sc->dev.sphere_mat = sc->sphere_mat.as<int32_t>();
  if (hb.use_filter) {
    sc->dev.oct_box = sc->oct_box.as<float4>();
    sc->dev.oct_face_off = sc->oct_face_off.as<int32_t>();
    sc->dev.oct_face_leaf = sc->oct_face_leaf.as<int32_t>();
  }
  sc->build_ms = hb.bake_ms + std::chrono::duration<float, std::milli>(std::chrono::high_resolution_clock::now() - t_start).count();
  {
    std::lock_guard<std::mutex> lk(g_scenes_mu);
    g_scenes.push_back(sc);
  }
  *out = sc;
  return RT_OK;
}

}  // namespace

#include "rt_gpu_build.inl"

namespace {
bool wants_gpu_build(const RtSceneDesc *desc) {
  const long long n = (long long)desc->n_faces + desc->n_spheres;
  return g_opt_gpu_build != 0 && n >= (g_opt_gpu_build == 1 ? 64 : (long long)g_opt_gpu_build_min);
}
}  // namespace

extern "C" int rt_scene_create(const RtSceneDesc *desc, RtScene **out) {
  if (!desc || !out) return fail(RT_ERR_INVALID, "null argument");
  const auto t0 = std::chrono::high_resolution_clock::now();
  auto since = [&] { return std::chrono::duration<float, std::milli>(std::chrono::high_resolution_clock::now() - t0).count(); };
  int rc = validate_scene_desc(desc);
  if (rc) return rc;
  if ((rc = ensure_device())) return rc;
  const bool trace = getenv("RT_BUILD_TRACE") != nullptr;
  if (trace) fprintf(stderr, "[rt build] %8.3f ms  scene description validated\n", since());
  if (wants_gpu_build(desc)) {
    bool too_deep = false;
    if ((rc = gpu_build_scene(desc, out, &too_deep))) return rc;
    if (trace) fprintf(stderr, "[rt build] %8.3f ms  rt_scene_create done (build arenas released)\n", since());
    // (build_ms = the whole call: validation, build, release of the build arenas)
    if (!too_deep) { (*out)->build_ms = since(); return RT_OK; }
    // (never taken: both builders guard their depth)
  }
  HostBake hb;
  bake_scene(desc, hb);
  return upload_scene(hb, out);
}

namespace {
void release_workspace(RtScene *sc) {
  if (use_device(sc->device)) return;
  cudaDeviceSynchronize();
  for (auto &l : sc->levels) l.release();
  sc->levels.clear();
  if (sc->graph_exec) { cudaGraphExecDestroy(sc->graph_exec); sc->graph_exec = nullptr; }
  sc->graph_key.clear();
  sc->fq_o.release(); sc->fq_d.release(); sc->fq_x.release(); sc->frec_a.release(); sc->frec_b.release();
  for (int k = 0; k < 2; ++k) { sc->slot_rgba[k].release(); sc->slot_busy[k] = false; }
  sc->out_rgba.release(); sc->out_face.release(); sc->out_t.release(); sc->out_rgbf.release();
  sc->in_a.release(); sc->in_b.release();
}
}  // namespace

extern "C" void rt_scene_destroy(RtScene *sc) {
  if (!sc) return;
  {
    std::lock_guard<std::mutex> lk(g_scenes_mu);
    g_scenes.erase(std::remove(g_scenes.begin(), g_scenes.end(), sc), g_scenes.end());
  }
  use_device(sc->device);
  sc->nodes.release(); sc->prims.release(); sc->shade.release(); sc->mats.release();
  sc->spheres.release(); sc->sphere_mat.release();
  sc->oct_box.release(); sc->oct_face_off.release(); sc->oct_face_leaf.release(); sc->prim_order.release();
  for (auto &l : sc->levels) l.release();
  sc->frame_counts.release(); sc->frame_params.release();
  if (sc->graph_exec) cudaGraphExecDestroy(sc->graph_exec);
  if (sc->stream) cudaStreamDestroy(sc->stream);
  if (sc->copy_stream) cudaStreamDestroy(sc->copy_stream);
  for (int k = 0; k < 2; ++k) {
    if (sc->ev_rendered[k]) cudaEventDestroy(sc->ev_rendered[k]);
    if (sc->ev_copied[k]) cudaEventDestroy(sc->ev_copied[k]);
    sc->slot_rgba[k].release();
  }
  for (int k = 0; k < 4; ++k)
    if (sc->ev_chunk[k]) cudaEventDestroy(sc->ev_chunk[k]);
  sc->out_rgba.release(); sc->out_face.release(); sc->out_t.release(); sc->out_rgbf.release();
  sc->in_a.release(); sc->in_b.release();
  if (sc->h_counts) cudaFreeHost(sc->h_counts);
  sc->fq_o.release(); sc->fq_d.release(); sc->fq_x.release(); sc->frec_a.release(); sc->frec_b.release(); sc->fcounts.release();
  if (sc->h_fcounts) cudaFreeHost(sc->h_fcounts);
  delete sc;
}

extern "C" int rt_scene_root_box(const RtScene *sc, float mn[3], float mx[3]) {
  if (!sc) return fail(RT_ERR_INVALID, "null scene");
  memcpy(mn, sc->dev.root_min, 12);
  memcpy(mx, sc->dev.root_max, 12);
  return RT_OK;
}

extern "C" int rt_scene_info(const RtScene *sc, int64_t *n_nodes, int64_t *n_leaves, int64_t *n_tris,
                             int64_t *device_bytes, float *build_ms) {
  if (!sc) return fail(RT_ERR_INVALID, "null scene");
  if (n_nodes) *n_nodes = sc->dev.n_nodes;
  if (n_leaves) *n_leaves = sc->n_leaves;
  if (n_tris) *n_tris = sc->dev.n_prims;
  if (device_bytes) *device_bytes = (int64_t)sc->device_bytes();
  if (build_ms) *build_ms = sc->build_ms;
  return RT_OK;
}

extern "C" int rt_ref_octree_stats(const RtSceneDesc *desc, int32_t capacity, int64_t out[4]) {
  if (!desc || !out) return fail(RT_ERR_INVALID, "null argument");
  if (desc->n_faces > 0 && !desc->verts) return fail(RT_ERR_INVALID, "null verts");
  const rt::RefOctree oct = rt::build_ref_octree(desc->verts, desc->n_faces, capacity, 15, 0);
  out[0] = oct.n_leaves; out[1] = oct.n_inner; out[2] = oct.n_refs; out[3] = oct.max_leaf;
  return RT_OK;
}

// Host only: build the BVH rt_scene_create would build for the triangles of desc and check the
// invariants the device traversal relies on.
extern "C" int rt_bvh_check(const RtSceneDesc *desc, int32_t leaf_size, int64_t out[6]) {
  if (!desc || !out) return fail(RT_ERR_INVALID, "null argument");
  if (desc->n_faces < 0 || (desc->n_faces > 0 && !desc->verts)) return fail(RT_ERR_INVALID, "bad face arrays");
  const int T = desc->n_faces;
  std::vector<rt::Aabb> boxes((size_t)T);
  std::vector<uint8_t> kind((size_t)T, 0);
  for (int i = 0; i < T; ++i) {
    const float *v = desc->verts + (size_t)i * 9;
    for (int a = 0; a < 3; ++a) {
      if (!std::isfinite(v[a]) || !std::isfinite(v[3 + a]) || !std::isfinite(v[6 + a]))
        return fail(RT_ERR_INVALID, "face %d has a non-finite vertex coordinate", i);
      boxes[i].mn[a] = std::min(v[a], std::min(v[3 + a], v[6 + a]));
      boxes[i].mx[a] = std::max(v[a], std::max(v[3 + a], v[6 + a]));
    }
  }
  const int threads = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  const rt::BvhBuildResult bvh = rt::build_bvh(boxes, kind, leaf_size, 0.f, threads);
  if ((int)bvh.prim_order.size() != T) return fail(RT_ERR_INVALID, "prim_order has %zu entries for %d faces", bvh.prim_order.size(), T);
  // every node is reached once from the root, every soup slot lies in exactly one leaf, every leaf's box
  // (as stored in its parent) contains its primitives, every child box lies inside its parent's
  std::vector<uint8_t> seen_prim((size_t)T, 0), seen_node(bvh.nodes.size(), 0), seen_slot((size_t)T, 0);
  struct Item { int32_t node; int depth; float mn[3], mx[3]; };
  std::vector<Item> todo;
  const float inf = std::numeric_limits<float>::infinity();
  todo.push_back({0, 1, {-inf, -inf, -inf}, {inf, inf, inf}});
  int64_t n_leaves = 0, max_leaf = 0, max_depth = 0, refs = 0;
  while (!todo.empty()) {
    const Item it = todo.back();
    todo.pop_back();
    if (it.node < 0 || (size_t)it.node >= bvh.nodes.size()) return fail(RT_ERR_INVALID, "child index %d out of range", it.node);
    if (seen_node[(size_t)it.node]++) return fail(RT_ERR_INVALID, "node %d reached twice", it.node);
    max_depth = std::max<int64_t>(max_depth, it.depth);
    const rt::PairNode &pn = bvh.nodes[(size_t)it.node];
    for (int c = 0; c < 2; ++c) {
      int32_t code;
      memcpy(&code, &pn.q[12 + c], 4);
      if (code == rt::kEmptyLeaf) {
        // only the root of an EMPTY scene may have unused slots: the device's slab test would accept the
        // inverted box of an empty slot (see bvh_builder.cpp)
        if (T > 0) return fail(RT_ERR_INVALID, "node %d has an empty child slot", it.node);
        continue;
      }
      if (c == 1 && T == 1) {  // single primitive: both slots name the same leaf
        int32_t code0;
        memcpy(&code0, &pn.q[12], 4);
        if (code == code0) continue;
      }
      const float mn[3] = {pn.q[4 * c + 0], pn.q[4 * c + 2], pn.q[8 + 2 * c]};
      const float mx[3] = {pn.q[4 * c + 1], pn.q[4 * c + 3], pn.q[9 + 2 * c]};
      for (int a = 0; a < 3; ++a)
        if (!(mn[a] <= mx[a]) || mn[a] < it.mn[a] || mx[a] > it.mx[a])
          return fail(RT_ERR_INVALID, "node %d child %d: box not inside its parent's", it.node, c);
      if (code >= 0) {
        Item ch{code, it.depth + 1, {mn[0], mn[1], mn[2]}, {mx[0], mx[1], mx[2]}};
        todo.push_back(ch);
        continue;
      }
      const uint32_t lc = (uint32_t)~code;
      const int first = (int)(lc >> 5), count = (int)(lc & 15u) + 1;
      ++n_leaves;
      max_leaf = std::max<int64_t>(max_leaf, count);
      if (first < 0 || first + count > T) return fail(RT_ERR_INVALID, "leaf range [%d,%d) outside the soup", first, first + count);
      for (int k = first; k < first + count; ++k) {
        if (seen_slot[(size_t)k]++) return fail(RT_ERR_INVALID, "soup slot %d in two leaves", k);
        const int p = bvh.prim_order[(size_t)k];
        if (p < 0 || p >= T || seen_prim[(size_t)p]++) return fail(RT_ERR_INVALID, "face %d referenced twice or out of range", p);
        ++refs;
        for (int a = 0; a < 3; ++a)
          if (boxes[(size_t)p].mn[a] < mn[a] || boxes[(size_t)p].mx[a] > mx[a])
            return fail(RT_ERR_INVALID, "face %d sticks out of its leaf box", p);
      }
    }
  }
  if (refs != T) return fail(RT_ERR_INVALID, "%lld of %d faces are in leaves", (long long)refs, T);
  for (size_t i = 0; i < seen_node.size(); ++i)
    if (!seen_node[i]) return fail(RT_ERR_INVALID, "node %zu is unreachable", i);
  if (max_depth > RT_STACK_SIZE - 4) return fail(RT_ERR_INVALID, "tree depth %lld exceeds the traversal stack", (long long)max_depth);
  out[0] = (int64_t)bvh.nodes.size(); out[1] = n_leaves; out[2] = max_depth; out[3] = max_leaf; out[4] = refs;
  out[5] = (int64_t)(bvh.sah_cost * 1000.0);
  return RT_OK;
}

// Which builder made the scene's acceleration structures, and what came out.
extern "C" int rt_scene_build_info(const RtScene *sc, int64_t out[16]) {
  if (!sc || !out) return fail(RT_ERR_INVALID, "null argument");
  out[0] = sc->built_on_gpu ? 1 : 0;
  out[1] = sc->dev.n_nodes; out[2] = sc->n_leaves; out[3] = sc->bvh_depth;
  out[4] = (int64_t)(sc->sah_cost * 1000.0);
  for (int k = 0; k < 4; ++k) out[5 + k] = sc->oct_stats[k];
  out[9] = (int64_t)(sc->build_ms * 1000.f);
  for (int k = 0; k < 5; ++k) out[10 + k] = (int64_t)(sc->build_phase_ms[k] * 1000.f);
  out[15] = (int64_t)sc->build_rounds[0] * 1000 + sc->build_rounds[1];
  return RT_OK;
}

extern "C" int rt_scene_debug_bvh(const RtScene *sc_in, float *nodes, int64_t nodes_cap, int32_t *tri_face, int64_t tri_cap) {
  if (!sc_in) return fail(RT_ERR_INVALID, "null scene");
  RtScene *sc = const_cast<RtScene *>(sc_in);
  if (sc->built_on_gpu && sc->h_nodes.empty() && sc->dev.n_nodes > 0) {
    // scenes built on the device keep no host copy: fetch one on first use
    int rc = use_device(sc->device);
    if (rc) return rc;
    sc->h_nodes.resize((size_t)sc->dev.n_nodes);
    sc->h_prim_face.resize((size_t)sc->dev.n_prims);
    CUDA_TRY(cudaMemcpy(sc->h_nodes.data(), sc->nodes.p, sc->h_nodes.size() * sizeof(rt::PairNode), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(sc->h_prim_face.data(), sc->prim_order.p, sc->h_prim_face.size() * 4, cudaMemcpyDeviceToHost));
  }
  if (nodes) {
    if (nodes_cap < (int64_t)sc->h_nodes.size()) return fail(RT_ERR_INVALID, "nodes buffer too small");
    memcpy(nodes, sc->h_nodes.data(), sc->h_nodes.size() * sizeof(rt::PairNode));
  }
  if (tri_face) {
    if (tri_cap < (int64_t)sc->h_prim_face.size()) return fail(RT_ERR_INVALID, "tri buffer too small");
    memcpy(tri_face, sc->h_prim_face.data(), sc->h_prim_face.size() * 4);
  }
  return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// frame driver
// ---------------------------------------------------------------------------------------------
extern "C" int rt_local_rows(const RtParams *p) {
  if (!p || p->height <= 0) return 0;
  if (p->band_world <= 1) return p->height;
  const int B = std::max(1, p->band_rows);
  const int bands = (p->height + B - 1) / B;
  int rows = 0;
  for (int b = p->band_rank; b < bands; b += p->band_world) rows += std::min(B, p->height - b * B);
  return rows;
}

extern "C" int rt_local_row_map(const RtParams *p, int32_t *rows_out) {
  if (!p || !rows_out) return fail(RT_ERR_INVALID, "null argument");
  if (p->band_world <= 1) { for (int r = 0; r < p->height; ++r) rows_out[r] = r; return RT_OK; }
  const int B = std::max(1, p->band_rows);
  const int bands = (p->height + B - 1) / B;
  int k = 0;
  for (int b = p->band_rank; b < bands; b += p->band_world)
    for (int r = b * B; r < std::min(p->height, (b + 1) * B); ++r) rows_out[k++] = r;
  return RT_OK;
}

namespace {

// The 25 sample offsets of the reference's spherical light (createSpherePoint, src/flyscene.cpp:974-993)
// with the random_device draws replaced by the counter hash documented in rt_api.h (RtParams.sphere_seed).
// Types follow the reference's expressions: randomno, theta, phi, x, y, z are float; 2.0f * M_PI * randomno
// and acos(2.0 * randomno - 1.0) are evaluated in double; sin / cos of the float angles resolve to the
// float overloads (the translation unit is `using namespace std`); Vector3f(x, y, z) / 5 divides by 5.0f.
constexpr int RT_SPHERE_SAMPLES = 25;
inline uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
void sphere_offsets(uint32_t seed, float radius, float *out /*[25][3]*/) {
  for (int k = 0; k < RT_SPHERE_SAMPLES; ++k) {
    const uint32_t h = lowbias32(seed * 0x9E3779B9U + (uint32_t)k);
    const float randomno = (float)(h >> 8) * (1.0f / 16777216.0f);
    const float theta = (float)((double)2.0f * M_PI * (double)randomno);
    const float phi = (float)std::acos(2.0 * (double)randomno - 1.0);
    const float x = radius * sinf(phi) * cosf(theta);
    const float y = radius * sinf(phi) * sinf(theta);
    const float z = radius * cosf(phi);
    out[3 * k] = x / 5.0f; out[3 * k + 1] = y / 5.0f; out[3 * k + 2] = z / 5.0f;
  }
}

// Screen rectangle (in tiles of the frame's tile shape) of the scene's bounds: its tiles are handed out first
// (tile_xy in rt_device.cuh).  Only an ordering hint: a wrong or empty rectangle cannot change a pixel.
void set_tile_rect(FrameParams &fp, const RtScene *sc) {
  fp.tile_rect[0] = fp.tile_rect[1] = fp.tile_rect[2] = fp.tile_rect[3] = 0;
  fp.tile_order_on = 0; fp.tile_cull = 0;
  if ((!g_opt_tile_order && !g_opt_tile_cull) || sc->dev.n_prims <= 0) return;
  const int tw = 1 << fp.tile_w_log2, th = 32 >> fp.tile_w_log2;
  const int tiles_x = (fp.width + tw - 1) / tw, tiles_y = (fp.local_rows + th - 1) / th;
  // camera space of the bounds' corners: view_inv = [R | t] maps camera to world
  const float *m = fp.view_inv;
  const double R[9] = {m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10]}, t[3] = {m[3], m[7], m[11]};
  const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
  if (!(std::fabs(det) > 1e-20)) return;
  const double Ri[9] = {(R[4] * R[8] - R[5] * R[7]) / det, (R[2] * R[7] - R[1] * R[8]) / det, (R[1] * R[5] - R[2] * R[4]) / det,
                        (R[5] * R[6] - R[3] * R[8]) / det, (R[0] * R[8] - R[2] * R[6]) / det, (R[2] * R[3] - R[0] * R[5]) / det,
                        (R[3] * R[7] - R[4] * R[6]) / det, (R[1] * R[6] - R[0] * R[7]) / det, (R[0] * R[4] - R[1] * R[3]) / det};
  auto to_cam = [&](const double w[3], double c[3]) {
    const double d[3] = {w[0] - t[0], w[1] - t[1], w[2] - t[2]};
    for (int a = 0; a < 3; ++a) c[a] = Ri[3 * a] * d[0] + Ri[3 * a + 1] * d[1] + Ri[3 * a + 2] * d[2];
  };
  const double eye_w[3] = {fp.eye[0], fp.eye[1], fp.eye[2]};
  double e[3];
  to_cam(eye_w, e);
  double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
  for (int corner = 0; corner < 8; ++corner) {
    const double w[3] = {(corner & 1) ? sc->dev.bvh_max[0] : sc->dev.bvh_min[0], (corner & 2) ? sc->dev.bvh_max[1] : sc->dev.bvh_min[1],
                         (corner & 4) ? sc->dev.bvh_max[2] : sc->dev.bvh_min[2]};
    if (!std::isfinite(w[0]) || !std::isfinite(w[1]) || !std::isfinite(w[2])) return;
    double c[3];
    to_cam(w, c);
    const double dz = c[2] - e[2];
    if (!(dz < -1e-9)) return;  // a corner beside or behind the eye: no bounded rectangle
    const double sdist = (-1.0 - e[2]) / dz;  // the screen is the plane z = -1 of camera space (screenToWorld)
    if (!(sdist > 0.0)) return;
    const double sx = e[0] + sdist * (c[0] - e[0]), sy = e[1] + sdist * (c[1] - e[1]);
    const double px = (sx / fp.cam_sx + 1.0) * 0.5 * fp.viewport[2] + fp.viewport[0];
    const double py = (1.0 - sy / fp.cam_sy) * 0.5 * fp.viewport[3] + fp.viewport[1];
    xmin = std::min(xmin, px); xmax = std::max(xmax, px); ymin = std::min(ymin, py); ymax = std::max(ymax, py);
  }
  if (!(xmin <= xmax) || !(ymin <= ymax)) return;
  auto clampi = [](double v, int lo, int hi) { return (int)std::max((double)lo, std::min((double)hi, v)); };
  int x0 = clampi(std::floor(xmin / tw) - 1, 0, tiles_x), x1 = clampi(std::floor(xmax / tw) + 2, 0, tiles_x);
  int y0 = clampi(std::floor(ymin / th) - 1, 0, tiles_y), y1 = clampi(std::floor(ymax / th) + 2, 0, tiles_y);
  if (fp.band_world > 1) { y0 = 0; y1 = tiles_y; }  // interleaved bands: local rows are not global rows; order by column only
  if (x1 <= x0 || y1 <= y0 || (x0 == 0 && y0 == 0 && x1 == tiles_x && y1 == tiles_y)) return;
  fp.tile_rect[0] = x0; fp.tile_rect[1] = y0; fp.tile_rect[2] = x1; fp.tile_rect[3] = y1;
  // (pixels stored straight into a host frame -- rt_render's direct path sets the tile override -- keep row-major order)
  fp.tile_order_on = (g_opt_tile_order == 2 || (g_opt_tile_order == 1 && g_tile_override == 0)) ? 1 : 0;
  fp.tile_cull = g_opt_tile_cull ? 1 : 0;
}

int fill_frame(FrameParams &fp, const RtCamera *cam, const RtLights *lights, const RtParams *p) {
  memset(&fp, 0, sizeof(fp));
  if (!lights || !p) return fail(RT_ERR_INVALID, "null lights/params");
  if (lights->n < 0 || lights->n > RT_MAX_LIGHTS)
    return fail(RT_ERR_LIMIT, "%d lights: the reference's visibleLights[25] buffer caps lights at %d", lights->n, RT_MAX_LIGHTS);
  if (lights->n > 0 && !lights->pos) return fail(RT_ERR_INVALID, "null light positions");
  const bool sphere_mode = !p->point_light && !p->area_light;
  if (p->area_light && !p->point_light) {
    if (p->usteps <= 0 || p->vsteps <= 0 || p->usteps * p->vsteps > RT_MAX_SAMPLES)
      return fail(RT_ERR_LIMIT, "area grid %dx%d exceeds %d samples (visibleLights[25])", p->usteps, p->vsteps, RT_MAX_SAMPLES);
  }
  if (cam) {
    memcpy(fp.eye, cam->eye, 12);
    memcpy(fp.view_inv, cam->view_inv, 48);
    memcpy(fp.viewport, cam->viewport, 16);
    // Camera::getPerspectiveScale / screenToWorld scale, tucano/camera.hpp:166-168,263-266
    const float pscale = (float)((double)1.0f / std::tan((double)(cam->fovy / 2.0f) * (M_PI / (double)180.0f)));
    const float scale = (float)(1.0 / (double)pscale);
    fp.cam_sx = cam->aspect * scale;
    fp.cam_sy = scale;
  }
  fp.width = p->width; fp.height = p->height;
  fp.band_rows = std::max(1, p->band_rows); fp.band_rank = p->band_rank; fp.band_world = p->band_world;
  fp.local_rows = rt_local_rows(p);
  fp.out_full_frame = p->out_full_frame ? 1 : 0;
  fp.tile_w_log2 = g_tile_override ? g_tile_override : g_opt_tile_w_log2;
  fp.n_lights = lights->n;
  for (int i = 0; i < lights->n * 3; ++i) fp.lights[i] = lights->pos[i];
  memcpy(fp.light_color, lights->color, 12);
  fp.area_light = p->area_light; fp.point_light = p->point_light;
  fp.usteps = p->usteps; fp.vsteps = p->vsteps;
  fp.area_len_x = p->area_len_x; fp.area_len_y = p->area_len_y;
  fp.max_depth = p->max_depth;
  fp.guard_depth = 64;
  if (sphere_mode) {
    // 25 samples per light, indexed like a 25 x 1 grid
    fp.sphere_mode = 1;
    fp.usteps = RT_SPHERE_SAMPLES; fp.vsteps = 1;
    sphere_offsets(p->sphere_seed, p->sphere_radius, fp.sphere_off);
  }
  // area-light sample table for the scene lights (same expression as the device's area_sample)
  fp.have_sample_table = 0;
  if (!p->point_light && p->area_light) {
    const int S = p->usteps * p->vsteps;
    if (lights->n * S <= RT_SAMPLE_TABLE) {
      for (int l = 0; l < lights->n; ++l) {
        const float *c = lights->pos + 3 * l;
        const float ux = c[0] + p->area_len_x * 1.0f, vy = c[1] + p->area_len_y * 1.0f, uz = c[2] + p->area_len_x * 0.0f;
        for (int k = 0; k < S; ++k) {
          const int i = k / p->vsteps, j = k - i * p->vsteps;
          float *q = fp.sample_table + 3 * (l * S + k);
          q[0] = (float)((double)i + 0.5) * (ux / (float)p->usteps);
          q[1] = (float)((double)j + 0.5) * (vy / (float)p->vsteps);
          q[2] = uz;
        }
      }
      fp.have_sample_table = 1;
    }
  }
  return RT_OK;
}

struct EventTimer {
  struct Span { int cat; cudaEvent_t a, b; };
  std::vector<Span> spans;
  bool on = false;
  cudaStream_t st = nullptr;
  void begin(int cat) {
    if (!on) return;
    Span s; s.cat = cat;
    cudaEventCreate(&s.a); cudaEventCreate(&s.b);
    cudaEventRecord(s.a, st);
    spans.push_back(s);
  }
  void end() { if (on) cudaEventRecord(spans.back().b, st); }
  void collect(float out[4]) {
    for (auto &s : spans) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, s.a, s.b);
      out[s.cat] += ms;
      cudaEventDestroy(s.a); cudaEventDestroy(s.b);
    }
    spans.clear();
  }
};

template <class K>
int persistent_grid(K kernel, int block);

// Persistent grid sizes of the current device, recomputed when "persistent_ctas_per_sm" changes
const DeviceState &device_grids();

template <class K>
int persistent_grid(K kernel, int block) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0);
  if (per_sm < 1) per_sm = 1;
  if (g_opt_ctas_per_sm > 0) per_sm = std::min(per_sm, g_opt_ctas_per_sm);
  return per_sm * std::max(1, g_sm_count);
}

const DeviceState &device_grids() {
  DeviceState &d = g_devs[g_device];
  if (d.grids_opt != g_opt_ctas_per_sm) {
    d.g_trace[0][0][0] = persistent_grid(k_trace_nearest<false, false, false>, 128);
    d.g_trace[0][0][1] = persistent_grid(k_trace_nearest<false, false, true>, 128);
    d.g_trace[0][1][0] = persistent_grid(k_trace_nearest<false, true, false>, 128);
    d.g_trace[0][1][1] = persistent_grid(k_trace_nearest<false, true, true>, 128);
    d.g_trace[1][0][0] = persistent_grid(k_trace_nearest<true, false, false>, 128);
    d.g_trace[1][0][1] = persistent_grid(k_trace_nearest<true, false, true>, 128);
    d.g_trace[1][1][0] = persistent_grid(k_trace_nearest<true, true, false>, 128);
    d.g_trace[1][1][1] = persistent_grid(k_trace_nearest<true, true, true>, 128);
    d.g_shadow[0][0] = persistent_grid(k_shadow<false, false>, 128); d.g_shadow[0][1] = persistent_grid(k_shadow<false, true>, 128);
    d.g_shadow[1][0] = persistent_grid(k_shadow<true, false>, 128); d.g_shadow[1][1] = persistent_grid(k_shadow<true, true>, 128);
    d.g_frame[0][0] = persistent_grid(k_frame<false, false>, 128); d.g_frame[0][1] = persistent_grid(k_frame<false, true>, 128);
    d.g_frame[1][0] = persistent_grid(k_frame<true, false>, 128); d.g_frame[1][1] = persistent_grid(k_frame<true, true>, 128);
    d.grids_opt = g_opt_ctas_per_sm;
  }
  return d;
}

// Runs the wavefront pipeline.  Level 0 is either generated from the camera (n0 = local pixels)
// or taken from rays already stored in the level-0 queue buffers (explicit_rays).
//
// Bounded depth (0 <= max_depth <= kAsyncDepth): every level's buffers are sized for n0 rays up
// front, ray / hit counts stay on the device (FrameCounts) and the frame is a fixed sequence
// "memset, (K1, K2, K3) per level, K3b per level" with no host read-back -- captured once into a CUDA
// graph and replayed (the per-frame camera / lights / output pointers travel through the FrameParams
// device buffer, written by k_set_frame just before the replay).
// Unbounded depth (the reference's default) cannot pre-allocate 64 levels, so there the host reads
// the next level's ray count after each K3 and stops at the first empty level.
constexpr int kAsyncDepth = 8;

struct FramePlan {
  bool explicit_rays, trav_stats, async;
  int n0, J, Lmax, S, depth_cap;
};

// A scene without analytic spheres and without an octree candidate filter runs the PLAIN kernel variants
bool scene_is_plain(const RtScene *sc) { return sc->dev.n_spheres == 0 && sc->dev.oct_box == nullptr; }

// K1 for one level: nearest hit of the primary rays (generated in-kernel) or of the queued rays
void launch_trace(RtScene *sc, bool primary, bool stats, const FrameParams *fpp, const LevelBufs &lv, int level, int n_param,
                  FrameCounts *fc, cudaStream_t st) {
  const bool plain = scene_is_plain(sc);
  const int grid = device_grids().g_trace[primary][stats][plain];
  sc->dev.split_min = g_opt_donate_min;
#define RT_K1(PRIM_, STATS_, PLAIN_) k_trace_nearest<PRIM_, STATS_, PLAIN_><<<grid, 128, 0, st>>>(sc->dev, fpp, lv, level, n_param, fc)
  if (primary) {
    if (stats) { if (plain) RT_K1(true, true, true); else RT_K1(true, true, false); }
    else { if (plain) RT_K1(true, false, true); else RT_K1(true, false, false); }
  } else {
    if (stats) { if (plain) RT_K1(false, true, true); else RT_K1(false, true, false); }
    else { if (plain) RT_K1(false, false, true); else RT_K1(false, false, false); }
  }
#undef RT_K1
}

// K2 for one level: gate + sample rays of every hit
void launch_shadow(RtScene *sc, const FramePlan &pl, const FrameParams *fpp, const LevelBufs &lv, int level, FrameCounts *fc,
                   cudaStream_t st, int *launches) {
  const bool plain = scene_is_plain(sc);
  sc->dev.split_min = g_opt_donate_min;
  const int grid = device_grids().g_shadow[pl.trav_stats][plain];
#define RT_K2(STATS_, PLAIN_, PASS_) k_shadow<STATS_, PLAIN_><<<grid, 128, 0, st>>>(sc->dev, fpp, lv, level, pl.J, pl.Lmax, pl.S, fc, PASS_)
  // area mode: gate rays first, then the sample rays of the hits whose gate passed (compacted: full warps)
  for (int pass = pl.S > 0 ? 1 : 0; pass <= (pl.S > 0 ? 2 : 0); ++pass) {
    if (pl.trav_stats) { if (plain) RT_K2(true, true, pass); else RT_K2(true, false, pass); }
    else { if (plain) RT_K2(false, true, pass); else RT_K2(false, false, pass); }
    *launches += 1;
  }
#undef RT_K2
}

// enqueue the launches of levels [0, depth_cap] and the folds (async mode: all of them)
int enqueue_frame_async(RtScene *sc, const FramePlan &pl, cudaStream_t st, int *launches) {
  const FrameParams *fpp = sc->frame_params.as<FrameParams>();
  FrameCounts *fc = sc->frame_counts.as<FrameCounts>();
  const int elem_blocks = std::max(1, std::min((pl.n0 + 127) / 128, g_sm_count * 16));
  CUDA_TRY(cudaMemsetAsync(fc, 0, sizeof(FrameCounts), st));
  for (int level = 0; level <= pl.depth_cap; ++level) {
    LevelBufs lv = sc->levels[level].bufs();
    LevelBufs nx = sc->levels[level + 1].bufs();
    const int n_param = level == 0 ? pl.n0 : -1;
    launch_trace(sc, level == 0 && !pl.explicit_rays, pl.trav_stats, fpp, lv, level, n_param, fc, st);
    launch_shadow(sc, pl, fpp, lv, level, fc, st, launches);
    k_shade<<<elem_blocks, 128, 0, st>>>(sc->dev, fpp, lv, nx, level, pl.J, pl.Lmax, pl.S, fc, 0);
    *launches += 2;
  }
  for (int level = pl.depth_cap - 1; level >= 0; --level) {
    k_fold<<<elem_blocks, 256, 0, st>>>(fpp, sc->levels[level].bufs(), sc->levels[level + 1].bufs(), level,
                                        level == 0 ? pl.n0 : -1, fc);
    *launches += 1;
  }
  CUDA_TRY(cudaGetLastError());
  return RT_OK;
}

// Frame graph with conditional nodes: the kernels of bounce level k+1 (and, nested inside, everything
// deeper, and the fold of level k) sit in the body of an IF node whose condition K3 of level k sets to
// "some ray was spawned".  A depth cap above the scene's natural depth then costs nothing: on the
// cube scene (natural depth 1, cap 3) a frame is 7 kernel launches instead of 15.
//   root : memset, K1_0, K2_0, K3_0, IF(h1){ body_1 }
//   body_k: K1_k, K2_k, K3_k, IF(h_{k+1}){ body_{k+1} }, fold_{k-1}
// Built by capturing the launches of each level into its (body) graph and splicing the explicit
// conditional node into the capture.  Any failure returns an error and the caller falls back to the
// plain captured graph.
int enqueue_level(RtScene *sc, const FramePlan &pl, cudaStream_t cs, int level, cudaGraphConditionalHandle next_cond) {
  const FrameParams *fpp = sc->frame_params.as<FrameParams>();
  FrameCounts *fc = sc->frame_counts.as<FrameCounts>();
  const int elem_blocks = std::max(1, std::min((pl.n0 + 127) / 128, g_sm_count * 16));
  LevelBufs lv = sc->levels[level].bufs();
  LevelBufs nx = sc->levels[level + 1].bufs();
  const int n_param = level == 0 ? pl.n0 : -1;
  int launches = 0;
  launch_trace(sc, level == 0 && !pl.explicit_rays, pl.trav_stats, fpp, lv, level, n_param, fc, cs);
  launch_shadow(sc, pl, fpp, lv, level, fc, cs, &launches);
  k_shade<<<elem_blocks, 128, 0, cs>>>(sc->dev, fpp, lv, nx, level, pl.J, pl.Lmax, pl.S, fc, next_cond);
  CUDA_TRY(cudaGetLastError());
  return RT_OK;
}

// An IF node costs about as much as three empty launches (~10 us measured), so level 1 -- non-empty in
// any scene with a reflective or transparent surface -- stays unconditional in the root graph and the
// conditionals start at level 2:
//   root  : memset, K1_0, K2_0, K3_0, K1_1, K2_1, K3_1, IF(h2){ body_2 }, fold_0
//   body_k: K1_k, K2_k, K3_k, IF(h_{k+1}){ body_{k+1} }, fold_{k-1}          (k >= 2)
constexpr int kFirstConditionalLevel = 2;

int build_level_graph(RtScene *sc, const FramePlan &pl, cudaStream_t cs, cudaGraph_t graph, int first_level) {
  const FrameParams *fpp = sc->frame_params.as<FrameParams>();
  FrameCounts *fc = sc->frame_counts.as<FrameCounts>();
  const int elem_blocks = std::max(1, std::min((pl.n0 + 127) / 128, g_sm_count * 16));
  // levels captured straight into this graph: the root takes 0 .. kFirstConditionalLevel-1, a body takes one
  const int last_level = first_level == 0 ? std::min(pl.depth_cap, kFirstConditionalLevel - 1) : first_level;
  const bool has_next = last_level < pl.depth_cap;
  cudaGraphConditionalHandle h_next = 0;
  if (has_next) CUDA_TRY(cudaGraphConditionalHandleCreate(&h_next, graph, 0, cudaGraphCondAssignDefault));
  CUDA_TRY(cudaStreamBeginCaptureToGraph(cs, graph, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  int rc = RT_OK;
  cudaGraph_t body = nullptr;
  do {
    if (first_level == 0) {
      cudaError_t e = cudaMemsetAsync(fc, 0, sizeof(FrameCounts), cs);
      if (e != cudaSuccess) { rc = fail(RT_ERR_CUDA, "memset node: %s", cudaGetErrorString(e)); break; }
    }
    for (int level = first_level; level <= last_level && rc == RT_OK; ++level)
      rc = enqueue_level(sc, pl, cs, level, level == last_level ? h_next : 0);
    if (rc) break;
    if (has_next) {
      cudaStreamCaptureStatus status;
      const cudaGraphNode_t *deps = nullptr;
      size_t ndeps = 0;
      cudaError_t e = cudaStreamGetCaptureInfo_v2(cs, &status, nullptr, nullptr, &deps, &ndeps);
      if (e != cudaSuccess) { rc = fail(RT_ERR_CUDA, "cudaStreamGetCaptureInfo: %s", cudaGetErrorString(e)); break; }
      cudaGraphNodeParams cp = {};
      cp.type = cudaGraphNodeTypeConditional;
      cp.conditional.handle = h_next;
      cp.conditional.type = cudaGraphCondTypeIf;
      cp.conditional.size = 1;
      cudaGraphNode_t cnode;
      e = cudaGraphAddNode(&cnode, graph, deps, ndeps, &cp);
      if (e != cudaSuccess) { rc = fail(RT_ERR_CUDA, "conditional node: %s", cudaGetErrorString(e)); break; }
      body = cp.conditional.phGraph_out[0];
      e = cudaStreamUpdateCaptureDependencies(cs, &cnode, 1, cudaStreamSetCaptureDependencies);
      if (e != cudaSuccess) { rc = fail(RT_ERR_CUDA, "cudaStreamUpdateCaptureDependencies: %s", cudaGetErrorString(e)); break; }
    }
    // folds owned by this graph, deepest first: a body folds the level above it, the root folds 0 .. last_level-1
    const int fold_hi = first_level == 0 ? last_level - 1 : first_level - 1;
    const int fold_lo = first_level == 0 ? 0 : first_level - 1;
    for (int k = fold_hi; k >= fold_lo; --k)
      k_fold<<<elem_blocks, 256, 0, cs>>>(fpp, sc->levels[k].bufs(), sc->levels[k + 1].bufs(), k, k == 0 ? pl.n0 : -1, fc);
  } while (false);
  cudaGraph_t same = nullptr;
  cudaError_t e = cudaStreamEndCapture(cs, &same);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(RT_ERR_CUDA, "cudaStreamEndCapture (level %d): %s", first_level, cudaGetErrorString(e));
  if (has_next) return build_level_graph(sc, pl, cs, body, last_level + 1);
  return RT_OK;
}

int build_conditional_frame_graph(RtScene *sc, const FramePlan &pl, cudaGraphExec_t *exec_out) {
  cudaGraph_t graph = nullptr;
  CUDA_TRY(cudaGraphCreate(&graph, 0));
  int rc = build_level_graph(sc, pl, sc->stream, graph, 0);
  if (rc == RT_OK) {
    cudaError_t e = cudaGraphInstantiate(exec_out, graph, 0);
    if (e != cudaSuccess) rc = fail(RT_ERR_CUDA, "cudaGraphInstantiate (conditional): %s", cudaGetErrorString(e));
  }
  cudaGraphDestroy(graph);
  return rc;
}

// ---------------------------------------------------------------------------------------------
// fused frame: one memset + one cooperative launch of k_frame (rt_frame.cuh)
// ---------------------------------------------------------------------------------------------
constexpr int RT_RETRY_WAVEFRONT = 1;  // run_fused: the frame must be rendered by the wavefront path instead

struct FusedShape { int J, Lmax, S, depth_cap; bool unbounded; };

// Which frames take the fused kernel.  It needs the shadow jobs of a hit in a 64-bit mask and a depth cap within
// its slabs.  Whether it is also FASTER depends on the frame: a warp of k_frame walks its tile through every
// stage on its own, which costs nothing to launch and keeps everything in registers, but the warps of an SM then
// sit in different loops of a 100 KB kernel (instruction-cache misses are its top stall) and a tile's shadow jobs
// run one after the other instead of across the machine.  Measured (B200): the 1-GPU 1080p headline frame takes
// 0.40 ms fused and 0.335 ms as a wavefront graph, but a quarter of that frame (4 GPUs) takes 0.10 ms fused and
// 0.157 ms as a graph, whose ten dependent launches no longer shrink with the frame.  So the default is "auto":
// small frames -- the bands of a multi-GPU frame, interactive previews -- and frames of unbounded depth (whose
// wavefront form needs a host read-back per level) are fused, large single-GPU frames are not.
bool fused_shape(const FrameParams &fp, long long n_rays, bool plain_scene, FusedShape *out) {
  FusedShape f;
  f.Lmax = std::max(1, fp.n_lights);
  f.S = fp.point_light ? 0 : fp.usteps * fp.vsteps;
  f.J = f.Lmax + f.Lmax * f.S;
  f.unbounded = fp.max_depth < 0;
  f.depth_cap = f.unbounded ? RT_FUSED_MAX_LEVELS - 1 : fp.max_depth;
  if (out) *out = f;
  if (!g_opt_fused || f.J > RT_FUSED_MAX_JOBS || f.depth_cap > RT_FUSED_MAX_LEVELS - 1) return false;
  // (PLAIN scenes -- no spheres, no octree filter, e.g. the bundled cube -- have a much smaller fused kernel, which
  // beats the wavefront graph at every frame size: 1080p headline frame 0.276 vs 0.301 ms)
  if (g_opt_fused == 2 && !f.unbounded && !plain_scene && n_rays > 1000LL * g_opt_fused_max_kpix) return false;
  return true;
}

int fused_reserve(RtScene *sc, size_t n0, int depth_cap) {
  const size_t n = std::max<size_t>(n0, 1);
  const size_t q_slabs = (size_t)depth_cap + 1, r_slabs = (size_t)std::max(depth_cap, 1);
  int rc;
  if ((rc = sc->fq_o.reserve(q_slabs * n * 16)) || (rc = sc->fq_d.reserve(q_slabs * n * 16)) ||
      (rc = sc->fq_x.reserve(q_slabs * n * 16)) || (rc = sc->frec_a.reserve(r_slabs * n * 16)) ||
      (rc = sc->frec_b.reserve(r_slabs * n * 8)) || (rc = sc->fcounts.reserve(sizeof(FusedCounts))))
    return rc;
  return RT_OK;
}

int run_fused(RtScene *sc, FrameParams &fp, bool explicit_rays, int n0, cudaStream_t st, bool own_stream, RtStats *stats) {
  FusedShape sh;
  fused_shape(fp, n0, scene_is_plain(sc), &sh);
  int rc;
  if ((rc = fused_reserve(sc, (size_t)n0, sh.depth_cap))) return rc;
  FusedBufs fb;
  fb.q_o = sc->fq_o.as<float4>(); fb.q_d = sc->fq_d.as<float4>(); fb.q_x = sc->fq_x.as<float4>();
  fb.rec_a = sc->frec_a.as<float4>(); fb.rec_b = sc->frec_b.as<int2>();
  fb.n_cap = std::max(n0, 1);
  FusedCounts *fc = sc->fcounts.as<FusedCounts>();
  const bool trav_stats = g_opt_stats != 0;
  const bool want_stats = stats != nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  if (want_stats) { cudaEventCreate(&ev_a); cudaEventCreate(&ev_b); }
  CUDA_TRY(cudaMemsetAsync(fc, 0, sizeof(FusedCounts), st));
  if (want_stats) cudaEventRecord(ev_a, st);
  int explicit0 = explicit_rays ? 1 : 0, J = sh.J, Lmax = sh.Lmax, S = sh.S, depth_cap = sh.depth_cap, cont_min = g_opt_cont_min;
  void *args[] = {&sc->dev, &fp, &fb, &fc, &n0, &explicit0, &J, &Lmax, &S, &depth_cap, &cont_min};
  const bool plain = scene_is_plain(sc);
  const void *kern = trav_stats ? (plain ? (const void *)k_frame<true, true> : (const void *)k_frame<true, false>)
                                : (plain ? (const void *)k_frame<false, true> : (const void *)k_frame<false, false>);
  CUDA_TRY(cudaLaunchCooperativeKernel(kern, dim3(device_grids().g_frame[trav_stats][plain]), dim3(128), args, 0, st));
  if (want_stats) cudaEventRecord(ev_b, st);
  if (want_stats || sh.unbounded) {
    CUDA_TRY(cudaMemcpyAsync(sc->h_fcounts, fc, sizeof(FusedCounts), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const FusedCounts &h = *sc->h_fcounts;
    if (sh.unbounded && h.overflow) {
      // some chain wanted to go deeper than the slabs allow: the frame is incomplete
      if (want_stats) { cudaEventDestroy(ev_a); cudaEventDestroy(ev_b); }
      return RT_RETRY_WAVEFRONT;
    }
    if (want_stats) {
      memset(stats, 0, sizeof(*stats));
      cudaEventElapsedTime(&stats->ms_total, ev_a, ev_b);
      cudaEventDestroy(ev_a); cudaEventDestroy(ev_b);
      stats->rays_primary = n0;
      stats->pixels = n0;
      stats->rays_shadow = (int64_t)h.ctr.shadow_rays;
      stats->shadow_rays_traced = (int64_t)h.ctr.shadow_rays_traced;
      int64_t secondary = 0;
      int lv_used = 1;
      for (int l = 1; l <= RT_FUSED_MAX_LEVELS; ++l) { secondary += h.n_spawn[l]; if (h.n_spawn[l] > 0) lv_used = l + 1; }
      stats->rays_secondary = secondary;
      stats->levels = lv_used;
      stats->kernel_launches = 1;
      stats->box_tests = (int64_t)h.ctr.box_tests;
      stats->tri_tests = (int64_t)h.ctr.tri_tests;
      stats->shade_samples = (int64_t)h.ctr.shade_samples;
      stats->box_tests_shadow = (int64_t)h.ctr.box_tests_k2;
      stats->tri_tests_shadow = (int64_t)h.ctr.tri_tests_k2;
      stats->filter_checks = (int64_t)h.ctr.filter_checks;
      stats->filter_slow = (int64_t)h.ctr.filter_slow;
      stats->filter_rejects = (int64_t)h.ctr.filter_rejects;
      // phase shares: warp-cycles (clock64) inside the kernel, scaled to its CUDA-event duration
      const double all = (double)h.phase_clk[3];
      if (all > 0.0) {
        stats->ms_trace = (float)(stats->ms_total * (double)h.phase_clk[0] / all);
        stats->ms_shadow = (float)(stats->ms_total * (double)h.phase_clk[1] / all);
        stats->ms_shade = (float)(stats->ms_total * (double)h.phase_clk[2] / all);
      }
      stats->fused = 1;
    }
  }
  if (own_stream) CUDA_TRY(cudaStreamSynchronize(st));
  return RT_OK;
}

int run_pipeline(RtScene *sc, const FrameParams &fp_in, bool explicit_rays, int n0, uchar4 *d_rgba, int32_t *d_face,
                 float *d_t, float *d_rgbf, cudaStream_t user_stream, RtStats *stats) {
  // the legacy default stream cannot be captured: run on the scene's own stream and join at the end
  const bool own_stream = user_stream == nullptr;
  cudaStream_t st = own_stream ? sc->stream : user_stream;
  if (own_stream) CUDA_TRY(cudaDeviceSynchronize());  // order after whatever the caller queued on stream 0

  FrameParams fp = fp_in;
  fp.out_rgba = d_rgba; fp.out_face = d_face; fp.out_t = d_t; fp.out_rgbf = d_rgbf;

  if (fused_shape(fp, n0, scene_is_plain(sc), nullptr)) {
    const int frc = run_fused(sc, fp, explicit_rays, n0, st, own_stream, stats);
    if (frc != RT_RETRY_WAVEFRONT) return frc;
  }

  const bool want_stats = stats != nullptr;
  FramePlan pl;
  pl.explicit_rays = explicit_rays;
  pl.trav_stats = g_opt_stats != 0;
  pl.n0 = n0;
  pl.Lmax = std::max(1, fp.n_lights);
  pl.S = fp.point_light ? 0 : fp.usteps * fp.vsteps;
  pl.J = pl.Lmax + pl.Lmax * pl.S;
  if ((unsigned long long)n0 * (unsigned long long)pl.J >= 0xffffffffull)
    return fail(RT_ERR_LIMIT, "%d rays x %d shadow jobs exceed 2^32; render the frame in bands", n0, pl.J);
  pl.async = fp.max_depth >= 0 && fp.max_depth <= kAsyncDepth;
  pl.depth_cap = fp.max_depth >= 0 ? fp.max_depth : fp.guard_depth;
  if (pl.depth_cap + 2 > RT_MAX_LEVELS) return fail(RT_ERR_LIMIT, "max_depth %d too large", fp.max_depth);

  EventTimer timer; timer.on = want_stats; timer.st = st;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  if (want_stats) { cudaEventCreate(&ev_a); cudaEventCreate(&ev_b); cudaEventRecord(ev_a, st); }

  FrameCounts *fc = sc->frame_counts.as<FrameCounts>();
  FrameParams *fpp = sc->frame_params.as<FrameParams>();
  int launches = 0, rc;
  k_set_frame<<<1, 128, 0, st>>>(fp, fpp);
  ++launches;

  if (pl.async && !want_stats) {
    // ---- bounded depth: CUDA graph replay ----
    if ((int)sc->levels.size() < pl.depth_cap + 2) sc->levels.resize(pl.depth_cap + 2);
    for (int l = 0; l <= pl.depth_cap; ++l)
      if ((rc = sc->levels[l].reserve((size_t)n0, (size_t)pl.J))) return rc;
    const std::vector<long long> key = {g_alloc_generation.load(), n0, pl.J, pl.Lmax, pl.S, pl.depth_cap, g_opt_graph_cond, g_opt_ctas_per_sm, g_opt_donate_min,
                                        (long long)pl.explicit_rays, (long long)pl.trav_stats};
    if (sc->graph_exec == nullptr || key != sc->graph_key) {
      if (sc->graph_exec) { cudaGraphExecDestroy(sc->graph_exec); sc->graph_exec = nullptr; }
      if (g_opt_graph_cond && pl.depth_cap >= kFirstConditionalLevel) {
        if (build_conditional_frame_graph(sc, pl, &sc->graph_exec) == RT_OK) sc->graph_key = key;
        else { cudaGetLastError(); sc->graph_exec = nullptr; }  // fall back to the plain captured graph below
      }
    }
    if (sc->graph_exec == nullptr || key != sc->graph_key) {
      cudaGraph_t graph = nullptr;
      CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      rc = enqueue_frame_async(sc, pl, st, &launches);
      cudaError_t e = cudaStreamEndCapture(st, &graph);
      if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
      if (e != cudaSuccess) return fail(RT_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
      e = cudaGraphInstantiate(&sc->graph_exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) { sc->graph_exec = nullptr; return fail(RT_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e)); }
      sc->graph_key = key;
    }
    CUDA_TRY(cudaGraphLaunch(sc->graph_exec, st));
    if (own_stream) CUDA_TRY(cudaStreamSynchronize(st));
    return RT_OK;
  }

  // ---- stream launches (statistics requested, or unbounded depth with host-visible ray counts) ----
  if (pl.async) {
    if ((int)sc->levels.size() < pl.depth_cap + 2) sc->levels.resize(pl.depth_cap + 2);
    for (int l = 0; l <= pl.depth_cap; ++l)
      if ((rc = sc->levels[l].reserve((size_t)n0, (size_t)pl.J))) return rc;
  } else {
    if (sc->levels.size() < 2) sc->levels.resize(2);
    if ((rc = sc->levels[0].reserve((size_t)n0, (size_t)pl.J))) return rc;
  }
  CUDA_TRY(cudaMemsetAsync(fc, 0, sizeof(FrameCounts), st));
  const int elem_blocks = std::max(1, std::min((n0 + 127) / 128, g_sm_count * 16));
  int levels_run = 0;
  int cur_n = n0;  // host knowledge of the level's ray count (exact in sync mode, upper bound in async mode)
  for (int level = 0; level <= pl.depth_cap; ++level) {
    LevelBufs lv = sc->levels[level].bufs();
    const int n_param = level == 0 ? n0 : -1;
    timer.begin(0);
    launch_trace(sc, level == 0 && !explicit_rays, pl.trav_stats, fpp, lv, level, n_param, fc, st);
    timer.end();
    timer.begin(1);
    launch_shadow(sc, pl, fpp, lv, level, fc, st, &launches);
    timer.end();
    const bool may_spawn = level < pl.depth_cap;
    if (!pl.async) {
      if ((int)sc->levels.size() < level + 2) sc->levels.resize(level + 2);
      if (may_spawn && (rc = sc->levels[level + 1].reserve((size_t)cur_n, (size_t)pl.J))) return rc;
    }
    LevelBufs nx = sc->levels[std::min(level + 1, (int)sc->levels.size() - 1)].bufs();
    timer.begin(2);
    k_shade<<<elem_blocks, 128, 0, st>>>(sc->dev, fpp, lv, nx, level, pl.J, pl.Lmax, pl.S, fc, 0);
    timer.end();
    launches += 2;
    levels_run = level + 1;
    CUDA_TRY(cudaGetLastError());
    if (!may_spawn) break;
    if (!pl.async) {
      CUDA_TRY(cudaMemcpyAsync(&sc->h_counts->n_rays[level + 1], &fc->n_rays[level + 1], 4, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      cur_n = sc->h_counts->n_rays[level + 1];
      if (cur_n <= 0) break;
    }
  }
  for (int level = levels_run - 2; level >= 0; --level) {
    timer.begin(2);
    k_fold<<<elem_blocks, 256, 0, st>>>(fpp, sc->levels[level].bufs(), sc->levels[level + 1].bufs(), level,
                                        level == 0 ? n0 : -1, fc);
    timer.end();
    ++launches;
  }
  CUDA_TRY(cudaGetLastError());

  if (want_stats) {
    cudaEventRecord(ev_b, st);
    CUDA_TRY(cudaMemcpyAsync(sc->h_counts, fc, sizeof(FrameCounts), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    memset(stats, 0, sizeof(*stats));
    float cat[4] = {0, 0, 0, 0};
    timer.collect(cat);
    cudaEventElapsedTime(&stats->ms_total, ev_a, ev_b);
    cudaEventDestroy(ev_a); cudaEventDestroy(ev_b);
    const FrameCounts &h = *sc->h_counts;
    stats->ms_trace = cat[0]; stats->ms_shadow = cat[1]; stats->ms_shade = cat[2];
    stats->rays_primary = n0;
    stats->pixels = n0;
    stats->rays_shadow = (int64_t)h.ctr.shadow_rays;
    stats->shadow_rays_traced = (int64_t)h.ctr.shadow_rays_traced;
    int64_t secondary = 0;
    int lv_used = 1;
    for (int l = 1; l < RT_MAX_LEVELS; ++l) { secondary += h.n_rays[l]; if (h.n_rays[l] > 0) lv_used = l + 1; }
    stats->rays_secondary = secondary;
    stats->levels = lv_used;
    stats->kernel_launches = launches;
    stats->fused = pl.async ? 0 : 2;  // 2: per-level host read-back (unbounded depth, or a cap above the graph's 8 levels)
    stats->box_tests = (int64_t)h.ctr.box_tests;
    stats->tri_tests = (int64_t)h.ctr.tri_tests;
    stats->shade_samples = (int64_t)h.ctr.shade_samples;
    stats->box_tests_shadow = (int64_t)h.ctr.box_tests_k2;
    stats->tri_tests_shadow = (int64_t)h.ctr.tri_tests_k2;
    stats->filter_checks = (int64_t)h.ctr.filter_checks;
    stats->filter_slow = (int64_t)h.ctr.filter_slow;
    stats->filter_rejects = (int64_t)h.ctr.filter_rejects;
  }
  if (own_stream) CUDA_TRY(cudaStreamSynchronize(st));
  return RT_OK;
}

}  // namespace

extern "C" int rt_render_device(RtScene *sc, const RtCamera *cam, const RtLights *lights, const RtParams *p,
                                void *d_rgba, int32_t *d_face, float *d_t, float *d_rgb_f32, void *stream,
                                RtStats *stats) {
  if (!sc || !cam || !lights || !p || !d_rgba) return fail(RT_ERR_INVALID, "null argument");
  if (p->width <= 0 || p->height <= 0) return fail(RT_ERR_INVALID, "bad image size %dx%d", p->width, p->height);
  if (p->band_world > 1 && (p->band_rank < 0 || p->band_rank >= p->band_world)) return fail(RT_ERR_INVALID, "bad band rank");
  int rc = use_device(sc->device);
  if (rc) return rc;
  FrameParams fp;
  if ((rc = fill_frame(fp, cam, lights, p))) return rc;
  set_tile_rect(fp, sc);
  const long long n0 = (long long)fp.local_rows * fp.width;
  if (n0 > 0x7fffffffLL / 32) return fail(RT_ERR_LIMIT, "image too large for one call (%lld pixels); shard it in bands", n0);
  if (n0 == 0) return RT_OK;
  return run_pipeline(sc, fp, false, (int)n0, (uchar4 *)d_rgba, d_face, d_t, d_rgb_f32, (cudaStream_t)stream, stats);
}

extern "C" int rt_render(RtScene *sc, const RtCamera *cam, const RtLights *lights, const RtParams *p, uint8_t *rgba_out,
                         int32_t *face_out, float *t_out, float *rgb_f32_out, RtStats *stats) {
  if (!sc || !p || !rgba_out) return fail(RT_ERR_INVALID, "null argument");
  int rc = use_device(sc->device);
  if (rc) return rc;
  const size_t n = (size_t)rt_local_rows(p) * (size_t)std::max(0, p->width);
  if ((rc = sc->out_rgba.reserve(n * 4))) return rc;
  if (face_out && (rc = sc->out_face.reserve(n * 4))) return rc;
  if (t_out && (rc = sc->out_t.reserve(n * 4))) return rc;
  if (rgb_f32_out && (rc = sc->out_rgbf.reserve(n * 12))) return rc;
  RtParams local_p = *p;
  local_p.out_full_frame = 0;  // host outputs always hold this call's rows only
  p = &local_p;
  // everything on the scene's own stream: frame, copies, one synchronisation at the end
  cudaStream_t st = sc->stream;
  // A page-locked host frame (cudaHostAlloc / cudaHostRegister, e.g. a pinned tensor) is mapped into the device's
  // address space: the kernels that finish a pixel store its uchar4 straight into it, the 8.3 MB of a 1080p frame
  // cross PCIe as posted writes while the frame is still being rendered, and no device->host copy follows.
  if (g_opt_host_direct && n > 0 && (g_opt_host_direct == 2 || n * 4 <= (size_t)g_opt_direct_max_mb << 20)) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, rgba_out) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer != nullptr) {
      g_tile_override = g_opt_direct_tile_w_log2;
      rc = rt_render_device(sc, cam, lights, p, at.devicePointer, face_out ? sc->out_face.as<int32_t>() : nullptr,
                            t_out ? sc->out_t.as<float>() : nullptr, rgb_f32_out ? sc->out_rgbf.as<float>() : nullptr, st, stats);
      g_tile_override = 0;
      if (rc) return rc;
      if (face_out) CUDA_TRY(cudaMemcpyAsync(face_out, sc->out_face.p, n * 4, cudaMemcpyDeviceToHost, st));
      if (t_out) CUDA_TRY(cudaMemcpyAsync(t_out, sc->out_t.p, n * 4, cudaMemcpyDeviceToHost, st));
      if (rgb_f32_out) CUDA_TRY(cudaMemcpyAsync(rgb_f32_out, sc->out_rgbf.p, n * 12, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      return RT_OK;
    }
    cudaGetLastError();  // (an unregistered host pointer is reported as an error by older drivers)
  }
  // Only the packed frame is wanted, the whole image is rendered by this call and the scene takes the fused kernel
  // at any size: render it in row chunks (one launch each) and copy chunk c to the host while chunk c+1 is being
  // rendered -- the 8.3 MB of a 1080p frame otherwise add 0.17 ms of PCIe time behind a 0.28 ms frame.
  // (Large frames of any scene -- from 32 MB on, e.g. the 8K frame -- are chunked the same way through the wavefront
  // kernels when the rows divide evenly, so that every chunk replays the same graph.  Not below: the 4K frame of C4,
  // 31.6 MB and six bounce levels deep, took 13.6 instead of 9.8 ms in four chunks.)
  const bool big_frame = n * 4 >= ((size_t)32 << 20) && p->height % 16 == 0;
  if (g_opt_chunks > 1 && !face_out && !t_out && !rgb_f32_out && !stats && cam && lights && p->band_world <= 1 &&
      p->height >= 64 * g_opt_chunks && p->width > 0 && (scene_is_plain(sc) || big_frame)) {
    const int C = scene_is_plain(sc) ? g_opt_chunks : 4, H = p->height, W = p->width;
    const int B = (((H + C - 1) / C) + 3) & ~3;  // rows per chunk, a multiple of the 8x4 tile height
    FrameParams probe;
    if ((rc = fill_frame(probe, cam, lights, p))) return rc;
    if (fused_shape(probe, (long long)B * W, scene_is_plain(sc), nullptr) || big_frame) {
      if (!sc->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&sc->copy_stream, cudaStreamNonBlocking));
      for (int c = 0; c < C; ++c)
        if (!sc->ev_chunk[c]) CUDA_TRY(cudaEventCreateWithFlags(&sc->ev_chunk[c], cudaEventDisableTiming));
      for (int c = 0; c < C; ++c) {
        RtParams cp = *p;
        cp.band_rows = B; cp.band_rank = c; cp.band_world = C; cp.out_full_frame = 1;
        const int rows = rt_local_rows(&cp);
        if (rows <= 0) continue;
        if ((rc = rt_render_device(sc, cam, lights, &cp, sc->out_rgba.p, nullptr, nullptr, nullptr, st, nullptr))) return rc;
        CUDA_TRY(cudaEventRecord(sc->ev_chunk[c], st));
        CUDA_TRY(cudaStreamWaitEvent(sc->copy_stream, sc->ev_chunk[c], 0));
        const size_t off = (size_t)c * (size_t)B * (size_t)W * 4;
        CUDA_TRY(cudaMemcpyAsync(rgba_out + off, (const uint8_t *)sc->out_rgba.p + off, (size_t)rows * (size_t)W * 4,
                                 cudaMemcpyDeviceToHost, sc->copy_stream));
      }
      CUDA_TRY(cudaStreamSynchronize(sc->copy_stream));
      return RT_OK;
    }
  }
  rc = rt_render_device(sc, cam, lights, p, sc->out_rgba.p, face_out ? sc->out_face.as<int32_t>() : nullptr,
                        t_out ? sc->out_t.as<float>() : nullptr, rgb_f32_out ? sc->out_rgbf.as<float>() : nullptr,
                        st, stats);
  if (rc) return rc;
  if (n == 0) return RT_OK;
  CUDA_TRY(cudaMemcpyAsync(rgba_out, sc->out_rgba.p, n * 4, cudaMemcpyDeviceToHost, st));
  if (face_out) CUDA_TRY(cudaMemcpyAsync(face_out, sc->out_face.p, n * 4, cudaMemcpyDeviceToHost, st));
  if (t_out) CUDA_TRY(cudaMemcpyAsync(t_out, sc->out_t.p, n * 4, cudaMemcpyDeviceToHost, st));
  if (rgb_f32_out) CUDA_TRY(cudaMemcpyAsync(rgb_f32_out, sc->out_rgbf.p, n * 12, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return RT_OK;
}

extern "C" int rt_render_submit(RtScene *sc, const RtCamera *cam, const RtLights *lights, const RtParams *p,
                                uint8_t *rgba_out, int *ticket) {
  if (!sc || !p || !rgba_out || !ticket) return fail(RT_ERR_INVALID, "null argument");
  int rc = use_device(sc->device);
  if (rc) return rc;
  const int slot = sc->submit_seq & 1;
  if (sc->slot_busy[slot])
    return fail(RT_ERR_INVALID, "two frames already in flight: rt_render_wait(ticket %d) first", sc->submit_seq - 2);
  if (!sc->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&sc->copy_stream, cudaStreamNonBlocking));
  for (int k = 0; k < 2; ++k) {
    if (!sc->ev_rendered[k]) CUDA_TRY(cudaEventCreateWithFlags(&sc->ev_rendered[k], cudaEventDisableTiming));
    if (!sc->ev_copied[k]) CUDA_TRY(cudaEventCreateWithFlags(&sc->ev_copied[k], cudaEventDisableTiming));
  }
  const size_t n = (size_t)rt_local_rows(p) * (size_t)std::max(0, p->width);
  if ((rc = sc->slot_rgba[slot].reserve(std::max<size_t>(n, 1) * 4))) return rc;
  RtParams local_p = *p;
  local_p.out_full_frame = 0;
  rc = rt_render_device(sc, cam, lights, &local_p, sc->slot_rgba[slot].p, nullptr, nullptr, nullptr, sc->stream, nullptr);
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(sc->ev_rendered[slot], sc->stream));
  CUDA_TRY(cudaStreamWaitEvent(sc->copy_stream, sc->ev_rendered[slot], 0));
  if (n) CUDA_TRY(cudaMemcpyAsync(rgba_out, sc->slot_rgba[slot].p, n * 4, cudaMemcpyDeviceToHost, sc->copy_stream));
  CUDA_TRY(cudaEventRecord(sc->ev_copied[slot], sc->copy_stream));
  sc->slot_busy[slot] = true;
  *ticket = sc->submit_seq++;
  return RT_OK;
}

extern "C" int rt_render_wait(RtScene *sc, int ticket) {
  if (!sc) return fail(RT_ERR_INVALID, "null scene");
  if (use_device(sc->device)) return RT_ERR_NO_DEVICE;
  if (ticket < 0 || ticket >= sc->submit_seq) return fail(RT_ERR_INVALID, "unknown ticket %d", ticket);
  if (ticket < sc->submit_seq - 2 || !sc->slot_busy[ticket & 1]) return RT_OK;  // already complete
  if (ticket == sc->submit_seq - 1 && sc->slot_busy[(ticket & 1) ^ 1]) {
    // frames complete in submission order: finishing the newer one finishes the older one too
    CUDA_TRY(cudaEventSynchronize(sc->ev_copied[(ticket & 1) ^ 1]));
    sc->slot_busy[(ticket & 1) ^ 1] = false;
  }
  CUDA_TRY(cudaEventSynchronize(sc->ev_copied[ticket & 1]));
  sc->slot_busy[ticket & 1] = false;
  return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// cross-process framebuffer sharing (one process per GPU): rank 0 allocates the frame, the other ranks
// map it over NVLink peer access and their kernels store pixels straight into it
// ---------------------------------------------------------------------------------------------
extern "C" int rt_shared_frame_create(size_t bytes, void **d_ptr, unsigned char handle[64]) {
  if (!d_ptr || !handle || bytes == 0) return fail(RT_ERR_INVALID, "bad argument");
  int rc = ensure_device();
  if (rc) return rc;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  // the frame, then 4 KiB of per-rank completion flags (rt_shared_frame_signal / rt_shared_frame_wait)
  const size_t frame_al = (bytes + 255) & ~(size_t)255;
  CUDA_TRY(cudaMalloc(d_ptr, frame_al + 4096));
  CUDA_TRY(cudaMemset((char *)*d_ptr + frame_al, 0, 4096));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, *d_ptr);
  if (e != cudaSuccess) { cudaFree(*d_ptr); *d_ptr = nullptr; return fail(RT_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
  memcpy(handle, &h, 64);
  return RT_OK;
}

extern "C" int rt_shared_frame_open(const unsigned char handle[64], void **d_ptr) {
  if (!d_ptr || !handle) return fail(RT_ERR_INVALID, "bad argument");
  int rc = ensure_device();
  if (rc) return rc;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  CUDA_TRY(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return RT_OK;
}

extern "C" int rt_shared_frame_close(void *d_ptr, int owner) {
  if (!d_ptr) return RT_OK;
  if (owner) CUDA_TRY(cudaFree(d_ptr));
  else CUDA_TRY(cudaIpcCloseMemHandle(d_ptr));
  return RT_OK;
}

extern "C" int rt_device_copy_to_host(void *host, const void *d_ptr, size_t bytes) {
  if (!host || !d_ptr) return fail(RT_ERR_INVALID, "null argument");
  CUDA_TRY(cudaMemcpy(host, d_ptr, bytes, cudaMemcpyDeviceToHost));
  return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// batched per-function entry points
// ---------------------------------------------------------------------------------------------
extern "C" int rt_trace_rays(RtScene *sc, int64_t n, const float *origins, const float *dirs, const RtLights *lights,
                             const RtParams *p, float *rgb_out, int32_t *face_out, float *t_out) {
  if (!sc || !origins || !dirs || !lights || !p || !rgb_out) return fail(RT_ERR_INVALID, "null argument");
  if (n <= 0) return RT_OK;
  if (n > (1 << 26)) return fail(RT_ERR_LIMIT, "at most 2^26 rays per call");
  int rc = use_device(sc->device);
  if (rc) return rc;
  FrameParams fp;
  if ((rc = fill_frame(fp, nullptr, lights, p))) return rc;
  fp.width = (int)n; fp.height = 1; fp.local_rows = 1; fp.band_world = 1;
  const int Lmax = std::max(1, fp.n_lights);
  const int S = fp.point_light ? 0 : fp.usteps * fp.vsteps;
  std::vector<float> ho((size_t)n * 4), hd((size_t)n * 4);
  for (int64_t i = 0; i < n; ++i) {
    for (int k = 0; k < 3; ++k) { ho[4 * i + k] = origins[3 * i + k]; hd[4 * i + k] = dirs[3 * i + k]; }
    ho[4 * i + 3] = bits(0); hd[4 * i + 3] = 0.f;
  }
  FusedShape fsh;
  const bool fused = fused_shape(fp, n, scene_is_plain(sc), &fsh);
  if (fused) {
    // level-0 slab of the fused frame's ray queue: (o, flags), (d, lp.x), (lp.y, lp.z, parent = none, ray index)
    if ((rc = fused_reserve(sc, (size_t)n, fsh.depth_cap))) return rc;
    std::vector<float> hx((size_t)n * 4, 0.f);
    for (int64_t i = 0; i < n; ++i) { hx[4 * i + 2] = bits(-1); hx[4 * i + 3] = bits((int32_t)i); }
    CUDA_TRY(cudaMemcpy(sc->fq_o.p, ho.data(), (size_t)n * 16, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(sc->fq_d.p, hd.data(), (size_t)n * 16, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(sc->fq_x.p, hx.data(), (size_t)n * 16, cudaMemcpyHostToDevice));
  }
  if (!fused || fsh.unbounded) {  // the wavefront path (also the fall-back of an unbounded frame that went too deep)
    if ((int)sc->levels.size() < 1) sc->levels.resize(1);
    if ((rc = sc->levels[0].reserve((size_t)n, (size_t)(Lmax + Lmax * S)))) return rc;
    CUDA_TRY(cudaMemcpy(sc->levels[0].ray_o.p, ho.data(), (size_t)n * 16, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(sc->levels[0].ray_d.p, hd.data(), (size_t)n * 16, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemset(sc->levels[0].ray_l.p, 0, (size_t)n * 8));
  }
  if ((rc = sc->out_rgbf.reserve((size_t)n * 12)) || (rc = sc->out_face.reserve((size_t)n * 4)) ||
      (rc = sc->out_t.reserve((size_t)n * 4)))
    return rc;
  if ((rc = run_pipeline(sc, fp, true, (int)n, nullptr, sc->out_face.as<int32_t>(), sc->out_t.as<float>(),
                         sc->out_rgbf.as<float>(), 0, nullptr)))
    return rc;
  CUDA_TRY(cudaStreamSynchronize(0));
  CUDA_TRY(cudaMemcpy(rgb_out, sc->out_rgbf.p, (size_t)n * 12, cudaMemcpyDeviceToHost));
  if (face_out) CUDA_TRY(cudaMemcpy(face_out, sc->out_face.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  if (t_out) CUDA_TRY(cudaMemcpy(t_out, sc->out_t.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return RT_OK;
}

extern "C" int rt_light_strikes(RtScene *sc, int64_t n, const float *hit_points, const RtLights *lights, uint8_t *visible_out) {
  if (!sc || !hit_points || !lights || !visible_out) return fail(RT_ERR_INVALID, "null argument");
  if (n <= 0 || lights->n <= 0) return RT_OK;
  int rc = use_device(sc->device);
  if (rc) return rc;
  RtParams p; rt_default_params(&p);
  FrameParams fp;
  if ((rc = fill_frame(fp, nullptr, lights, &p))) return rc;
  if ((rc = sc->in_a.reserve((size_t)n * 12)) || (rc = sc->in_b.reserve((size_t)n * lights->n))) return rc;
  CUDA_TRY(cudaMemcpy(sc->in_a.p, hit_points, (size_t)n * 12, cudaMemcpyHostToDevice));
  const int64_t jobs = n * lights->n;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((jobs + 127) / 128, g_sm_count * 16));
  k_light_strikes<<<blocks, 128>>>(sc->dev, fp, n, sc->in_a.as<float>(), sc->in_b.as<uint8_t>());
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(visible_out, sc->in_b.p, (size_t)jobs, cudaMemcpyDeviceToHost));
  return RT_OK;
}

extern "C" int rt_box_intersect(RtScene *sc, int64_t n, const float *origins, const float *dests, uint8_t *hit_out) {
  if (!sc || !origins || !dests || !hit_out) return fail(RT_ERR_INVALID, "null argument");
  if (n <= 0) return RT_OK;
  int rc = use_device(sc->device);
  if (rc) return rc;
  if ((rc = sc->in_a.reserve((size_t)n * 24)) || (rc = sc->in_b.reserve((size_t)n))) return rc;
  float *d_o = sc->in_a.as<float>(), *d_d = d_o + (size_t)n * 3;
  CUDA_TRY(cudaMemcpy(d_o, origins, (size_t)n * 12, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(d_d, dests, (size_t)n * 12, cudaMemcpyHostToDevice));
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, g_sm_count * 16));
  k_box_intersect<<<blocks, 256>>>(sc->dev, n, d_o, d_d, sc->in_b.as<uint8_t>());
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(hit_out, sc->in_b.p, (size_t)n, cudaMemcpyDeviceToHost));
  return RT_OK;
}

extern "C" int rt_box_intersect_box(const float mn[3], const float mx[3], int64_t n, const float *origins,
                                    const float *dests, uint8_t *hit_out) {
  if (!mn || !mx || !origins || !dests || !hit_out) return fail(RT_ERR_INVALID, "null argument");
  if (n <= 0) return RT_OK;
  int rc = ensure_device();
  if (rc) return rc;
  DevBuf in, out;
  if ((rc = in.reserve((size_t)n * 24)) || (rc = out.reserve((size_t)n))) { in.release(); out.release(); return rc; }
  float *d_o = in.as<float>(), *d_d = d_o + (size_t)n * 3;
  cudaMemcpy(d_o, origins, (size_t)n * 12, cudaMemcpyHostToDevice);
  cudaMemcpy(d_d, dests, (size_t)n * 12, cudaMemcpyHostToDevice);
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, g_sm_count * 16));
  k_box_intersect_box<<<blocks, 256>>>(make_float3(mn[0], mn[1], mn[2]), make_float3(mx[0], mx[1], mx[2]), n, d_o, d_d,
                                       out.as<uint8_t>());
  cudaError_t e = cudaMemcpy(hit_out, out.p, (size_t)n, cudaMemcpyDeviceToHost);
  in.release(); out.release();
  if (e != cudaSuccess) return fail(RT_ERR_CUDA, "box_intersect_box: %s", cudaGetErrorString(e));
  return RT_OK;
}

extern "C" int rt_ray_triangle(RtScene *sc, int64_t n, const float *origins, const float *dirs, const int32_t *faces,
                               float *t_out) {
  if (!sc || !origins || !dirs || !faces || !t_out) return fail(RT_ERR_INVALID, "null argument");
  if (n <= 0) return RT_OK;
  int rc = use_device(sc->device);
  if (rc) return rc;
  if ((rc = sc->in_a.reserve((size_t)n * 28)) || (rc = sc->in_b.reserve((size_t)n * 4))) return rc;
  float *d_o = sc->in_a.as<float>(), *d_d = d_o + (size_t)n * 3;
  int32_t *d_f = (int32_t *)(d_d + (size_t)n * 3);
  CUDA_TRY(cudaMemcpy(d_o, origins, (size_t)n * 12, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(d_d, dirs, (size_t)n * 12, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(d_f, faces, (size_t)n * 4, cudaMemcpyHostToDevice));
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, g_sm_count * 16));
  k_ray_triangle<<<blocks, 256>>>(sc->dev, n, d_o, d_d, d_f, sc->in_b.as<float>());
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(t_out, sc->in_b.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return RT_OK;
}

extern "C" int rt_octree_candidates(RtScene *sc, const float origin[3], const float dest[3], int32_t *ids, int32_t cap) {
  if (!sc || !origin || !dest || (!ids && cap > 0)) return fail(RT_ERR_INVALID, "null argument");
  int rc = use_device(sc->device);
  if (rc) return rc;
  const int T = sc->dev.n_faces;
  if (T == 0) return 0;
  if ((rc = sc->in_b.reserve((size_t)T))) return rc;
  const int blocks = std::max(1, std::min((T + 255) / 256, g_sm_count * 16));
  k_octree_candidates<<<blocks, 256>>>(sc->dev, make_float3(origin[0], origin[1], origin[2]),
                                       make_float3(dest[0], dest[1], dest[2]), sc->in_b.as<uint8_t>());
  CUDA_TRY(cudaGetLastError());
  std::vector<uint8_t> flag((size_t)T);
  CUDA_TRY(cudaMemcpy(flag.data(), sc->in_b.p, (size_t)T, cudaMemcpyDeviceToHost));
  int count = 0;
  for (int f = 0; f < T; ++f)
    if (flag[f]) { if (count < cap) ids[count] = f; ++count; }
  return count;
}

extern "C" int rt_phong_shade(RtScene *sc, int64_t n, const float *origins, const float *hit_points, const int32_t *faces,
                              const RtLights *lights, const RtParams *p, float *rgb_out) {
  if (!sc || !origins || !hit_points || !faces || !lights || !p || !rgb_out) return fail(RT_ERR_INVALID, "null argument");
  if (n <= 0) return RT_OK;
  int rc = use_device(sc->device);
  if (rc) return rc;
  FrameParams fp;
  if ((rc = fill_frame(fp, nullptr, lights, p))) return rc;
  if ((rc = sc->in_a.reserve((size_t)n * 28)) || (rc = sc->in_b.reserve((size_t)n * 12))) return rc;
  float *d_o = sc->in_a.as<float>(), *d_h = d_o + (size_t)n * 3;
  int32_t *d_f = (int32_t *)(d_h + (size_t)n * 3);
  CUDA_TRY(cudaMemcpy(d_o, origins, (size_t)n * 12, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(d_h, hit_points, (size_t)n * 12, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(d_f, faces, (size_t)n * 4, cudaMemcpyHostToDevice));
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + 127) / 128, g_sm_count * 16));
  k_phong_shade<<<blocks, 128>>>(sc->dev, fp, n, d_o, d_h, d_f, sc->in_b.as<float>());
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(rgb_out, sc->in_b.p, (size_t)n * 12, cudaMemcpyDeviceToHost));
  return RT_OK;
}

extern "C" int rt_screen_to_world(const RtCamera *cam, int64_t n, const float *pixels_xy, float *out) {
  if (!cam || !pixels_xy || !out) return fail(RT_ERR_INVALID, "null argument");
  if (n <= 0) return RT_OK;
  int rc = ensure_device();
  if (rc) return rc;
  RtParams p; rt_default_params(&p);
  RtLights l{}; l.n = 0; l.pos = nullptr;
  FrameParams fp;
  if ((rc = fill_frame(fp, cam, &l, &p))) return rc;
  float *d_in = nullptr, *d_out = nullptr;
  CUDA_TRY(cudaMalloc((void **)&d_in, (size_t)n * 8));
  if (cudaMalloc((void **)&d_out, (size_t)n * 12) != cudaSuccess) { cudaFree(d_in); return fail(RT_ERR_CUDA, "cudaMalloc failed"); }
  cudaMemcpy(d_in, pixels_xy, (size_t)n * 8, cudaMemcpyHostToDevice);
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, g_sm_count * 16));
  k_screen_to_world<<<blocks, 256>>>(fp, n, d_in, d_out);
  cudaError_t e = cudaMemcpy(out, d_out, (size_t)n * 12, cudaMemcpyDeviceToHost);
  cudaFree(d_in); cudaFree(d_out);
  if (e != cudaSuccess) return fail(RT_ERR_CUDA, "screen_to_world: %s", cudaGetErrorString(e));
  return RT_OK;
}

// Host-side restatement of the light sample positions handed to the shadow kernel (the device
// computes the same expression per job; this entry point serves Flyscene::createSpherePoint).
extern "C" int rt_light_samples(const RtParams *p, const float light[3], float *out) {
  if (!p || !light || !out) return fail(RT_ERR_INVALID, "null argument");
  if (p->point_light) { out[0] = light[0]; out[1] = light[1]; out[2] = light[2]; return 1; }
  if (!p->area_light) {
    float off[RT_SPHERE_SAMPLES * 3];
    sphere_offsets(p->sphere_seed, p->sphere_radius, off);
    for (int k = 0; k < RT_SPHERE_SAMPLES * 3; ++k) out[k] = off[k] + light[k % 3];
    return RT_SPHERE_SAMPLES;
  }
  if (p->usteps * p->vsteps > RT_MAX_SAMPLES) return fail(RT_ERR_LIMIT, "too many samples");
  const float ux = light[0] + p->area_len_x * 1.0f, vy = light[1] + p->area_len_y * 1.0f, uz = light[2] + p->area_len_x * 0.0f;
  int k = 0;
  for (int i = 0; i < p->usteps; ++i)
    for (int j = 0; j < p->vsteps; ++j) {
      out[3 * k] = (float)((double)i + 0.5) * (ux / (float)p->usteps);
      out[3 * k + 1] = (float)((double)j + 0.5) * (vy / (float)p->vsteps);
      out[3 * k + 2] = uz;
      ++k;
    }
  return k;
}

// Device self-test: div3 (the shared-reciprocal normalisation of the shading kernels) against IEEE division
extern "C" int rt_selftest_div3(int64_t n_trials, uint32_t seed, int32_t exp_range, int64_t *mismatches) {
  if (!mismatches || n_trials < 0) return fail(RT_ERR_INVALID, "bad argument");
  int rc = ensure_device();
  if (rc) return rc;
  unsigned long long *d_bad = nullptr;
  CUDA_TRY(cudaMalloc(&d_bad, 8));
  cudaMemset(d_bad, 0, 8);
  k_selftest_div3<<<g_sm_count * 8, 256>>>((unsigned long long)n_trials, seed, exp_range, d_bad);
  unsigned long long h = 0;
  cudaError_t e = cudaMemcpy(&h, d_bad, 8, cudaMemcpyDeviceToHost);
  cudaFree(d_bad);
  if (e != cudaSuccess) return fail(RT_ERR_CUDA, "selftest: %s", cudaGetErrorString(e));
  *mismatches = (int64_t)h;
  return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// output: writePPMImage, tucano/utils/ppmIO.hpp:130-151
// ---------------------------------------------------------------------------------------------
extern "C" int rt_write_ppm(const char *path, const uint8_t *rgba, int32_t width, int32_t height, int32_t binary) {
  if (!path || !rgba || width <= 0 || height <= 0) return fail(RT_ERR_INVALID, "bad argument");
  FILE *f = fopen(path, "wb");
  if (!f) return fail(RT_ERR_IO, "cannot open %s for writing", path);
  if (binary) {
    fprintf(f, "P6\n%d %d\n255\n", width, height);
    std::vector<uint8_t> row((size_t)width * 3);
    for (int j = 0; j < height; ++j) {
      for (int i = 0; i < width; ++i) {
        const uint8_t *px = rgba + ((size_t)j * width + i) * 4;
        row[3 * i] = px[0]; row[3 * i + 1] = px[1]; row[3 * i + 2] = px[2];
      }
      fwrite(row.data(), 1, row.size(), f);
    }
  } else {
    // "r g b " per pixel, newline per row, exactly the reference's text layout.  Every value 0..255
    // is pre-formatted once ("123 "), a pixel is three table copies; rows go out in 4 MiB chunks.
    char tab[256][4];
    uint8_t len[256];
    for (int v = 0; v < 256; ++v) {
      int n = 0;
      if (v >= 100) tab[v][n++] = (char)('0' + v / 100);
      if (v >= 10) tab[v][n++] = (char)('0' + (v / 10) % 10);
      tab[v][n++] = (char)('0' + v % 10);
      tab[v][n++] = ' ';
      len[v] = (uint8_t)n;
    }
    fprintf(f, "P3\n%d %d\n255\n", width, height);
    std::vector<char> buf((size_t)4 << 20);
    const size_t row_max = (size_t)width * 12 + 1;
    if (buf.size() < row_max) buf.resize(row_max);
    size_t used = 0;
    for (int j = 0; j < height; ++j) {
      if (used + row_max > buf.size()) { fwrite(buf.data(), 1, used, f); used = 0; }
      char *o = buf.data() + used;
      const uint8_t *px = rgba + (size_t)j * width * 4;
      for (int i = 0; i < width; ++i, px += 4) {
        memcpy(o, tab[px[0]], 4); o += len[px[0]];
        memcpy(o, tab[px[1]], 4); o += len[px[1]];
        memcpy(o, tab[px[2]], 4); o += len[px[2]];
      }
      *o++ = '\n';
      used = (size_t)(o - buf.data());
    }
    fwrite(buf.data(), 1, used, f);
  }
  fclose(f);
  return RT_OK;
}

#include "rt_multi.inl"
