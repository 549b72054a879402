// rt_build.cuh -- the acceleration structures built ON the GPU (sm_100a).
//
// Replaces, for large scenes, the host build behind rt_scene_create:
//   BoxTree::BoxTree / split / clasifyFace   src/boxTree.cpp:11-31, 88-147, 203-336   (reference octree, kept as the
//                                                                                      candidate filter, bit-identical)
//   + the BVH the device traverses and the per-triangle constants of rayTriangleIntersection (src/flyscene.cpp:787-819).
// (citations relative to /root/reference).  Compiled with -fmad=false like the rest of the library: the octree
// classification and the baked constants must be the floats the host code -- and the reference -- compute.
//
//   1. reference octree, level by level: one thread per (face of a node being split, octant) evaluates the
//      reference's membership test (vertex in box, else its mis-normalised separating-axis test) -> per-octant counts,
//      (face, octant) pairs; octants are then classified (empty / leaf / split / "exactly capacity": hidden) exactly as
//      BoxTree::split does, leaves append their (face, leaf) references, split octants form the next level;
//   2. primitive boxes (sliver faces widened from their octree leaves, see rt_api.cu), 63-bit Morton codes, radix sort;
//   3. BVH by PLOC (parallel locally-ordered clustering, Meister & Bittner 2017): every cluster looks for the
//      neighbour within +-kPlocRadius positions of the Morton order whose union with it has the smallest area;
//      mutual nearest neighbours merge; repeat until one cluster is left.  Unlike a plain Morton-split LBVH this
//      bottom-up agglomeration gives trees of SAH quality;
//   4. top-down pass in reverse creation order: first primitive slot and depth of every node (leaf order = soup
//      order); subtrees of at most `leaf` primitives become leaves; 64-byte pair nodes as bvh_builder.hpp defines them;
//   5. bake: primitive soup (80 B), shading table (112 B) with the host's expressions.
#pragma once

#include <cub/cub.cuh>

#include "rt_device.cuh"

namespace rtb {

using rtd::PRIM_ILLUM9;
using rtd::PRIM_SPHERE;

// ---- float3 algebra with the evaluation order of host/vec3.hpp (which is Eigen 3.3.7's) ----
struct B3 { float x, y, z; };
__device__ __forceinline__ B3 b3(float x, float y, float z) { B3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ B3 b3(const float *p) { return b3(p[0], p[1], p[2]); }
__device__ __forceinline__ B3 operator+(B3 a, B3 b) { return b3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ B3 operator-(B3 a, B3 b) { return b3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ B3 operator*(float s, B3 a) { return b3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ B3 operator/(B3 a, float s) { return b3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ float bdot(B3 a, B3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
__device__ __forceinline__ B3 bcross(B3 a, B3 b) { return b3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ B3 bnormalized(B3 a) {
  const float z = bdot(a, a);
  if (z > 0.f) return a / sqrtf(z);
  return a;
}
__device__ __forceinline__ float bget(B3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// ---------------------------------------------------------------------------------------------
// 1. reference octree (host/ref_octree.cpp, statement for statement)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool oct_inside(B3 mn, B3 mx, B3 v) {
  return mn.x <= v.x && mx.x >= v.x && mn.y <= v.y && mx.y >= v.y && mn.z <= v.z && mx.z >= v.z;
}
__device__ __forceinline__ bool oct_axis_ok(float p0, float p1, float rad) {
  const float lo = fminf(p1, p0) , hi = fmaxf(p1, p0);  // std::min / std::max of finite values
  return !(lo > rad || hi < -rad);
}
__device__ __forceinline__ bool oct_probe_x(float a, float b, float fa, float fb, B3 u, B3 v, B3 h) {
  return oct_axis_ok(a * u.y - b * u.z, a * v.y - b * v.z, fa * h.y + fb * h.z);
}
__device__ __forceinline__ bool oct_probe_y(float a, float b, float fa, float fb, B3 u, B3 v, B3 h) {
  return oct_axis_ok(-a * u.x + b * u.z, -a * v.x + b * v.z, fa * h.x + fb * h.z);
}
__device__ __forceinline__ bool oct_probe_z(float a, float b, float fa, float fb, B3 u, B3 v, B3 h) {
  return oct_axis_ok(a * u.x - b * u.y, a * v.x - b * v.y, fa * h.x + fb * h.y);
}
__device__ __forceinline__ bool oct_plane_overlaps(B3 n, B3 p, B3 h) {
  float lo[3], hi[3];
  for (int i = 0; i < 3; ++i) {
    const float ni = bget(n, i), hi_ = bget(h, i), pi = bget(p, i);
    if (ni > 0.0f) { lo[i] = -hi_ - pi; hi[i] = hi_ - pi; }
    else { lo[i] = hi_ - pi; hi[i] = -hi_ - pi; }
  }
  if (bdot(n, b3(lo)) > 0.0f) return false;
  return bdot(n, b3(hi)) >= 0.0f;
}
// Does the reference put the face into the octant [mn, mx]?  (any vertex inside, else its SAT variant)
__device__ __forceinline__ bool oct_belongs(const float *verts, int face, B3 mn, B3 mx) {
  const float *p = verts + (size_t)face * 9;
  const B3 v0 = b3(p), v1 = b3(p + 3), v2 = b3(p + 6);
  if (oct_inside(mn, mx, v0) || oct_inside(mn, mx, v1) || oct_inside(mn, mx, v2)) return true;
  const B3 c = b3(mn.x + (mx.x - mn.x) / 2.f, mn.y + (mx.y - mn.y) / 2.f, mn.z + (mx.z - mn.z) / 2.f);
  const B3 h = bnormalized(mx - c);
  const B3 a = bnormalized(v0 - c), b = bnormalized(v1 - c), cc = bnormalized(v2 - c);
  const B3 e0 = b - a, e1 = cc - b, e2 = a - cc;
  float fx = fabsf(e0.x), fy = fabsf(e0.y), fz = fabsf(e0.z);
  if (!oct_probe_x(e0.z, e0.y, fz, fy, a, cc, h)) return false;
  if (!oct_probe_y(e0.z, e0.x, fz, fx, a, cc, h)) return false;
  if (!oct_probe_z(e0.y, e0.x, fy, fx, b, cc, h)) return false;
  fx = fabsf(e1.x); fy = fabsf(e1.y); fz = fabsf(e1.z);
  if (!oct_probe_x(e1.z, e1.y, fz, fy, a, cc, h)) return false;
  if (!oct_probe_y(e1.z, e1.x, fz, fx, a, cc, h)) return false;
  if (!oct_probe_z(e1.y, e1.x, fy, fx, a, b, h)) return false;
  fx = fabsf(e2.x); fy = fabsf(e2.y); fz = fabsf(e2.z);
  if (!oct_probe_x(e2.z, e2.y, fz, fy, a, b, h)) return false;
  if (!oct_probe_y(e2.z, e2.x, fz, fx, a, b, h)) return false;
  if (!oct_probe_z(e2.y, e2.x, fy, fx, b, cc, h)) return false;
  for (int k = 0; k < 3; ++k) {
    const float lo = fminf(fminf(bget(a, k), bget(b, k)), bget(cc, k)), hi = fmaxf(fmaxf(bget(a, k), bget(b, k)), bget(cc, k));
    if (lo > bget(h, k) || hi < -bget(h, k)) return false;
  }
  return oct_plane_overlaps(bnormalized(bcross(a - b, a - cc)), a, h);
}
// octant k of [lo, hi] with the reference's own expression trees (left-to-right sums, src/boxTree.cpp:103-120)
__device__ __forceinline__ void oct_child_box(B3 lo, B3 hi, int k, B3 &cmn, B3 &cmx) {
  const float dx = (hi.x - lo.x) / 2, dy = (hi.y - lo.y) / 2, dz = (hi.z - lo.z) / 2;
  const B3 vx = b3(dx, 0, 0), vy = b3(0, dy, 0), vz = b3(0, 0, dz);
  switch (k) {
    case 0: cmn = lo; cmx = lo + vx + vy + vz; break;
    case 1: cmn = lo + vz; cmx = lo + vx + vy + 2 * vz; break;
    case 2: cmn = lo + vy; cmx = lo + vx + 2 * vy + vz; break;
    case 3: cmn = lo + vy + vz; cmx = lo + vx + 2 * vy + 2 * vz; break;
    case 4: cmn = lo + vx; cmx = lo + 2 * vx + vy + vz; break;
    case 5: cmn = lo + vx + vz; cmx = hi - vy; break;
    case 6: cmn = lo + vx + vy; cmx = hi - vz; break;
    default: cmn = lo + vx + vy + vz; cmx = hi; break;
  }
}

struct OctLevelNode { float mn[3], mx[3]; int32_t node_id; int32_t pad; };
enum OctState : int32_t { OCT_EMPTY = 0, OCT_LEAF = 1, OCT_SPLIT = 2, OCT_HIDDEN = 3 };
// counters of one octree level (zeroed before the level) ...
struct OctLevelCtr {
  unsigned int n_out;          // (face, octant) memberships found by k_oct_classify
  unsigned int n_new_nodes;    // non-empty octants = octree nodes created by this level
  unsigned int n_split;        // octants that are split again = nodes of the next level
  unsigned int n_leaf_refs;    // memberships that end in a leaf
  unsigned int n_next_pairs;   // memberships handed to the next level
  unsigned int refs_written, pairs_written;  // k_oct_route cursors
  unsigned int overflow;
};
// ... and of the whole octree (BoxTree shape: what rt_ref_octree_stats reports for the host build)
struct OctTotals { unsigned int n_leaves, n_inner, max_leaf, pad; unsigned long long n_refs; };

// (pair, octant) -> membership; counts per octant, compacted (face, octant slot) pairs
__global__ void k_oct_classify(const float *__restrict__ verts, const int2 *__restrict__ pairs, const unsigned n_pairs,
                               const OctLevelNode *__restrict__ level, unsigned int *child_cnt, int2 *out, const unsigned out_cap,
                               OctLevelCtr *ctr) {
  const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned pi = g >> 3, k = g & 7u;
  bool in = false;
  int face = 0, child = 0;
  if (pi < n_pairs) {
    const int2 pr = pairs[pi];
    face = pr.x;
    const OctLevelNode nd = level[pr.y];
    B3 cmn, cmx;
    oct_child_box(b3(nd.mn), b3(nd.mx), (int)k, cmn, cmx);
    in = oct_belongs(verts, face, cmn, cmx);
    child = pr.y * 8 + (int)k;
  }
  if (in) atomicAdd(&child_cnt[child], 1u);
  const int slot = rtd::warp_append(reinterpret_cast<int *>(&ctr->n_out), in);
  if (in) {
    if ((unsigned)slot < out_cap) out[slot] = make_int2(face, child);
    else atomicExch(&ctr->overflow, 1u);
  }
}

// BoxTree::split's classification of every octant of this level (src/boxTree.cpp:122-146): flags for the two prefix
// sums that number the new nodes and the nodes of the next level (deterministic ids, no atomics)
__global__ void k_oct_flags(const unsigned n_children, const unsigned int *__restrict__ child_cnt, const int capacity, const int depth,
                            int32_t *flag_node, int32_t *flag_split) {
  const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_children) return;
  const unsigned cnt = child_cnt[c];
  flag_node[c] = cnt > 0u ? 1 : 0;
  flag_split[c] = ((int)cnt > capacity && depth > 0) ? 1 : 0;
}

__global__ void k_oct_decide(const OctLevelNode *__restrict__ level, const unsigned n_children, const unsigned int *__restrict__ child_cnt,
                             const int capacity, const int depth, const int32_t *__restrict__ flag_node,
                             const int32_t *__restrict__ pos_node, const int32_t *__restrict__ flag_split,
                             const int32_t *__restrict__ pos_split, const int node_base, int32_t *child_state, float4 *level_box,
                             OctLevelNode *next_level, OctLevelCtr *ctr, OctTotals *tot) {
  const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_children) return;
  if (c == n_children - 1u) {
    ctr->n_new_nodes = (unsigned)(pos_node[c] + flag_node[c]);
    ctr->n_split = (unsigned)(pos_split[c] + flag_split[c]);
  }
  const unsigned cnt = child_cnt[c];
  if (cnt == 0u) { child_state[c] = OCT_EMPTY; return; }
  const OctLevelNode nd = level[c >> 3];
  B3 cmn, cmx;
  oct_child_box(b3(nd.mn), b3(nd.mx), (int)(c & 7u), cmn, cmx);
  const int local = pos_node[c], id = node_base + local;
  level_box[2 * (size_t)local] = make_float4(cmn.x, cmn.y, cmn.z, __int_as_float(nd.node_id));
  level_box[2 * (size_t)local + 1] = make_float4(cmx.x, cmx.y, cmx.z, 0.f);
  if ((int)cnt < capacity || depth <= 0) {
    atomicAdd(&tot->n_leaves, 1u);
    atomicAdd(&tot->n_refs, (unsigned long long)cnt);
    atomicMax(&tot->max_leaf, cnt);
    atomicAdd(&ctr->n_leaf_refs, cnt);
    child_state[c] = (id << 2) | OCT_LEAF;
    return;
  }
  atomicAdd(&tot->n_inner, 1u);
  if (flag_split[c]) {
    const int slot = pos_split[c];
    OctLevelNode nx;
    nx.mn[0] = cmn.x; nx.mn[1] = cmn.y; nx.mn[2] = cmn.z; nx.mx[0] = cmx.x; nx.mx[1] = cmx.y; nx.mx[2] = cmx.z;
    nx.node_id = id; nx.pad = 0;
    next_level[slot] = nx;
    atomicAdd(&ctr->n_next_pairs, cnt);
    child_state[c] = (slot << 2) | OCT_SPLIT;
  } else {
    child_state[c] = OCT_HIDDEN;  // exactly `capacity` faces: neither a leaf nor split -- its faces are never offered
  }
}

// (face, octant) pairs -> leaf references or pairs of the next level
__global__ void k_oct_route(const int2 *__restrict__ in, const unsigned n_in, const int32_t *__restrict__ child_state, int2 *refs,
                            const unsigned refs_cap, int2 *next_pairs, const unsigned next_cap, OctLevelCtr *ctr) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  int face = 0, st = OCT_EMPTY, val = 0;
  if (i < n_in) {
    const int2 pr = in[i];
    face = pr.x;
    const int s = child_state[pr.y];
    st = s & 3; val = s >> 2;
  }
  const int rs = rtd::warp_append(reinterpret_cast<int *>(&ctr->refs_written), st == OCT_LEAF);
  if (st == OCT_LEAF) {
    if ((unsigned)rs < refs_cap) refs[rs] = make_int2(face, val);
    else atomicExch(&ctr->overflow, 1u);
  }
  const int ns = rtd::warp_append(reinterpret_cast<int *>(&ctr->pairs_written), st == OCT_SPLIT);
  if (st == OCT_SPLIT) {
    if ((unsigned)ns < next_cap) next_pairs[ns] = make_int2(face, val);
    else atomicExch(&ctr->overflow, 1u);
  }
}

__global__ void k_iota_pairs(int2 *pairs, const unsigned n) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pairs[i] = make_int2((int)i, 0);
}
// (face, leaf) references of one level -> 64-bit sort keys (face major): the sorted keys ARE the CSR rows
__global__ void k_pack_ref_keys(const int2 *__restrict__ refs, const unsigned n, unsigned long long *keys) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = ((unsigned long long)(unsigned)refs[i].x << 32) | (unsigned)refs[i].y;
}
// sorted keys -> face_off[T+1], face_leaf[R]
__global__ void k_face_csr(const unsigned long long *__restrict__ keys, const unsigned R, const int T, int32_t *face_off, int32_t *face_leaf) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  const unsigned long long k = keys[i];
  const int f = (int)(k >> 32);
  face_leaf[i] = (int)(unsigned)(k & 0xffffffffull);
  const int fprev = i > 0 ? (int)(keys[i - 1] >> 32) : -1;
  for (int g = fprev + 1; g <= f; ++g) face_off[g] = (int)i;
  if (i == R - 1) for (int g = f + 1; g <= T; ++g) face_off[g] = (int)R;
}

// ---------------------------------------------------------------------------------------------
// 2. primitive boxes, Morton codes
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float hdot3(const float *a, const float *b) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); }

// order-preserving float <-> unsigned for atomicMin / atomicMax
__device__ __forceinline__ unsigned f2ord(float f) { const unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// min / max over all vertex coordinates: ord[0..2] = min xyz, ord[3..5] = max xyz as order-preserving unsigneds
// (the host initialises them and derives both the reference root box -- BoundingBox(Mesh&), src/boundingBox.cpp:14-43,
// whose max starts at FLT_MIN -- and the scene bounds from them)
__global__ void k_vertex_bounds(const float *__restrict__ verts, const int T, unsigned int *ord) {
  float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}, mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T; i += gridDim.x * blockDim.x) {
    const float *v = verts + (size_t)i * 9;
    for (int a = 0; a < 3; ++a) {
      mn[a] = fminf(mn[a], fminf(v[a], fminf(v[3 + a], v[6 + a])));
      mx[a] = fmaxf(mx[a], fmaxf(v[a], fmaxf(v[3 + a], v[6 + a])));
    }
  }
  for (int a = 0; a < 3; ++a) {
    for (int off = 16; off > 0; off >>= 1) {
      mn[a] = fminf(mn[a], __shfl_down_sync(0xffffffffu, mn[a], off));
      mx[a] = fmaxf(mx[a], __shfl_down_sync(0xffffffffu, mx[a], off));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&ord[a], f2ord(mn[a]));
      atomicMax(&ord[3 + a], f2ord(mx[a]));
    }
  }
}

// boxes[i] = (min.xyz, max.xyz) of primitive i (faces 0..T-1, then spheres), padded by `pad`; sliver faces as in
// rt_api.cu bake_scene: sin^2 <= 1e-7 -> union with the face's octree leaf boxes, sin^2 <= 1e-2 -> extra padding.
__global__ void k_prim_boxes(const float *__restrict__ verts, const int T, const float *__restrict__ spheres, const int S,
                             const float4 *__restrict__ oct_box, const int32_t *__restrict__ face_off,
                             const int32_t *__restrict__ face_leaf, const float pad, float *boxes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T + S) return;
  float mn[3], mx[3];
  if (i < T) {
    const float *v = verts + (size_t)i * 9;
    for (int a = 0; a < 3; ++a) {
      mn[a] = fminf(v[a], fminf(v[3 + a], v[6 + a]));
      mx[a] = fmaxf(v[a], fmaxf(v[3 + a], v[6 + a]));
    }
    if (oct_box != nullptr) {
      const float e0[3] = {v[6] - v[0], v[7] - v[1], v[8] - v[2]}, e1[3] = {v[3] - v[0], v[4] - v[1], v[5] - v[2]};
      const float d00 = hdot3(e0, e0), d01 = hdot3(e0, e1), d11 = hdot3(e1, e1);
      const float det = d00 * d11 - d01 * d01;
      const float sin2 = det / (d00 * d11);
      if (!(sin2 > 1e-2f)) {
        if (!(sin2 > 1e-7f) || !isfinite(1.f / det)) {
          for (int k = face_off[i]; k < face_off[i + 1]; ++k) {
            const float4 lo = oct_box[2 * (size_t)face_leaf[k]], hi = oct_box[2 * (size_t)face_leaf[k] + 1];
            mn[0] = fminf(mn[0], lo.x); mn[1] = fminf(mn[1], lo.y); mn[2] = fminf(mn[2], lo.z);
            mx[0] = fmaxf(mx[0], hi.x); mx[1] = fmaxf(mx[1], hi.y); mx[2] = fmaxf(mx[2], hi.z);
          }
        } else {
          const float extra = 1e-6f * sqrtf(fmaxf(d00, d11)) / sin2;
          for (int a = 0; a < 3; ++a) { mn[a] -= extra; mx[a] += extra; }
        }
      }
    }
  } else {
    const float *s = spheres + (size_t)(i - T) * 4;
    for (int a = 0; a < 3; ++a) { mn[a] = s[a] - s[3]; mx[a] = s[a] + s[3]; }
  }
  for (int a = 0; a < 3; ++a) {
    boxes[(size_t)i * 6 + a] = mn[a] - pad;
    boxes[(size_t)i * 6 + 3 + a] = mx[a] + pad;
  }
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long x) {
  x &= 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
struct Bounds6 { float lo[3], hi[3]; };
__global__ void k_morton(const float *__restrict__ boxes, const int N, const Bounds6 sb, unsigned long long *keys, int32_t *ids) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  unsigned long long code = 0;
  for (int a = 0; a < 3; ++a) {
    const float lo = sb.lo[a], hi = sb.hi[a];
    const float c = 0.5f * (boxes[(size_t)i * 6 + a] + boxes[(size_t)i * 6 + 3 + a]);
    const float ext = hi - lo;
    float u = ext > 0.f ? (c - lo) / ext : 0.f;
    u = fminf(fmaxf(u, 0.f), 1.f);
    const unsigned long long q = (unsigned long long)fminf(u * 2097152.f, 2097151.f);
    code |= spread21(q) << a;
  }
  keys[i] = code;
  ids[i] = i;
}

// ---------------------------------------------------------------------------------------------
// 3. PLOC
// ---------------------------------------------------------------------------------------------
constexpr int kPlocRadius = 16;

struct BuildNodes {      // binary tree under construction: nodes 0..N-1 are the leaves (sorted primitives)
  float *box;            // [2N][6]
  int32_t *left, *right; // children (-1 for leaves)
  int32_t *count;        // primitives below
  int32_t *first;        // first soup slot (top-down pass)
  int32_t *depth;
};

__global__ void k_ploc_init(const float *__restrict__ prim_boxes, const int32_t *__restrict__ sorted_ids, const int N, BuildNodes bn,
                            int32_t *clusters) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int p = sorted_ids[i];
  for (int a = 0; a < 6; ++a) bn.box[(size_t)i * 6 + a] = prim_boxes[(size_t)p * 6 + a];
  bn.left[i] = -1; bn.right[i] = -1; bn.count[i] = 1;
  clusters[i] = i;
}

__device__ __forceinline__ float union_half_area(const float *a, const float *b) {
  const float dx = fmaxf(a[3], b[3]) - fminf(a[0], b[0]), dy = fmaxf(a[4], b[4]) - fminf(a[1], b[1]),
              dz = fmaxf(a[5], b[5]) - fminf(a[2], b[2]);
  return dx * dy + dy * dz + dz * dx;
}

// nearest neighbour of every cluster within the search radius (ties: the lower position)
__global__ void k_ploc_nn(const int32_t *__restrict__ clusters, const int n, const float *__restrict__ box, int32_t *nn) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float bi[6];
  const float *pb = box + (size_t)clusters[i] * 6;
  for (int a = 0; a < 6; ++a) bi[a] = pb[a];
  const int lo = max(0, i - kPlocRadius), hi = min(n - 1, i + kPlocRadius);
  float best = 3.4e38f;
  int bj = -1;
  for (int j = lo; j <= hi; ++j) {
    if (j == i) continue;
    const float A = union_half_area(bi, box + (size_t)clusters[j] * 6);
    if (A < best) { best = A; bj = j; }
  }
  nn[i] = bj;
}

// flags: merge[i] = 1 if position i creates a node (mutual pair, lower position), keep[i] = 0 if it vanishes
__global__ void k_ploc_flags(const int32_t *__restrict__ nn, const int n, int32_t *merge, int32_t *keep, int *n_merged) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool m = false;
  if (i < n) {
    const int j = nn[i];
    const bool mutual = j >= 0 && nn[j] == i;
    m = mutual && i < j;
    merge[i] = m ? 1 : 0;
    keep[i] = (mutual && i > j) ? 0 : 1;
  }
  const unsigned b = __ballot_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(n_merged, __popc(b));
}

__global__ void k_ploc_merge(const int32_t *__restrict__ clusters, const int32_t *__restrict__ nn, const int n,
                             const int32_t *__restrict__ merge, const int32_t *__restrict__ merge_pos, const int32_t *__restrict__ keep,
                             const int32_t *__restrict__ keep_pos, const int next_node, BuildNodes bn, int32_t *clusters_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!keep[i]) return;
  int c = clusters[i];
  if (merge[i]) {
    const int l = c, r = clusters[nn[i]];
    const int id = next_node + merge_pos[i];
    const float *bl = bn.box + (size_t)l * 6, *br = bn.box + (size_t)r * 6;
    float *bo = bn.box + (size_t)id * 6;
    for (int a = 0; a < 3; ++a) { bo[a] = fminf(bl[a], br[a]); bo[3 + a] = fmaxf(bl[3 + a], br[3 + a]); }
    bn.left[id] = l; bn.right[id] = r;
    bn.count[id] = bn.count[l] + bn.count[r];
    c = id;
  }
  clusters_out[keep_pos[i]] = c;
}

// ---------------------------------------------------------------------------------------------
// 4. top-down pass, pair nodes
// ---------------------------------------------------------------------------------------------
__global__ void k_topdown(const int id_begin, const int id_end, BuildNodes bn, int32_t *max_depth) {
  const int id = id_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= id_end) return;
  const int l = bn.left[id], r = bn.right[id];
  const int f = bn.first[id], d = bn.depth[id];
  bn.first[l] = f; bn.first[r] = f + bn.count[l];
  bn.depth[l] = d + 1; bn.depth[r] = d + 1;
  atomicMax(max_depth, d + 1);
}
__global__ void k_leaf_order(const int N, const BuildNodes bn, const int32_t *__restrict__ sorted_ids, int32_t *prim_order) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) prim_order[bn.first[i]] = sorted_ids[i];
}
// inner[k] = 1 for tree nodes that become pair nodes (more than `leaf` primitives), in REVERSE id order (root first)
__global__ void k_pair_flags(const int N, const int n_nodes, const BuildNodes bn, const int leaf, int32_t *flag) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_nodes - N) return;
  const int id = n_nodes - 1 - k;
  flag[k] = bn.count[id] > leaf ? 1 : 0;
}
__global__ void k_emit_pairs(const int N, const int n_nodes, const BuildNodes bn, const int leaf, const int T,
                             const int32_t *__restrict__ flag, const int32_t *__restrict__ pos, const int32_t *__restrict__ prim_order,
                             float4 *nodes_out, unsigned int *n_leaves, double *sah) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  // surface-area-heuristic cost with the host builder's definition (bvh_builder.cpp): inner boxes + leaf boxes x count
  double cost = 0.0;
  unsigned leaves = 0;
  if (k < n_nodes - N && flag[k]) {
    const int id = n_nodes - 1 - k;
    const int ch[2] = {bn.left[id], bn.right[id]};
    int code[2];
    cost = (double)union_half_area(bn.box + (size_t)id * 6, bn.box + (size_t)id * 6);
    for (int c = 0; c < 2; ++c) {
      const int cid = ch[c];
      if (bn.count[cid] > leaf) {
        code[c] = pos[n_nodes - 1 - cid];
      } else {
        const int first = bn.first[cid], cnt = bn.count[cid];
        bool mixed = false;
        for (int s = first; s < first + cnt; ++s) mixed |= prim_order[s] >= T;
        code[c] = ~((first << 5) | ((mixed ? 1 : 0) << 4) | (cnt - 1));
        leaves++;
        cost += (double)union_half_area(bn.box + (size_t)cid * 6, bn.box + (size_t)cid * 6) * cnt;
      }
    }
    const float *b0 = bn.box + (size_t)ch[0] * 6, *b1 = bn.box + (size_t)ch[1] * 6;
    float4 *o = nodes_out + (size_t)pos[k] * 4;
    o[0] = make_float4(b0[0], b0[3], b0[1], b0[4]);
    o[1] = make_float4(b1[0], b1[3], b1[1], b1[4]);
    o[2] = make_float4(b0[2], b0[5], b1[2], b1[5]);
    o[3] = make_float4(__int_as_float(code[0]), __int_as_float(code[1]), 0.f, 0.f);
  }
  for (int off = 16; off > 0; off >>= 1) {
    cost += __shfl_down_sync(0xffffffffu, cost, off);
    leaves += __shfl_down_sync(0xffffffffu, leaves, off);
  }
  if ((threadIdx.x & 31) == 0) {
    if (leaves) atomicAdd(n_leaves, leaves);
    if (cost != 0.0) atomicAdd(sah, cost);
  }
}

// ---------------------------------------------------------------------------------------------
// 5. bake (the expressions of rt_api.cu bake_scene / src/flyscene.cpp:792-810)
// ---------------------------------------------------------------------------------------------
__global__ void k_bake_prims(const int N, const int T, const int32_t *__restrict__ prim_order, const float *__restrict__ verts,
                             const float *__restrict__ fnormals, const int32_t *__restrict__ mat_id, const int32_t *__restrict__ mat_illum,
                             const float *__restrict__ spheres, const int32_t *__restrict__ sphere_mat, float4 *prims) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= N) return;
  const int p = prim_order[slot];
  float4 *q = prims + (size_t)slot * 5;
  if (p < T) {
    const float *v = verts + (size_t)p * 9, *n = fnormals + (size_t)p * 3;
    const float *a = v, *b = v + 3, *c = v + 6;
    const float e0[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    const float e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    const float d00 = hdot3(e0, e0), d01 = hdot3(e0, e1), d11 = hdot3(e1, e1);
    const float inv = 1 / (d00 * d11 - d01 * d01);
    const int illum = mat_illum[mat_id[p]];
    q[0] = make_float4(n[0], n[1], n[2], hdot3(n, a));
    q[1] = make_float4(a[0], a[1], a[2], __int_as_float(p));
    q[2] = make_float4(e0[0], e0[1], e0[2], d00);
    q[3] = make_float4(e1[0], e1[1], e1[2], d11);
    q[4] = make_float4(d01, inv, __int_as_float(illum == 9 ? (int)PRIM_ILLUM9 : 0), 0.f);
  } else {
    const int si = p - T;
    const float *s = spheres + (size_t)si * 4;
    const int illum = mat_illum[sphere_mat[si]];
    q[0] = make_float4(s[0], s[1], s[2], s[3]);
    q[1] = make_float4(0.f, 0.f, 0.f, __int_as_float(p));
    q[2] = make_float4(0.f, 0.f, 0.f, 0.f);
    q[3] = make_float4(0.f, 0.f, 0.f, 0.f);
    q[4] = make_float4(0.f, 0.f, __int_as_float((int)(PRIM_SPHERE | (illum == 9 ? PRIM_ILLUM9 : 0))), 0.f);
  }
}
__global__ void k_bake_shade(const int T, const float *__restrict__ verts, const float *__restrict__ fnormals,
                             const float *__restrict__ vnormals, const int32_t *__restrict__ mat_id, float4 *shade) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  const float *v = verts + (size_t)i * 9, *vn = vnormals + (size_t)i * 9, *n = fnormals + (size_t)i * 3;
  float4 *q = shade + (size_t)i * 7;
  q[0] = make_float4(v[0], v[1], v[2], __int_as_float(mat_id[i]));
  q[1] = make_float4(v[3], v[4], v[5], 0.f);
  q[2] = make_float4(v[6], v[7], v[8], 0.f);
  q[3] = make_float4(vn[0], vn[1], vn[2], 0.f);
  q[4] = make_float4(vn[3], vn[4], vn[5], 0.f);
  q[5] = make_float4(vn[6], vn[7], vn[8], 0.f);
  q[6] = make_float4(n[0], n[1], n[2], 0.f);
}
// union of the root pair's two child boxes (DevScene.bvh_min / bvh_max)
__global__ void k_root_bounds(const float4 *__restrict__ nodes, float *out6) {
  const float4 q0 = nodes[0], q1 = nodes[1], q2 = nodes[2];
  out6[0] = fminf(q0.x, q1.x); out6[1] = fminf(q0.z, q1.z); out6[2] = fminf(q2.x, q2.z);
  out6[3] = fmaxf(q0.y, q1.y); out6[4] = fmaxf(q0.w, q1.w); out6[5] = fmaxf(q2.y, q2.w);
}

}  // namespace rtb
