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
//   3. BVH top down by binned surface-area-heuristic splits (the algorithm of host/bvh_builder.cpp), one launch per
//      tree level, every node split by its own thread block or warp with its bins in shared memory;
//   4. 64-byte pair nodes as bvh_builder.hpp defines them, in breadth-first order;
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
// 3. BVH: binned surface-area-heuristic splits, top down, one level per launch
//
// The algorithm of host/bvh_builder.cpp (16 centroid bins per axis, cost = area x count, median fall-back, depth
// guard), so the tree has the host tree's quality; what changes is who does the work: every node of a level is split
// by its own group of threads -- a 1024-thread block for nodes of more than kSahBigNode primitives, a single warp
// otherwise -- and all nodes of a level are split by one launch.  A group makes three passes over its node's
// primitives (centroid bounds; bins; stable partition into the other index buffer) with its bins in shared memory.
// Primitives arrive in Morton order, so the lanes of a warp mostly fall into one or two bins: lanes with the same
// bin are combined with match / redux before the shared-memory atomics.
// ---------------------------------------------------------------------------------------------
constexpr int kSahBins = 16;
constexpr int kSahMaxDepth = 60;    // device traversal stack holds 64 entries (bvh_builder.cpp kMaxDepth)
constexpr int kSahBigNode = 2048;

struct SahNodes {   // 2N entries; node 0 is the root, children are created in pairs (right = left + 1)
  int32_t *lo, *cnt, *left, *depth;
  float *box;       // [6] bounds of the node's (padded) primitive boxes, written by its parent
};

struct SahShared {
  unsigned int cb[6];
  unsigned int bin_box[3][kSahBins][6];
  int bin_cnt[3][kSahBins];
  unsigned int child_box[2][6];
  int axis, split, nleft, fallback, done_l, done_r;
  float cmin[3], scale[3];
};

__device__ __forceinline__ float sah_half_area(const float *mn, const float *mx) {
  const float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
  if (dx < 0 || dy < 0 || dz < 0) return 0.f;
  return dx * dy + dy * dz + dz * dx;
}
__device__ __forceinline__ int sah_bin(float c, float cmin, float scale) {
  int b = (int)((c - cmin) * scale);
  return b < 0 ? 0 : (b >= kSahBins ? kSahBins - 1 : b);
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_sah_split(const int32_t *__restrict__ sel, const int32_t *__restrict__ act,
                                                     const int child_base, const int leaf, const float *__restrict__ pbox,
                                                     const int32_t *__restrict__ in, int32_t *out, int32_t *final_order, SahNodes nd) {
  __shared__ SahShared s;
  const int tid = threadIdx.x, lane = tid & 31;
  const int k = sel[blockIdx.x], node = act[k];
  const int lo = nd.lo[node], cnt = nd.cnt[node], depth = nd.depth[node];
  const unsigned FULL = 0xffffffffu;
  // ---- init ----
  for (int i = tid; i < 3 * kSahBins * 6; i += BLOCK) {
    const int v = i % 6;
    (&s.bin_box[0][0][0])[i] = v < 3 ? 0xffffffffu : 0u;  // ordered encodings: min slots start at +max, max slots at -max
  }
  for (int i = tid; i < 3 * kSahBins; i += BLOCK) (&s.bin_cnt[0][0])[i] = 0;
  if (tid < 6) s.cb[tid] = tid < 3 ? 0xffffffffu : 0u;
  if (tid < 12) (&s.child_box[0][0])[tid] = (tid % 6) < 3 ? 0xffffffffu : 0u;
  if (tid == 0) { s.done_l = 0; s.done_r = 0; }
  __syncthreads();
  // ---- pass 1: centroid bounds ----
  {
    unsigned int mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
    for (int i = tid; i < cnt; i += BLOCK) {
      const float *b = pbox + (size_t)in[lo + i] * 6;
      for (int a = 0; a < 3; ++a) {
        const unsigned int c = f2ord(0.5f * (b[a] + b[3 + a]));
        mn[a] = min(mn[a], c); mx[a] = max(mx[a], c);
      }
    }
    for (int a = 0; a < 3; ++a) {
      const unsigned int wmn = __reduce_min_sync(FULL, mn[a]), wmx = __reduce_max_sync(FULL, mx[a]);
      if (lane == 0) { atomicMin(&s.cb[a], wmn); atomicMax(&s.cb[3 + a], wmx); }
    }
  }
  __syncthreads();
  if (tid < 3) {
    const float cmin = ord2f(s.cb[tid]), cext = ord2f(s.cb[3 + tid]) - cmin;
    s.cmin[tid] = cmin;
    s.scale[tid] = cext > 0.f ? (float)kSahBins / cext : 0.f;  // 0: the axis offers no split
  }
  int need = 0;
  while ((1 << need) < cnt) ++need;
  const bool use_sah = depth + need < kSahMaxDepth - 1;
  __syncthreads();
  // ---- pass 2: bins ----
  if (use_sah) {
    for (int base = 0; base < cnt; base += BLOCK) {
      const int i = base + tid;
      const bool valid = i < cnt;
      float bx[6] = {0, 0, 0, 0, 0, 0};
      if (valid) {
        const float *b = pbox + (size_t)in[lo + i] * 6;
        for (int v = 0; v < 6; ++v) bx[v] = b[v];
      }
      unsigned int ob[6];
      for (int v = 0; v < 6; ++v) ob[v] = f2ord(bx[v]);
      for (int a = 0; a < 3; ++a) {
        const float scale = s.scale[a];
        if (!(scale > 0.f)) continue;  // (uniform across the block)
        const int bin = valid ? sah_bin(0.5f * (bx[a] + bx[3 + a]), s.cmin[a], scale) : 64 + lane;
        const unsigned m = __match_any_sync(FULL, bin);
        unsigned int r[6];
        for (int v = 0; v < 3; ++v) { r[v] = __reduce_min_sync(m, ob[v]); r[3 + v] = __reduce_max_sync(m, ob[3 + v]); }
        if (valid && lane == __ffs(m) - 1) {
          atomicAdd(&s.bin_cnt[a][bin], __popc(m));
          for (int v = 0; v < 3; ++v) { atomicMin(&s.bin_box[a][bin][v], r[v]); atomicMax(&s.bin_box[a][bin][3 + v], r[3 + v]); }
        }
      }
    }
  }
  __syncthreads();
  // ---- split selection (bvh_builder.cpp sah_partition: axes in order, first strictly better candidate wins) ----
  if (tid == 0) {
    float best_cost = 3.402823466e+38f;
    int best_axis = -1, best_split = -1, best_left = 0;
    if (use_sah) {
      for (int a = 0; a < 3; ++a) {
        if (!(s.scale[a] > 0.f)) continue;
        float right_area[kSahBins];
        int right_cnt[kSahBins];
        float amn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}, amx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
        int c = 0;
        for (int b = kSahBins - 1; b > 0; --b) {
          if (s.bin_cnt[a][b] > 0)
            for (int v = 0; v < 3; ++v) { amn[v] = fminf(amn[v], ord2f(s.bin_box[a][b][v])); amx[v] = fmaxf(amx[v], ord2f(s.bin_box[a][b][3 + v])); }
          c += s.bin_cnt[a][b];
          right_area[b] = sah_half_area(amn, amx); right_cnt[b] = c;
        }
        float lmn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}, lmx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
        int lc = 0;
        for (int b = 0; b < kSahBins - 1; ++b) {
          if (s.bin_cnt[a][b] > 0)
            for (int v = 0; v < 3; ++v) { lmn[v] = fminf(lmn[v], ord2f(s.bin_box[a][b][v])); lmx[v] = fmaxf(lmx[v], ord2f(s.bin_box[a][b][3 + v])); }
          lc += s.bin_cnt[a][b];
          if (lc == 0 || right_cnt[b + 1] == 0) continue;
          const float cost = sah_half_area(lmn, lmx) * (float)lc + right_area[b + 1] * (float)right_cnt[b + 1];
          if (cost < best_cost) { best_cost = cost; best_axis = a; best_split = b; best_left = lc; }
        }
      }
    }
    s.fallback = best_axis < 0 ? 1 : 0;
    s.axis = best_axis < 0 ? 0 : best_axis;
    s.split = best_split;
    s.nleft = best_axis < 0 ? cnt / 2 : best_left;  // no usable split: halve the (Morton-ordered) range
  }
  __syncthreads();
  // ---- pass 3: stable partition into the other buffer, child boxes ----
  const int axis = s.axis, split = s.split, nleft = s.nleft;
  const bool fallback = s.fallback != 0;
  const float cmin = s.cmin[axis], scale = s.scale[axis];
  const bool leaf_l = nleft <= leaf, leaf_r = cnt - nleft <= leaf;
  unsigned int cb[2][6];
  for (int c = 0; c < 2; ++c)
    for (int v = 0; v < 6; ++v) cb[c][v] = v < 3 ? 0xffffffffu : 0u;
  for (int base = 0; base < cnt; base += BLOCK) {
    const int i = base + tid;
    const bool valid = i < cnt;
    int p = 0;
    bool left = false;
    if (valid) {
      p = in[lo + i];
      const float *b = pbox + (size_t)p * 6;
      left = fallback ? (i < nleft) : (sah_bin(0.5f * (b[axis] + b[3 + axis]), cmin, scale) <= split);
      const int c = left ? 0 : 1;
      for (int v = 0; v < 3; ++v) { cb[c][v] = min(cb[c][v], f2ord(b[v])); cb[c][3 + v] = max(cb[c][3 + v], f2ord(b[3 + v])); }
    }
    int rank_l, total_l;
    if constexpr (BLOCK == 32) {
      const unsigned m = __ballot_sync(FULL, left);
      rank_l = __popc(m & ((1u << lane) - 1u)); total_l = __popc(m);
    } else {
      typedef cub::BlockScan<int, BLOCK> Scan;
      __shared__ typename Scan::TempStorage tmp;
      Scan(tmp).ExclusiveSum(left ? 1 : 0, rank_l, total_l);
    }
    if (valid) {
      const int dst = left ? lo + s.done_l + rank_l : lo + nleft + s.done_r + (tid - rank_l);
      out[dst] = p;
      if (left ? leaf_l : leaf_r) final_order[dst] = p;
    }
    __syncthreads();
    if (tid == 0) { s.done_l += total_l; s.done_r += min(BLOCK, cnt - base) - total_l; }
    __syncthreads();
  }
  for (int c = 0; c < 2; ++c)
    for (int v = 0; v < 6; ++v) {
      const unsigned int w = v < 3 ? __reduce_min_sync(FULL, cb[c][v]) : __reduce_max_sync(FULL, cb[c][v]);
      if (lane == 0) { if (v < 3) atomicMin(&s.child_box[c][v], w); else atomicMax(&s.child_box[c][v], w); }
    }
  __syncthreads();
  if (tid < 2) {
    const int c = tid, id = child_base + 2 * k + c;
    nd.lo[id] = c == 0 ? lo : lo + nleft;
    nd.cnt[id] = c == 0 ? nleft : cnt - nleft;
    nd.depth[id] = depth + 1;
    nd.left[id] = -1;
    for (int v = 0; v < 6; ++v) nd.box[(size_t)id * 6 + v] = ord2f(s.child_box[c][v]);
    if (c == 0) nd.left[node] = id;
  }
}

// the children created by one level: which of them are split next (in id order: deterministic node numbering) ...
__global__ void k_sah_flags(const int first, const int n, const SahNodes nd, const int leaf, int32_t *flag_act) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag_act[i] = nd.cnt[first + i] > leaf ? 1 : 0;
}
__global__ void k_sah_lists(const int first, const int n, const int32_t *__restrict__ flag_act, const int32_t *__restrict__ pos_act,
                            int32_t *act, int32_t *totals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flag_act[i]) act[pos_act[i]] = first + i;
  if (i == n - 1) totals[0] = pos_act[i] + flag_act[i];
}
// ... and by which kind of group: rank k of the active list -> list of the big / of the small nodes
__global__ void k_sah_kind(const int32_t *__restrict__ act, const int n_act, const SahNodes nd, int32_t *flag_big, int32_t *flag_small) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_act) return;
  const bool big = nd.cnt[act[k]] > kSahBigNode;
  flag_big[k] = big ? 1 : 0;
  flag_small[k] = big ? 0 : 1;
}
__global__ void k_sah_select(const int n_act, const int32_t *__restrict__ flag_big, const int32_t *__restrict__ pos_big,
                             const int32_t *__restrict__ pos_small, int32_t *sel_big, int32_t *sel_small, int32_t *totals) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_act) return;
  if (flag_big[k]) sel_big[pos_big[k]] = k; else sel_small[pos_small[k]] = k;
  if (k == n_act - 1) { totals[1] = pos_big[k] + flag_big[k]; totals[2] = pos_small[k] + (flag_big[k] ? 0 : 1); }
}
__global__ void k_sah_root(SahNodes nd, const int N, int32_t *act) {
  nd.lo[0] = 0; nd.cnt[0] = N; nd.depth[0] = 0; nd.left[0] = -1;
  act[0] = 0;
}

// ---------------------------------------------------------------------------------------------
// 4. pair nodes
// ---------------------------------------------------------------------------------------------
__global__ void k_pair_flags(const int n_nodes, const SahNodes nd, const int leaf, int32_t *flag) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id < n_nodes) flag[id] = nd.cnt[id] > leaf ? 1 : 0;
}
__global__ void k_emit_pairs(const int n_nodes, const SahNodes nd, const int leaf, const int T, const int32_t *__restrict__ flag,
                             const int32_t *__restrict__ pos, const int32_t *__restrict__ prim_order, float4 *nodes_out,
                             unsigned int *n_leaves, int32_t *max_depth, double *sah) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  // surface-area-heuristic cost with the host builder's definition (bvh_builder.cpp): inner boxes + leaf boxes x count
  double cost = 0.0;
  unsigned leaves = 0;
  int deepest = 0;
  if (id < n_nodes && flag[id]) {
    const int ch[2] = {nd.left[id], nd.left[id] + 1};
    int code[2];
    const float *b0 = nd.box + (size_t)ch[0] * 6, *b1 = nd.box + (size_t)ch[1] * 6;
    {
      float mn[3], mx[3];
      for (int v = 0; v < 3; ++v) { mn[v] = fminf(b0[v], b1[v]); mx[v] = fmaxf(b0[3 + v], b1[3 + v]); }
      cost = (double)sah_half_area(mn, mx);
    }
    for (int c = 0; c < 2; ++c) {
      const int cid = ch[c];
      if (nd.cnt[cid] > leaf) {
        code[c] = pos[cid];
      } else {
        const int first = nd.lo[cid], cnt = nd.cnt[cid];
        bool mixed = false;
        for (int q = first; q < first + cnt; ++q) mixed |= prim_order[q] >= T;
        code[c] = ~((first << 5) | ((mixed ? 1 : 0) << 4) | (cnt - 1));
        leaves++;
        deepest = max(deepest, nd.depth[cid]);
        const float *b = nd.box + (size_t)cid * 6;
        cost += (double)sah_half_area(b, b + 3) * cnt;
      }
    }
    float4 *o = nodes_out + (size_t)pos[id] * 4;
    o[0] = make_float4(b0[0], b0[3], b0[1], b0[4]);
    o[1] = make_float4(b1[0], b1[3], b1[1], b1[4]);
    o[2] = make_float4(b0[2], b0[5], b1[2], b1[5]);
    o[3] = make_float4(__int_as_float(code[0]), __int_as_float(code[1]), 0.f, 0.f);
  }
  for (int off = 16; off > 0; off >>= 1) {
    cost += __shfl_down_sync(0xffffffffu, cost, off);
    leaves += __shfl_down_sync(0xffffffffu, leaves, off);
    deepest = max(deepest, __shfl_down_sync(0xffffffffu, deepest, off));
  }
  if ((threadIdx.x & 31) == 0) {
    if (leaves) { atomicAdd(n_leaves, leaves); atomicMax(max_depth, deepest); }
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
