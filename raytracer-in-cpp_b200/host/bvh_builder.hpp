// bvh_builder.hpp -- host build of the flattened BVH that replaces the reference's octree.
//
// Replaces BoxTree::BoxTree / split / clasifyFace (src/boxTree.cpp:11-31, 88-147, 203-336 of the
// reference) and BoundingBox::BoundingBox(Mesh&) (src/boundingBox.cpp:14-43).  The reference
// octree is an index of *candidate* faces; any structure that never hides a face the reference's
// ray-triangle test would accept gives the same nearest hit, so the device path uses a binary
// SAH BVH instead: binned surface-area-heuristic build, <= 4 primitives per leaf, emitted as
// 64-byte "pair nodes" (both children's AABBs + child codes in one node, four float4) in depth-
// first order, with the primitives re-ordered into leaf order.
#pragma once

#include <cstdint>
#include <vector>

namespace rt {

struct Aabb {
  float mn[3], mx[3];
};

// One 64-byte node: the AABBs of its two children and their codes.
//   q0 = (c0.min.x, c0.max.x, c0.min.y, c0.max.y)
//   q1 = (c1.min.x, c1.max.x, c1.min.y, c1.max.y)
//   q2 = (c0.min.z, c0.max.z, c1.min.z, c1.max.z)
//   q3 = (code0, code1, 0, 0) as int32 bit patterns
// code >= 0 : index of an inner pair node;  code < 0 : leaf, ~code = (first_prim << 5) | (mixed << 4) | (count-1)
// An unused child slot has an inverted box (min = +inf, max = -inf) and code = kEmptyLeaf.
struct PairNode {
  float q[16];
};

constexpr int32_t kEmptyLeaf = ~((int32_t)0x7fffffff);  // decodes to no primitives (see builder)

struct BvhBuildResult {
  std::vector<PairNode> nodes;     // nodes[0] is the root pair
  std::vector<int32_t> prim_order; // soup slot -> input primitive index
  int64_t n_leaves = 0;
  int max_depth = 0;
  Aabb bounds;                     // union of all primitive boxes (unpadded)
  double sah_cost = 0.0;
};

inline int32_t leaf_code(int32_t first, int32_t count, bool mixed) {
  return ~((first << 5) | ((mixed ? 1 : 0) << 4) | (count - 1));
}

// prim_boxes: one (unpadded) AABB per primitive.  pad is added on every side of every box before
// the build so that ulp-level differences between the slab test and the reference's hit point can
// never cull a primitive the reference would hit.  kind[i] != 0 marks non-triangle primitives
// (spheres): leaves containing one are flagged "mixed".
BvhBuildResult build_bvh(const std::vector<Aabb> &prim_boxes, const std::vector<uint8_t> &kind,
                         int max_leaf_size, float pad, int threads);

}  // namespace rt
