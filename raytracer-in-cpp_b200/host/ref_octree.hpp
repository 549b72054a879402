// ref_octree.hpp -- the reference's BoxTree as a *candidate filter* for the BVH path.
//
// The reference finds hits only among the faces its octree offers for a ray (BoxTree::intersect,
// src/boxTree.cpp:150-173 of /root/reference).  Its classification of faces into octants is a
// heuristic (the tri-box SAT runs on normalised vectors, :236-240) and its box test produces NaNs
// for rays that lie in a split plane (src/boundingBox.cpp:56-61), so the candidate set is not
// simply "every face the ray could hit": real hits are sometimes missing (e.g. the centre row of an
// even-height image, whose rays have dir.y == 0 on the y = 0 split plane) and degenerate faces
// report phantom hits only where the octree happens to offer them.
//
// To reproduce the reference's image exactly, the device BVH finds triangle hits as usual and then
// asks this structure whether the reference would have had the face among its candidates:
//     candidate(face, ray)  <=>  some octree leaf that lists `face` is reached, i.e. every box on the
//                                path root -> leaf passes BoundingBox::boxIntersect(origin, dest).
// That needs, per face, the leaves listing it (CSR) and, per octree node, its box and parent.
// The build below re-derives the reference octree (same boxes bit for bit, same face lists).
#pragma once

#include <cstdint>
#include <vector>

namespace rt {

struct RefOctree {
  // per node: box and parent (-1 for the root); node 0 is the root
  std::vector<float> box;        // [n][6] min xyz, max xyz
  std::vector<int32_t> parent;   // [n]
  std::vector<uint8_t> is_leaf;  // [n] reachable, non-empty leaf
  // CSR face -> leaves
  std::vector<int32_t> face_off;   // [T+1]
  std::vector<int32_t> face_leaf;  // [refs]
  int64_t n_leaves = 0, n_inner = 0, n_refs = 0, max_leaf = 0;
  bool root_is_leaf = false;  // T <= capacity: every face is a candidate whenever the root box is hit
  int n_nodes() const { return (int)parent.size(); }
};

// verts: [T][3][3] world-space vertices (exactly the floats the reference reads).
// capacity = 1000 and MAX_DEPTH = 15 in the reference (src/flyscene.cpp:86, src/boxTree.cpp:3).
RefOctree build_ref_octree(const float *verts, int32_t n_faces, int capacity, int max_depth, int threads);

}  // namespace rt
