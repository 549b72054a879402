// bvh_builder.cpp -- see bvh_builder.hpp.
#include "bvh_builder.hpp"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <future>
#include <limits>
#include <thread>

namespace rt {

namespace {

constexpr int kBins = 16;
constexpr int kMaxDepth = 60;  // device traversal stack holds 64 entries

struct Box {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
  float mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  void grow(const float *lo, const float *hi) {
    for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], lo[a]); mx[a] = std::max(mx[a], hi[a]); }
  }
  void grow(const Box &b) { grow(b.mn, b.mx); }
  void grow_pt(const float *p) { grow(p, p); }
  float half_area() const {
    float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
    if (dx < 0 || dy < 0 || dz < 0) return 0.f;
    return dx * dy + dy * dz + dz * dx;
  }
};

struct BuildNode {
  Box box;
  int32_t left = -1, right = -1;  // build-node indices
  int32_t first = 0, count = 0;   // leaf range in idx[]
  int32_t depth = 0;
};

struct Builder {
  const std::vector<Aabb> &boxes;  // padded primitive boxes
  std::vector<float> cent;         // [N][3] centroids
  std::vector<int32_t> idx;
  std::vector<BuildNode> nodes;
  std::atomic<int32_t> n_nodes{0};
  int max_leaf;
  int par_depth;  // spawn async tasks above this depth

  Builder(const std::vector<Aabb> &b, int leaf, int threads) : boxes(b), max_leaf(leaf) {
    const size_t N = b.size();
    cent.resize(N * 3);
    idx.resize(N);
    for (size_t i = 0; i < N; ++i) {
      idx[i] = (int32_t)i;
      for (int a = 0; a < 3; ++a) cent[3 * i + a] = 0.5f * (b[i].mn[a] + b[i].mx[a]);
    }
    nodes.resize(std::max<size_t>(1, 2 * N));
    par_depth = 0;
    while ((1 << par_depth) < threads * 2) ++par_depth;
    if (threads <= 1 || N < 50000) par_depth = -1;
  }

  int32_t alloc() { return n_nodes.fetch_add(1); }

  void build(int32_t ni, int32_t lo, int32_t hi, int depth) {
    BuildNode &n = nodes[ni];
    n.depth = depth;
    Box bb, cb;
    for (int32_t k = lo; k < hi; ++k) {
      const int32_t p = idx[k];
      bb.grow(boxes[p].mn, boxes[p].mx);
      cb.grow_pt(&cent[3 * (size_t)p]);
    }
    n.box = bb;
    const int32_t count = hi - lo;
    if (count <= max_leaf) { n.first = lo; n.count = count; return; }
    // Depth guard: SAH splits may be arbitrarily unbalanced; switch to median splits (which halve
    // the count) early enough that the finished tree is never deeper than kMaxDepth.
    int need = 0;
    while ((1 << need) < count) ++need;
    int32_t mid = -1;
    if (depth + need < kMaxDepth - 1) mid = sah_partition(lo, hi, cb);
    if (mid <= lo || mid >= hi) mid = median_partition(lo, hi, cb);
    const int32_t l = alloc(), r = alloc();
    n.left = l; n.right = r;
    if (depth <= par_depth && count > 20000) {
      auto fut = std::async(std::launch::async, [this, l, lo, mid, depth] { build(l, lo, mid, depth + 1); });
      build(r, mid, hi, depth + 1);
      fut.get();
    } else {
      build(l, lo, mid, depth + 1);
      build(r, mid, hi, depth + 1);
    }
  }

  int32_t median_partition(int32_t lo, int32_t hi, const Box &cb) {
    int axis = 0;
    float ext = -1.f;
    for (int a = 0; a < 3; ++a) if (cb.mx[a] - cb.mn[a] > ext) { ext = cb.mx[a] - cb.mn[a]; axis = a; }
    const int32_t mid = lo + (hi - lo) / 2;
    std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, [&](int32_t a, int32_t b) {
      const float ca = cent[3 * (size_t)a + axis], cb2 = cent[3 * (size_t)b + axis];
      return ca < cb2 || (ca == cb2 && a < b);
    });
    return mid;
  }

  int32_t sah_partition(int32_t lo, int32_t hi, const Box &cb) {
    float best_cost = FLT_MAX;
    int best_axis = -1, best_split = -1;
    for (int axis = 0; axis < 3; ++axis) {
      const float cmin = cb.mn[axis], cext = cb.mx[axis] - cb.mn[axis];
      if (!(cext > 0.f)) continue;
      const float scale = (float)kBins / cext;
      Box bin_box[kBins];
      int32_t bin_cnt[kBins] = {0};
      for (int32_t k = lo; k < hi; ++k) {
        const int32_t p = idx[k];
        int b = (int)((cent[3 * (size_t)p + axis] - cmin) * scale);
        b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
        bin_cnt[b]++;
        bin_box[b].grow(boxes[p].mn, boxes[p].mx);
      }
      float right_area[kBins];
      int32_t right_cnt[kBins];
      Box acc;
      int32_t c = 0;
      for (int b = kBins - 1; b > 0; --b) {
        acc.grow(bin_box[b]); c += bin_cnt[b];
        right_area[b] = acc.half_area(); right_cnt[b] = c;
      }
      Box lacc;
      int32_t lc = 0;
      for (int b = 0; b < kBins - 1; ++b) {
        lacc.grow(bin_box[b]); lc += bin_cnt[b];
        if (lc == 0 || right_cnt[b + 1] == 0) continue;
        const float cost = lacc.half_area() * (float)lc + right_area[b + 1] * (float)right_cnt[b + 1];
        if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = b; }
      }
    }
    if (best_axis < 0) return -1;
    const float cmin = cb.mn[best_axis], scale = (float)kBins / (cb.mx[best_axis] - cb.mn[best_axis]);
    auto it = std::partition(idx.begin() + lo, idx.begin() + hi, [&](int32_t p) {
      int b = (int)((cent[3 * (size_t)p + best_axis] - cmin) * scale);
      b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
      return b <= best_split;
    });
    return (int32_t)(it - idx.begin());
  }
};

struct Emitter {
  const Builder &b;
  const std::vector<uint8_t> &kind;
  BvhBuildResult &out;

  static void set_child(PairNode &pn, int c, const Box &bx, int32_t code) {
    if (c == 0) {
      pn.q[0] = bx.mn[0]; pn.q[1] = bx.mx[0]; pn.q[2] = bx.mn[1]; pn.q[3] = bx.mx[1];
      pn.q[8] = bx.mn[2]; pn.q[9] = bx.mx[2];
    } else {
      pn.q[4] = bx.mn[0]; pn.q[5] = bx.mx[0]; pn.q[6] = bx.mn[1]; pn.q[7] = bx.mx[1];
      pn.q[10] = bx.mn[2]; pn.q[11] = bx.mx[2];
    }
    static_assert(sizeof(int32_t) == sizeof(float), "");
    std::memcpy(&pn.q[12 + c], &code, 4);
  }

  int32_t code_of(int32_t ni, int depth) {
    const BuildNode &n = b.nodes[ni];
    out.max_depth = std::max(out.max_depth, depth);
    if (n.left < 0) {
      bool mixed = false;
      for (int32_t k = n.first; k < n.first + n.count; ++k) mixed |= (kind[b.idx[k]] != 0);
      out.n_leaves++;
      out.sah_cost += (double)n.box.half_area() * n.count;
      return leaf_code(n.first, n.count, mixed);
    }
    const int32_t pi = (int32_t)out.nodes.size();
    out.nodes.emplace_back();
    out.sah_cost += (double)n.box.half_area();
    // explicit recursion: depth is bounded by kMaxDepth + log2(16)
    const int32_t c0 = code_of(n.left, depth + 1);
    const int32_t c1 = code_of(n.right, depth + 1);
    PairNode pn{};
    set_child(pn, 0, b.nodes[n.left].box, c0);
    set_child(pn, 1, b.nodes[n.right].box, c1);
    out.nodes[pi] = pn;
    return pi;
  }
};

}  // namespace

BvhBuildResult build_bvh(const std::vector<Aabb> &prim_boxes, const std::vector<uint8_t> &kind,
                         int max_leaf_size, float pad, int threads) {
  BvhBuildResult out;
  const size_t N = prim_boxes.size();
  Box total;
  std::vector<Aabb> padded(N);
  for (size_t i = 0; i < N; ++i) {
    total.grow(prim_boxes[i].mn, prim_boxes[i].mx);
    for (int a = 0; a < 3; ++a) {
      padded[i].mn[a] = prim_boxes[i].mn[a] - pad;
      padded[i].mx[a] = prim_boxes[i].mx[a] + pad;
    }
  }
  for (int a = 0; a < 3; ++a) { out.bounds.mn[a] = total.mn[a]; out.bounds.mx[a] = total.mx[a]; }
  max_leaf_size = std::max(1, std::min(16, max_leaf_size));

  Box empty;  // inverted: never intersected
  const float inf = std::numeric_limits<float>::infinity();
  for (int a = 0; a < 3; ++a) { empty.mn[a] = inf; empty.mx[a] = -inf; }

  if (N == 0) {
    PairNode pn{};
    Emitter::set_child(pn, 0, empty, kEmptyLeaf);
    Emitter::set_child(pn, 1, empty, kEmptyLeaf);
    out.nodes.push_back(pn);
    return out;
  }

  Builder b(padded, max_leaf_size, threads);
  const int32_t root = b.alloc();
  b.build(root, 0, (int32_t)N, 0);
  out.prim_order = b.idx;

  Emitter em{b, kind, out};
  if (b.nodes[root].left < 0) {
    // The whole scene fits one leaf, but the root is a PAIR node.  An empty sibling slot is not an option:
    // the device's min/max slab test is symmetric, so the inverted (+inf, -inf) box of an empty slot passes
    // it, and the traversal would take the empty code for the bottom-of-stack sentinel and stop before the
    // real leaf.  Two primitives or more: split the leaf in two halves (median).  A single primitive:
    // both slots name the same leaf (testing it twice gives the same hit).
    PairNode pn{};
    if (N >= 2) {
      BuildNode &rn = b.nodes[root];
      Box cb;
      for (int32_t k = 0; k < (int32_t)N; ++k) cb.grow_pt(&b.cent[3 * (size_t)b.idx[k]]);
      const int32_t mid = b.median_partition(0, (int32_t)N, cb);
      const int32_t l = b.alloc(), r = b.alloc();
      rn.left = l; rn.right = r;
      b.build(l, 0, mid, 1);
      b.build(r, mid, (int32_t)N, 1);
      out.prim_order = b.idx;
      em.code_of(root, 0);
    } else {
      const int32_t c0 = em.code_of(root, 1);
      Emitter::set_child(pn, 0, b.nodes[root].box, c0);
      Emitter::set_child(pn, 1, b.nodes[root].box, c0);
      out.nodes.push_back(pn);
    }
  } else {
    em.code_of(root, 0);
  }
  const double root_area = std::max(1e-30, (double)b.nodes[root].box.half_area());
  out.sah_cost /= root_area;
  return out;
}

}  // namespace rt
