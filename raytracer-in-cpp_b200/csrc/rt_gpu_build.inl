// rt_gpu_build.inl -- host side of the GPU scene build (kernels: rt_build.cuh).  Included by rt_api.cu.
//
// rt_scene_create for large scenes: the triangle arrays are uploaded once and everything the frame kernels read -- the
// reference octree used as candidate filter (BoxTree::BoxTree / split / clasifyFace, src/boxTree.cpp:11-31, 88-147,
// 203-336 of the reference), the BVH, the 80-byte primitive soup and the 112-byte shading table -- is computed on the
// device.  The host only sequences the launches: one 32-byte read-back per octree level and two small ones per
// BVH level.  Same outputs as the host bake (bake_scene) except for the shape of the BVH, which never changes a frame:
// nearest hits are ordered by (t, face id) and shadow queries are boolean, whatever the traversal order.

namespace {

// bump allocator over a few large device allocations (dozens of cudaMalloc / cudaFree pairs would cost more than the build)
struct BuildArena {
  struct Chunk { char *p; size_t cap, used; };
  std::vector<Chunk> chunks;
  size_t chunk_bytes;
  explicit BuildArena(size_t chunk) : chunk_bytes(chunk) {}
  ~BuildArena() { for (auto &c : chunks) cudaFree(c.p); }
  void reset() { for (auto &c : chunks) c.used = 0; }
  void *alloc(size_t bytes) {
    bytes = (std::max<size_t>(bytes, 16) + 255) & ~(size_t)255;
    for (auto &c : chunks)
      if (c.cap - c.used >= bytes) { void *r = c.p + c.used; c.used += bytes; return r; }
    Chunk c{nullptr, std::max(bytes, chunk_bytes), 0};
    if (cudaMalloc((void **)&c.p, c.cap) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    c.used = bytes;
    chunks.push_back(c);
    return c.p;
  }
  template <class T> T *get(size_t n) { return (T *)alloc(n * sizeof(T)); }
};

#define RT_ARENA(var, arena, T, n)                                                          \
  T *var = (arena).get<T>(n);                                                               \
  if (!var) return fail(RT_ERR_CUDA, "GPU build: device allocation of %zu bytes failed", (size_t)(n) * sizeof(T))

inline unsigned host_f2ord(float f) { uint32_t u; memcpy(&u, &f, 4); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
inline float host_ord2f(unsigned u) { uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u; float f; memcpy(&f, &v, 4); return f; }
inline unsigned cdiv(size_t n, unsigned b) { return (unsigned)((n + b - 1) / b); }

int exclusive_scan(BuildArena &scratch, const int32_t *in, int32_t *out, int n) {
  size_t bytes = 0;
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n));
  void *tmp = scratch.alloc(bytes);
  if (!tmp) return fail(RT_ERR_CUDA, "GPU build: scan workspace");
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, n));
  return RT_OK;
}

struct GpuBuildTimes { float upload_ms = 0, octree_ms = 0, sort_ms = 0, bvh_ms = 0, emit_ms = 0, total_ms = 0; int bvh_levels = 0, oct_levels = 0; };

// Returns RT_OK with *out set, or an error; *too_deep is set (and RT_OK returned with *out == nullptr) when the
// tree is deeper than the traversal stack allows -- the caller then takes the host builder, which bounds
// its depth by construction.
int gpu_build_scene(const RtSceneDesc *desc, RtScene **out, bool *too_deep) {
  using namespace rtb;
  using clk = std::chrono::high_resolution_clock;
  auto ms_since = [](clk::time_point t0) { return std::chrono::duration<float, std::milli>(clk::now() - t0).count(); };
  *out = nullptr;
  *too_deep = false;
  const auto t_start = clk::now();
  const bool trace_on = getenv("RT_BUILD_TRACE") != nullptr;
  auto trace = [&](const char *what) {
    if (trace_on) fprintf(stderr, "[rt build] %8.3f ms  %s\n", std::chrono::duration<float, std::milli>(clk::now() - t_start).count(), what);
  };
  const int T = desc->n_faces, S = desc->n_spheres, N = T + S, M = desc->n_materials;
  if ((long long)N >= (1ll << 26)) return fail(RT_ERR_INVALID, "too many primitives (%d) for the leaf encoding", N);
  GpuBuildTimes tm;

  RtScene *sc = new RtScene();
  struct Guard { RtScene *s; ~Guard() { if (s) rt_scene_destroy(s); } } guard{sc};
  sc->device = g_device;
  sc->built_on_gpu = true;
  DevScene &dv = sc->dev;
  memcpy(dv.model, desc->model_matrix, sizeof(float) * 12);
  dv.n_faces = T; dv.n_spheres = S; dv.n_prims = N;

  BuildArena perm((size_t)N * 460 + (64u << 20)), scratch((size_t)N * 80 + (32u << 20));
  int rc;

  // ---- inputs -> device ----
  RT_ARENA(d_verts, perm, float, (size_t)std::max(T, 1) * 9);
  RT_ARENA(d_fn, perm, float, (size_t)std::max(T, 1) * 3);
  RT_ARENA(d_vn, perm, float, (size_t)std::max(T, 1) * 9);
  RT_ARENA(d_mat, perm, int32_t, (size_t)std::max(T, 1));
  RT_ARENA(d_illum, perm, int32_t, (size_t)M);
  if (T > 0) {
    CUDA_TRY(cudaMemcpyAsync(d_verts, desc->verts, (size_t)T * 36, cudaMemcpyHostToDevice, 0));
    CUDA_TRY(cudaMemcpyAsync(d_fn, desc->face_normals, (size_t)T * 12, cudaMemcpyHostToDevice, 0));
    CUDA_TRY(cudaMemcpyAsync(d_vn, desc->vertex_normals, (size_t)T * 36, cudaMemcpyHostToDevice, 0));
    CUDA_TRY(cudaMemcpyAsync(d_mat, desc->material_id, (size_t)T * 4, cudaMemcpyHostToDevice, 0));
  }
  {
    std::vector<int32_t> illum((size_t)M);
    std::vector<float> mats((size_t)M * 12, 0.f);
    for (int m = 0; m < M; ++m) {
      const RtMaterial &mt = desc->materials[m];
      illum[(size_t)m] = mt.illum;
      float *q = &mats[(size_t)m * 12];
      q[0] = mt.kd[0]; q[1] = mt.kd[1]; q[2] = mt.kd[2]; q[3] = mt.ns;
      q[4] = mt.ks[0]; q[5] = mt.ks[1]; q[6] = mt.ks[2]; q[7] = mt.ni;
      q[8] = bits(mt.illum);
    }
    CUDA_TRY(cudaMemcpy(d_illum, illum.data(), (size_t)M * 4, cudaMemcpyHostToDevice));
    if ((rc = upload(sc->mats, mats.data(), mats.size() * 4))) return rc;
  }
  if ((rc = upload(sc->spheres, desc->spheres, (size_t)S * 16)) || (rc = upload(sc->sphere_mat, desc->sphere_material, (size_t)S * 4)))
    return rc;

  // ---- vertex bounds: reference root box + scene bounds ----
  RT_ARENA(d_ord, perm, unsigned, 8);
  float vmin[3] = {0, 0, 0}, vmax[3] = {0, 0, 0};
  {
    const unsigned init[6] = {host_f2ord(FLT_MAX), host_f2ord(FLT_MAX), host_f2ord(FLT_MAX), host_f2ord(-FLT_MAX), host_f2ord(-FLT_MAX),
                              host_f2ord(-FLT_MAX)};
    unsigned got[6];
    CUDA_TRY(cudaMemcpyAsync(d_ord, init, sizeof(init), cudaMemcpyHostToDevice, 0));
    if (T > 0) k_vertex_bounds<<<std::min(cdiv((size_t)T, 256), 148u * 8u), 256>>>(d_verts, T, d_ord);
    CUDA_TRY(cudaMemcpy(got, d_ord, sizeof(got), cudaMemcpyDeviceToHost));
    for (int a = 0; a < 3; ++a) { vmin[a] = host_ord2f(got[a]); vmax[a] = host_ord2f(got[3 + a]); }
  }
  tm.upload_ms = ms_since(t_start);
  trace("inputs uploaded, vertex bounds");
  float gmin[3] = {1e30f, 1e30f, 1e30f}, gmax[3] = {-1e30f, -1e30f, -1e30f};
  for (int a = 0; a < 3; ++a) {
    // BoundingBox(Mesh&): min starts at FLT_MAX, max at FLT_MIN (the smallest POSITIVE float)
    dv.root_min[a] = T > 0 ? std::min(std::numeric_limits<float>::max(), vmin[a]) : std::numeric_limits<float>::max();
    dv.root_max[a] = T > 0 ? std::max(std::numeric_limits<float>::min(), vmax[a]) : std::numeric_limits<float>::min();
    if (T > 0) { gmin[a] = std::min(gmin[a], vmin[a]); gmax[a] = std::max(gmax[a], vmax[a]); }
  }
  for (int i = 0; i < S; ++i) {
    const float *s = desc->spheres + (size_t)i * 4;
    for (int a = 0; a < 3; ++a) { gmin[a] = std::min(gmin[a], s[a] - s[3]); gmax[a] = std::max(gmax[a], s[a] + s[3]); }
  }
  float diag = 1.f;
  if (N > 0) {
    const float dx = gmax[0] - gmin[0], dy = gmax[1] - gmin[1], dz = gmax[2] - gmin[2];
    diag = std::max(1.f, std::sqrt(dx * dx + dy * dy + dz * dz));
  }
  const float pad = 1e-5f * diag;
  dv.oct_eps = 1e-5f * diag;

  // ---- 1. reference octree, level by level ----
  const auto t_oct = clk::now();
  const int capacity = 1000 /* src/flyscene.cpp:86 */, oct_max_depth = 15 /* MAX_DEPTH, src/boxTree.cpp:3 */;
  bool use_filter = false;
  if (g_opt_ref_candidates && T > 0) {
    if (T <= capacity) {
      sc->oct_stats[0] = 1; sc->oct_stats[1] = 0; sc->oct_stats[2] = T; sc->oct_stats[3] = T;  // the root is the only leaf
    } else {
      use_filter = true;
      RT_ARENA(d_tot, perm, OctTotals, 1);
      RT_ARENA(d_ctr, perm, OctLevelCtr, 1);
      CUDA_TRY(cudaMemsetAsync(d_tot, 0, sizeof(OctTotals), 0));
      RT_ARENA(level, perm, OctLevelNode, 1);
      {
        OctLevelNode root{};
        memcpy(root.mn, dv.root_min, 12); memcpy(root.mx, dv.root_max, 12);
        root.node_id = 0;
        CUDA_TRY(cudaMemcpy(level, &root, sizeof(root), cudaMemcpyHostToDevice));
      }
      RT_ARENA(pairs, perm, int2, (size_t)T);
      k_iota_pairs<<<cdiv((size_t)T, 256), 256>>>(pairs, (unsigned)T);
      unsigned n_level = 1, n_pairs = (unsigned)T;
      int n_nodes = 1, depth = oct_max_depth;
      struct LevelOut { const float4 *box; unsigned n_nodes; const int2 *refs; unsigned n_refs; };
      std::vector<LevelOut> outs;
      size_t total_refs = 0;
      while (n_pairs > 0 && n_level > 0) {
        scratch.reset();
        const unsigned n_children = n_level * 8u;
        const size_t out_cap = (size_t)n_pairs * 8;
        if (out_cap >= (1ull << 31)) return fail(RT_ERR_LIMIT, "GPU build: octree level with %u memberships", n_pairs);
        RT_ARENA(child_cnt, scratch, unsigned, n_children);
        RT_ARENA(memb, scratch, int2, out_cap);
        RT_ARENA(flag_node, scratch, int32_t, n_children);
        RT_ARENA(flag_split, scratch, int32_t, n_children);
        RT_ARENA(pos_node, scratch, int32_t, n_children);
        RT_ARENA(pos_split, scratch, int32_t, n_children);
        RT_ARENA(child_state, scratch, int32_t, n_children);
        RT_ARENA(level_box, perm, float4, (size_t)n_children * 2);
        RT_ARENA(next_level, perm, OctLevelNode, n_children);
        CUDA_TRY(cudaMemsetAsync(child_cnt, 0, (size_t)n_children * 4, 0));
        CUDA_TRY(cudaMemsetAsync(d_ctr, 0, sizeof(OctLevelCtr), 0));
        k_oct_classify<<<cdiv(out_cap, 256), 256>>>(d_verts, pairs, n_pairs, level, child_cnt, memb, (unsigned)out_cap, d_ctr);
        k_oct_flags<<<cdiv(n_children, 256), 256>>>(n_children, child_cnt, capacity, depth, flag_node, flag_split);
        if ((rc = exclusive_scan(scratch, flag_node, pos_node, (int)n_children)) ||
            (rc = exclusive_scan(scratch, flag_split, pos_split, (int)n_children)))
          return rc;
        k_oct_decide<<<cdiv(n_children, 256), 256>>>(level, n_children, child_cnt, capacity, depth, flag_node, pos_node, flag_split,
                                                   pos_split, n_nodes, child_state, level_box, next_level, d_ctr, d_tot);
        OctLevelCtr ctr;
        CUDA_TRY(cudaMemcpy(&ctr, d_ctr, sizeof(ctr), cudaMemcpyDeviceToHost));
        if (ctr.overflow) return fail(RT_ERR_CUDA, "GPU build: octree membership buffer overflow");
        RT_ARENA(refs, perm, int2, std::max<size_t>(ctr.n_leaf_refs, 1));
        RT_ARENA(next_pairs, perm, int2, std::max<size_t>(ctr.n_next_pairs, 1));
        if (ctr.n_out > 0)
          k_oct_route<<<cdiv(ctr.n_out, 256), 256>>>(memb, ctr.n_out, child_state, refs, ctr.n_leaf_refs, next_pairs, ctr.n_next_pairs, d_ctr);
        outs.push_back({level_box, ctr.n_new_nodes, refs, ctr.n_leaf_refs});
        total_refs += ctr.n_leaf_refs;
        n_nodes += (int)ctr.n_new_nodes;
        level = next_level; n_level = ctr.n_split;
        pairs = next_pairs; n_pairs = ctr.n_next_pairs;
        --depth;
        ++tm.oct_levels;
      }
      if (total_refs >= (1ull << 31)) return fail(RT_ERR_LIMIT, "GPU build: %zu octree references", total_refs);
      // node boxes + parent links, concatenated in level order (node 0 = root, parent -1)
      if ((rc = sc->oct_box.reserve((size_t)n_nodes * 32))) return rc;
      {
        const float root_box[8] = {dv.root_min[0], dv.root_min[1], dv.root_min[2], bits(-1), dv.root_max[0], dv.root_max[1], dv.root_max[2], 0.f};
        CUDA_TRY(cudaMemcpyAsync(sc->oct_box.p, root_box, 32, cudaMemcpyHostToDevice, 0));
        size_t at = 1;
        for (const LevelOut &lo : outs) {
          if (lo.n_nodes) CUDA_TRY(cudaMemcpyAsync((char *)sc->oct_box.p + at * 32, lo.box, (size_t)lo.n_nodes * 32, cudaMemcpyDeviceToDevice, 0));
          at += lo.n_nodes;
        }
      }
      // (face, leaf) references -> CSR by face through one radix sort (deterministic row order)
      scratch.reset();
      const unsigned R = (unsigned)total_refs;
      if ((rc = sc->oct_face_off.reserve((size_t)(T + 1) * 4)) || (rc = sc->oct_face_leaf.reserve(std::max<size_t>(R, 1) * 4))) return rc;
      CUDA_TRY(cudaMemsetAsync(sc->oct_face_off.p, 0, (size_t)(T + 1) * 4, 0));
      if (R > 0) {
        RT_ARENA(keys_a, scratch, unsigned long long, R);
        RT_ARENA(keys_b, scratch, unsigned long long, R);
        size_t at = 0;
        for (const LevelOut &lo : outs) {
          if (lo.n_refs) k_pack_ref_keys<<<cdiv(lo.n_refs, 256), 256>>>(lo.refs, lo.n_refs, keys_a + at);
          at += lo.n_refs;
        }
        cub::DoubleBuffer<unsigned long long> kb(keys_a, keys_b);
        size_t bytes = 0;
        CUDA_TRY(cub::DeviceRadixSort::SortKeys(nullptr, bytes, kb, (int)R, 0, 64));
        void *tmp = scratch.alloc(bytes);
        if (!tmp) return fail(RT_ERR_CUDA, "GPU build: sort workspace");
        CUDA_TRY(cub::DeviceRadixSort::SortKeys(tmp, bytes, kb, (int)R, 0, 64));
        k_face_csr<<<cdiv(R, 256), 256>>>(kb.Current(), R, T, sc->oct_face_off.as<int32_t>(), sc->oct_face_leaf.as<int32_t>());
      }
      OctTotals tot;
      CUDA_TRY(cudaMemcpy(&tot, d_tot, sizeof(tot), cudaMemcpyDeviceToHost));
      sc->oct_stats[0] = tot.n_leaves; sc->oct_stats[1] = 1 + (int64_t)tot.n_inner; sc->oct_stats[2] = (int64_t)tot.n_refs;
      sc->oct_stats[3] = tot.max_leaf;
      dv.oct_box = sc->oct_box.as<float4>();
      dv.oct_face_off = sc->oct_face_off.as<int32_t>();
      dv.oct_face_leaf = sc->oct_face_leaf.as<int32_t>();
    }
  }
  tm.octree_ms = ms_since(t_oct);
  trace("octree");
  sc->octree_ms = tm.octree_ms;

  // ---- 2. primitive boxes, Morton order ----
  const auto t_sort = clk::now();
  scratch.reset();
  RT_ARENA(d_boxes, perm, float, (size_t)N * 6);
  k_prim_boxes<<<cdiv((size_t)N, 256), 256>>>(d_verts, T, sc->spheres.as<float>(), S, use_filter ? dv.oct_box : nullptr,
                                             use_filter ? dv.oct_face_off : nullptr, use_filter ? dv.oct_face_leaf : nullptr, pad, d_boxes);
  RT_ARENA(mk_a, scratch, unsigned long long, (size_t)N);
  RT_ARENA(mk_b, scratch, unsigned long long, (size_t)N);
  RT_ARENA(id_a, perm, int32_t, (size_t)N);
  RT_ARENA(id_b, perm, int32_t, (size_t)N);
  {
    Bounds6 sb;
    for (int a = 0; a < 3; ++a) { sb.lo[a] = gmin[a]; sb.hi[a] = gmax[a]; }
    k_morton<<<cdiv((size_t)N, 256), 256>>>(d_boxes, N, sb, mk_a, id_a);
  }
  cub::DoubleBuffer<unsigned long long> mk(mk_a, mk_b);
  cub::DoubleBuffer<int32_t> ids(id_a, id_b);
  {
    size_t bytes = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, mk, ids, N, 0, 63));
    void *tmp = scratch.alloc(bytes);
    if (!tmp) return fail(RT_ERR_CUDA, "GPU build: sort workspace");
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, bytes, mk, ids, N, 0, 63));
  }
  const int32_t *sorted_ids = ids.Current();
  tm.sort_ms = ms_since(t_sort);
  trace("boxes, Morton sort (queued)");

  // ---- 3. BVH: binned SAH splits, one launch pair (big nodes / small nodes) per tree level ----
  const auto t_bvh = clk::now();
  const int leaf = std::max(1, std::min(16, g_opt_leaf));
  const size_t NN = (size_t)2 * N;
  SahNodes nd;
  nd.lo = perm.get<int32_t>(NN); nd.cnt = perm.get<int32_t>(NN); nd.left = perm.get<int32_t>(NN); nd.depth = perm.get<int32_t>(NN);
  nd.box = perm.get<float>(NN * 6);
  if (!nd.lo || !nd.cnt || !nd.left || !nd.depth || !nd.box) return fail(RT_ERR_CUDA, "GPU build: node arrays");
  if ((rc = sc->prim_order.reserve((size_t)N * 4))) return rc;
  int32_t *prim_order = sc->prim_order.as<int32_t>();
  RT_ARENA(act_a, perm, int32_t, (size_t)N);
  RT_ARENA(act_b, perm, int32_t, (size_t)N);
  RT_ARENA(sel_big, perm, int32_t, (size_t)N);
  RT_ARENA(sel_small, perm, int32_t, (size_t)N);
  RT_ARENA(fl_a, perm, int32_t, (size_t)N);
  RT_ARENA(fl_b, perm, int32_t, (size_t)N);
  RT_ARENA(ps_a, perm, int32_t, (size_t)N);
  RT_ARENA(ps_b, perm, int32_t, (size_t)N);
  RT_ARENA(d_tot3, perm, int32_t, 4);
  int n_nodes_tree = 1;
  {
    // index buffers: `ids` (Morton order) is level 0's input, the other half of the double buffer its output
    int32_t *in = ids.Current(), *outb = ids.Alternate();
    int32_t *act = act_a, *act_next = act_b;
    k_sah_root<<<1, 1>>>(nd, N, act);
    int n_act = 1, level_first = 0;
    while (n_act > 0) {
      scratch.reset();
      // which active nodes are big (one 1024-thread block each), which small (one warp each)
      k_sah_kind<<<cdiv((size_t)n_act, 256), 256>>>(act, n_act, nd, fl_a, fl_b);
      if ((rc = exclusive_scan(scratch, fl_a, ps_a, n_act)) || (rc = exclusive_scan(scratch, fl_b, ps_b, n_act))) return rc;
      k_sah_select<<<cdiv((size_t)n_act, 256), 256>>>(n_act, fl_a, ps_a, ps_b, sel_big, sel_small, d_tot3);
      int tot[3] = {0, 0, 0};
      CUDA_TRY(cudaMemcpy(tot, d_tot3, sizeof(tot), cudaMemcpyDeviceToHost));
      const int n_big = tot[1], n_small = tot[2], child_base = n_nodes_tree;
      if (n_big > 0) k_sah_split<1024><<<n_big, 1024>>>(sel_big, act, child_base, leaf, d_boxes, in, outb, prim_order, nd);
      if (n_small > 0) k_sah_split<32><<<n_small, 32>>>(sel_small, act, child_base, leaf, d_boxes, in, outb, prim_order, nd);
      // the children: next level's active list
      const int n_children = 2 * n_act;
      level_first = child_base;
      n_nodes_tree += n_children;
      k_sah_flags<<<cdiv((size_t)n_children, 256), 256>>>(level_first, n_children, nd, leaf, fl_a);
      if ((rc = exclusive_scan(scratch, fl_a, ps_a, n_children))) return rc;
      k_sah_lists<<<cdiv((size_t)n_children, 256), 256>>>(level_first, n_children, fl_a, ps_a, act_next, d_tot3);
      CUDA_TRY(cudaMemcpy(tot, d_tot3, 4, cudaMemcpyDeviceToHost));
      n_act = tot[0];
      std::swap(act, act_next);
      std::swap(in, outb);
      ++tm.bvh_levels;
      if (tm.bvh_levels > kSahMaxDepth + 8) return fail(RT_ERR_CUDA, "GPU build: tree deeper than %d levels", kSahMaxDepth + 8);
    }
  }
  tm.bvh_ms = ms_since(t_bvh);
  trace("BVH levels");

  // ---- 4. pair nodes ----
  const auto t_emit = clk::now();
  scratch.reset();
  RT_ARENA(pflag, scratch, int32_t, (size_t)n_nodes_tree);
  RT_ARENA(ppos, scratch, int32_t, (size_t)n_nodes_tree);
  k_pair_flags<<<cdiv((size_t)n_nodes_tree, 256), 256>>>(n_nodes_tree, nd, leaf, pflag);
  if ((rc = exclusive_scan(scratch, pflag, ppos, n_nodes_tree))) return rc;
  int last_pos = 0, last_flag = 0;
  CUDA_TRY(cudaMemcpy(&last_pos, ppos + n_nodes_tree - 1, 4, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(&last_flag, pflag + n_nodes_tree - 1, 4, cudaMemcpyDeviceToHost));
  const int n_pairs_out = last_pos + last_flag;
  trace("pair flags scanned");
  if ((rc = sc->nodes.reserve((size_t)std::max(n_pairs_out, 1) * 64))) return rc;
  trace("nodes allocated");
  RT_ARENA(d_leaves, perm, unsigned, 4);
  RT_ARENA(d_depth, perm, int32_t, 4);
  RT_ARENA(d_sah, perm, double, 2);
  CUDA_TRY(cudaMemsetAsync(d_leaves, 0, 16, 0));
  CUDA_TRY(cudaMemsetAsync(d_depth, 0, 16, 0));
  CUDA_TRY(cudaMemsetAsync(d_sah, 0, 16, 0));
  k_emit_pairs<<<cdiv((size_t)n_nodes_tree, 256), 256>>>(n_nodes_tree, nd, leaf, T, pflag, ppos, prim_order, sc->nodes.as<float4>(), d_leaves,
                                                       d_depth, d_sah);
  int tree_depth = 0;
  CUDA_TRY(cudaMemcpy(&tree_depth, d_depth, 4, cudaMemcpyDeviceToHost));
  if (tree_depth > RT_STACK_SIZE - 3) { *too_deep = true; return RT_OK; }  // (cannot happen: the builder guards its depth)

  // ---- 5. bake ----
  trace("pairs emitted (depth read back)");
  if ((rc = sc->prims.reserve((size_t)N * 80)) || (rc = sc->shade.reserve((size_t)std::max(T, 1) * 112))) return rc;
  trace("soup + shading table allocated");
  k_bake_prims<<<cdiv((size_t)N, 256), 256>>>(N, T, prim_order, d_verts, d_fn, d_mat, d_illum, sc->spheres.as<float>(),
                                             sc->sphere_mat.as<int32_t>(), sc->prims.as<float4>());
  if (T > 0) k_bake_shade<<<cdiv((size_t)T, 256), 256>>>(T, d_verts, d_fn, d_vn, d_mat, sc->shade.as<float4>());
  RT_ARENA(d_rb, perm, float, 8);
  k_root_bounds<<<1, 1>>>(sc->nodes.as<float4>(), d_rb);
  float rb[6];
  unsigned n_leaves = 0;
  double sah[2] = {0, 0}, root_area = 0;
  CUDA_TRY(cudaMemcpy(rb, d_rb, sizeof(rb), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(&n_leaves, d_leaves, 4, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(sah, d_sah, 8, cudaMemcpyDeviceToHost));
  {
    const double dx = rb[3] - rb[0], dy = rb[4] - rb[1], dz = rb[5] - rb[2];
    root_area = std::max(1e-30, dx * dy + dy * dz + dz * dx);
  }
  CUDA_TRY(cudaGetLastError());
  trace("baked, results read back");
  memcpy(dv.bvh_min, rb, 12);
  memcpy(dv.bvh_max, rb + 3, 12);
  dv.n_nodes = n_pairs_out;
  dv.nodes = sc->nodes.as<float4>();
  dv.prims = sc->prims.as<float4>();
  dv.shade = sc->shade.as<float4>();
  dv.mats = sc->mats.as<float4>();
  dv.spheres = sc->spheres.as<float4>();
  dv.sphere_mat = sc->sphere_mat.as<int32_t>();
  sc->n_leaves = n_leaves;
  sc->bvh_depth = tree_depth;
  sc->sah_cost = sah[0] / root_area;
  tm.emit_ms = ms_since(t_emit);

  if ((rc = sc->frame_counts.reserve(sizeof(FrameCounts))) || (rc = sc->frame_params.reserve(sizeof(FrameParams)))) return rc;
  if (cudaMallocHost((void **)&sc->h_counts, sizeof(FrameCounts)) != cudaSuccess ||
      cudaMallocHost((void **)&sc->h_fcounts, sizeof(FusedCounts)) != cudaSuccess ||
      cudaStreamCreateWithFlags(&sc->stream, cudaStreamNonBlocking) != cudaSuccess)
    return fail(RT_ERR_CUDA, "cudaMallocHost failed");
  trace("workspace (pinned counters, stream)");
  tm.total_ms = ms_since(t_start);
  sc->build_ms = tm.total_ms;
  sc->build_phase_ms[0] = tm.upload_ms; sc->build_phase_ms[1] = tm.octree_ms; sc->build_phase_ms[2] = tm.sort_ms;
  sc->build_phase_ms[3] = tm.bvh_ms; sc->build_phase_ms[4] = tm.emit_ms;
  sc->build_rounds[0] = tm.oct_levels; sc->build_rounds[1] = tm.bvh_levels;
  {
    std::lock_guard<std::mutex> lk(g_scenes_mu);
    g_scenes.push_back(sc);
  }
  guard.s = nullptr;
  *out = sc;
  return RT_OK;
}

}  // namespace
