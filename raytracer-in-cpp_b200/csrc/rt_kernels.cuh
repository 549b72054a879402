// rt_kernels.cuh -- the four kernel families of the B200 render path (sm_100a).
//
//   K1 k_trace_nearest  persistent-thread nearest-hit wavefront kernel (ray generation fused for
//                       primary rays)            <- raytraceScene pixel loop src/flyscene.cpp:573-598,
//                                                   traceRay nearest hit :655-695, BoxTree::intersect
//   K2 k_shadow         any-hit shadow rays for the gate test and every area-light grid sample
//                                                <- lightStrikes :912-954, createSpherePoint :962-972
//   K3 k_shade          Phong + material switch + ballot/popc ray compaction into the next bounce
//                       queue, framebuffer write for rays that terminate at level 0
//                                                <- phongShade :822-859, material switch :712-760
//   K3b k_fold          folds the per-level records back to the pixel in the reference's own
//                       evaluation order and writes the packed uchar4 framebuffer
//                                                <- the unwinding of traceRay's recursion + ppmIO.hpp:145
// (citations relative to /root/reference).  Compiled with -fmad=false, see rt_device.cuh.
#pragma once

#include "rt_device.cuh"

namespace rtd {

// record types written by K3 and consumed by K3b

enum RecType : uint8_t {
  REC_TERMINAL = 0,        // rec.xyz is the final colour of this ray
  REC_MIRROR = 1,          // 0.15*P + 0.85*child                  illum 3,4   :738
  REC_MIRROR_FRESNEL = 2,  // f * (0.15*P + 0.85*child)            illum 5     :738-743
  REC_GLASS9 = 3,          // 0.10*P + 0.90*child                  illum 9     :718
  REC_REFRACT6 = 4         // 0.2*P + 0.8*child                    illum 6     :754
};

#define RT_MAX_LEVELS 68  // guard depth 64 + slack

struct LevelBufs {
  float4 *ray_o;      // (o.xyz, bits(flags)), flags bit0: light list = single point lp
  float4 *ray_d;      // (d.xyz, lp.x)
  float2 *ray_l;      // (lp.y, lp.z)
  float *hit_t;
  int32_t *hit_face;  // -1: no hit
  int32_t *hit_list;  // compacted indices of the rays that hit something (K1 -> K2, K3)
  float4 *hit_p;      // per hit slot: (hit point, bits((ray index << 1) | single-light flag)) (K1 -> K2)
  int32_t *lit_list;  // area mode: hit slots whose gate passed, i.e. the hits that shoot sample rays (K2 pass 1 -> pass 2)
  uint8_t *vis;       // [hit slot][J] visibility of every shadow job
  float4 *rec;        // (P or colour, fresnel factor)
  int32_t *child;     // slot of the child ray in the next level, -1 if none
  uint8_t *type;      // RecType
};

struct Counters {
  unsigned long long shadow_rays;         // reference census (K3)
  unsigned long long shadow_rays_traced;  // any-hit queries K2 actually traced (stats builds)
  unsigned long long secondary_rays;
  unsigned long long box_tests;      // K1 (nearest hit)
  unsigned long long tri_tests;      // K1
  unsigned long long shade_samples;  // K3
  unsigned long long box_tests_k2;   // K2 (shadow)
  unsigned long long tri_tests_k2;   // K2
  unsigned long long filter_checks, filter_slow, filter_rejects;  // candidate filter (K1 + K2)
};

// Device-resident per-frame state, zeroed by one memset at frame start.  Ray and hit counts of every
// bounce level live here, so no kernel launch depends on a host read-back.
struct FrameCounts {
  unsigned long long work_k1[RT_MAX_LEVELS];  // persistent-kernel work cursors
  unsigned long long work_k2[RT_MAX_LEVELS];
  int32_t n_rays[RT_MAX_LEVELS];              // rays queued at level k (k >= 1; level 0 comes as a parameter)
  int32_t n_hits[RT_MAX_LEVELS];              // rays of level k that hit a primitive
  int32_t n_lit[RT_MAX_LEVELS];               // ... of which the gate passed (area mode; K2 pass 1)
  uint32_t k3_done[RT_MAX_LEVELS];            // CTAs of K3(level k) that have finished (last one sets the graph condition)
  Counters ctr;
};

__device__ __forceinline__ int global_row(const FrameParams &fp, int local_row) {
  if (fp.band_world <= 1) return local_row;
  const int band = local_row / fp.band_rows, within = local_row - band * fp.band_rows;
  return (band * fp.band_world + fp.band_rank) * fp.band_rows + within;
}

// Framebuffer slot of local pixel i.  Normally the output holds only this rank's rows, packed; with
// fp.out_full_frame the output is the whole H x W frame (possibly a peer GPU's memory mapped over
// NVLink) and every row goes to its global position, which removes the separate tile gather.
__device__ __forceinline__ size_t fb_index(const FrameParams &fp, int i) {
  if (!fp.out_full_frame) return (size_t)i;
  const int row = i / fp.width, col = i - row * fp.width;
  return (size_t)global_row(fp, row) * fp.width + col;
}

// ---------------------------------------------------------------------------------------------
// K1: nearest hit.  Persistent CTAs; each warp pulls 32 work items at a time from a global cursor.
// Primary rays: one item = one pixel of an 8x4 tile (coherent warps); the ray is generated
// in-kernel and stored to the level-0 queue.  Rays that miss are finished here (BACKGROUND record,
// level-0 pixel written); rays that hit are compacted into hit_list with one atomic per warp.
// n0 >= 0: item count given by the host (level 0); n0 < 0: read fc->n_rays[level].
// ---------------------------------------------------------------------------------------------
// Work items a warp takes from the global cursor per atomic.  Small when the launch has little work
// (so that every resident warp gets some), up to 96 when there is plenty (atomic off the critical path).
__device__ __forceinline__ unsigned long long pool_batch(unsigned long long n_items) {
  const unsigned long long warps_total = (unsigned long long)gridDim.x * (blockDim.x >> 5);
  const unsigned long long per_warp = n_items / (warps_total * 8ull);
  return per_warp >= 96ull ? 96ull : (per_warp >= 64ull ? 64ull : 32ull);
}

template <bool PRIMARY, bool STATS, bool PLAIN>
__global__ void __launch_bounds__(128, 7) k_trace_nearest(const DevScene sc, const FrameParams *__restrict__ fpp,
                                                      const LevelBufs lv, const int level, const int n0,
                                                      FrameCounts *fc) {
  RT_STAGE_FRAME_PARAMS(fpp);
  int32_t *face_out = level == 0 ? fp.out_face : nullptr;
  float *t_out = level == 0 ? fp.out_t : nullptr;
  uchar4 *fb = level == 0 ? fp.out_rgba : nullptr;
  float *rgb_f32 = level == 0 ? fp.out_rgbf : nullptr;
  const int lane = threadIdx.x & 31;
  TravStats st; st.box_tests = 0; st.tri_tests = 0; st.filter_checks = 0; st.filter_slow = 0; st.filter_rejects = 0;
  const int twl = fp.tile_w_log2, tile_w = 1 << twl, tile_h = 32 >> twl;
  const int tiles_x = PRIMARY ? (fp.width + tile_w - 1) >> twl : 1;
  const int n = n0 >= 0 ? n0 : fc->n_rays[level];
  const unsigned long long n_items =
      PRIMARY ? (unsigned long long)tiles_x * (unsigned long long)((fp.local_rows + tile_h - 1) / tile_h) * 32ull : (unsigned long long)n;
  unsigned long long *cursor = &fc->work_k1[level];
  const unsigned long long batch = pool_batch(n_items);
  unsigned long long pool_next = 0, pool_end = 0;  // warp-local batch of work items (warp-uniform)

  Trav<false, STATS, PLAIN> tr;
  int stack[RT_STACK_SIZE];
  tr.idle();
  // work donation between the lanes of a warp (Trav::run_split): pays for shadow rays of small launches (K2), not
  // for nearest-hit rays (C3 K1 0.272 -> 0.30 ms, C5 K1 2.10 -> 2.40 ms), so it stays off here unless forced (>= 100)
  const bool split = !PLAIN && sc.split_min >= 100;

  for (;;) {
    if (pool_next >= pool_end) {
      // one global atomic per `batch` items; the warp then walks the batch 32 items at a time
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(cursor, batch);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base >= n_items) break;
      pool_next = base;
      pool_end = base + batch < n_items ? base + batch : n_items;
    }
    const unsigned long long item = pool_next + (unsigned long long)lane;
    pool_next += 32ull;
    int i = 0, single = 0;
    bool valid = false, active = false;
    float fin_t = RT_NO_HIT_T;
    int fin_id = -1;
    if (item < pool_end) {
      V3 o, d;
      V3 screen = mk(0.f, 0.f, 0.f);
      bool tri_enabled = true;
      if constexpr (PRIMARY) {
        const int tile = (int)(item >> 5), in_tile = (int)(item & 31);
        int tx, ty;
        tile_xy(fp, tile, tiles_x, (fp.local_rows + tile_h - 1) / tile_h, tx, ty);
        const int px = (tx << twl) + (in_tile & (tile_w - 1)), py = ty * tile_h + (in_tile >> twl);
        valid = px < fp.width && py < fp.local_rows;
        i = py * fp.width + px;
        o = ld3(fp.eye);
        d = mk(0.f, 0.f, 0.f);
        if (tile_outside(fp, tx, ty)) {  // (warp-uniform) the whole tile is BACKGROUND, src/flyscene.cpp:658-665
          if (valid) {
            lv.rec[i] = make_float4(1.f, 1.f, 1.f, 1.f);
            lv.type[i] = (uint8_t)REC_TERMINAL;
            if (fb) fb[fb_index(fp, i)] = make_uchar4(255, 255, 255, 255);
            if (rgb_f32) { rgb_f32[3 * (size_t)i] = 1.f; rgb_f32[3 * (size_t)i + 1] = 1.f; rgb_f32[3 * (size_t)i + 2] = 1.f; }
            if (face_out) face_out[i] = -1;
            if (t_out) t_out[i] = RT_NO_HIT_T;
          }
          continue;
        }
        if (valid) {
          screen = screen_to_world(fp, (float)px, (float)global_row(fp, py));
          d = sub(screen, o);  // :619, not normalised
        }
      } else {
        i = (int)item;
        valid = true;
        const float4 ro = lv.ray_o[i], rd = lv.ray_d[i];
        o = mk(ro); d = mk(rd);
        single = __float_as_int(ro.w) & 1;
      }
      if (valid) {
        // one exact reciprocal of the direction serves both root-box tests and the traversal
        const V3 rdir = recip_dir(d);
        if constexpr (PRIMARY) {
          // raytraceScene's root-box pre-cull on (origin, screen), src/flyscene.cpp:576
          // (scenes with analytic spheres -- not a reference feature -- skip it, like rt_oracle.c)
          tri_enabled = ref_box_intersect_quick(sc.root_min, sc.root_max, o, screen, rdir) || (!PLAIN && sc.n_spheres > 0);
        }
        // traceRay's own root test on (origin, origin+direction), src/flyscene.cpp:655
        // (the reference divides by (o + d) - o, which differs from d by up to ulp(|o|) / |d|: more than the
        // margin of the decisive test for a slow, far-away secondary ray, so that test gets its own reciprocal)
        const V3 dest = add(o, d);
        tri_enabled = tri_enabled && ref_box_intersect_quick(sc.root_min, sc.root_max, o, dest, recip_dir(sub(dest, o)));
        if (tri_enabled || (!PLAIN && sc.n_spheres > 0)) {  // else: missed the root box, BACKGROUND
          tr.init(o, d, dest, tri_enabled, rdir);
          active = true;
          // work donation may reuse this lane's ray registers once its own ray is done: keep the ray in the queue
          if (PRIMARY && split) {
            lv.ray_o[i] = make_float4(o.x, o.y, o.z, __int_as_float(0));
            lv.ray_d[i] = make_float4(d.x, d.y, d.z, 0.f);
          }
        }
      }
    }
    // ---- traverse: the whole warp, lanes without a ray idle (or help, run_split) inside ----
    if (split) {
      tr.run_split(sc, st, stack, active, sc.split_min - 100);
      if (active) { fin_t = tr.fin_t; fin_id = tr.fin_id; }
    } else {
      for (;;) {
        tr.run(sc, st, stack, 0xffffffffu);
        if (active && tr.finish(sc, st)) {
          active = false;
          fin_t = tr.best_t;
          fin_id = tr.best_id;
        }
        if (!__any_sync(0xffffffffu, active)) break;
      }
    }
    // ---- retire (all 32 lanes converge here: one hit-list atomic per warp) ----
    const bool hit = valid && fin_id >= 0;
    const int slot = warp_append(&fc->n_hits[level], hit);
    if (valid) {
      if (hit) {
        V3 ro = tr.o, rd = tr.d;
        if (split) {
          ro = mk(lv.ray_o[i]); rd = mk(lv.ray_d[i]);
        } else if (PRIMARY) {
          lv.ray_o[i] = make_float4(tr.o.x, tr.o.y, tr.o.z, __int_as_float(0));
          lv.ray_d[i] = make_float4(tr.d.x, tr.d.y, tr.d.z, 0.f);
        }
        lv.hit_t[i] = fin_t;
        lv.hit_face[i] = fin_id;
        lv.hit_list[slot] = i;
        const V3 hit = add(ro, mul(fin_t, rd));  // src/flyscene.cpp:695
        lv.hit_p[slot] = make_float4(hit.x, hit.y, hit.z, __int_as_float((i << 1) | single));
      } else {
        // BACKGROUND, src/flyscene.cpp:658-665 / :684-691
        lv.rec[i] = make_float4(1.f, 1.f, 1.f, 1.f);
        lv.type[i] = (uint8_t)REC_TERMINAL;
        if (level == 0) {
          if (fb) fb[fb_index(fp, i)] = make_uchar4(255, 255, 255, 255);
          if (rgb_f32) { rgb_f32[3 * (size_t)i] = 1.f; rgb_f32[3 * (size_t)i + 1] = 1.f; rgb_f32[3 * (size_t)i + 2] = 1.f; }
        }
      }
      if (level == 0) {
        if (face_out) face_out[i] = fin_id;
        if (t_out) t_out[i] = fin_t;
      }
    }
  }
  if (STATS) {
    warp_sum_add(&fc->ctr.box_tests, st.box_tests);
    warp_sum_add(&fc->ctr.tri_tests, st.tri_tests);
    warp_sum_add(&fc->ctr.filter_checks, st.filter_checks);
    warp_sum_add(&fc->ctr.filter_slow, st.filter_slow);
    warp_sum_add(&fc->ctr.filter_rejects, st.filter_rejects);
  }
}

// ---------------------------------------------------------------------------------------------
// light list / sample generation shared by K2 and K3
// ---------------------------------------------------------------------------------------------
struct RayLights {
  int n;        // lights visible to this ray's shading (scene lights, or the single inherited point)
  bool single;
  V3 lp;
};
__device__ __forceinline__ RayLights ray_lights(const FrameParams &fp, float4 ro, float4 rd, float2 rl) {
  RayLights r;
  r.single = (__float_as_int(ro.w) & 1) != 0;
  r.n = r.single ? 1 : fp.n_lights;
  r.lp = mk(rd.w, rl.x, rl.y);
  return r;
}
__device__ __forceinline__ V3 light_pos(const FrameParams &fp, const RayLights &rl, int l) {
  return rl.single ? rl.lp : ld3(fp.lights + 3 * l);
}
// arealight::getPointLights (arealight.hpp:15-25) through createAreaLight(light, 0.3, 0.15, u, v)
// (src/flyscene.cpp:956-972): sample k = i*vsteps + j, i outer.
__device__ __forceinline__ V3 area_sample(const FrameParams &fp, V3 c, int k) {
  // spherical mode (:974-993): Vector3f(x, y, z) / 5 + lightPoint, offsets from the host
  if (fp.sphere_mode) return add(ld3(fp.sphere_off + 3 * k), c);
  const int i = k / fp.vsteps, j = k - i * fp.vsteps;
  const float ux = c.x + fp.area_len_x * 1.0f;  // uvec = corner + lengthX*(1,0,0)
  const float vy = c.y + fp.area_len_y * 1.0f;  // vvec = corner + lengthY*(0,1,0)
  const float uz = c.z + fp.area_len_x * 0.0f;
  const float sx = (float)((double)i + 0.5) * (ux / (float)fp.usteps);
  const float sy = (float)((double)j + 0.5) * (vy / (float)fp.vsteps);
  return mk(sx, sy, uz);
}

// sample s of light l as seen by a ray: from the per-frame table for scene lights, computed for the
// single inherited light of mirror children
__device__ __forceinline__ V3 light_sample(const FrameParams &fp, const RayLights &rl, int l, int s, int S) {
  if (!rl.single && fp.have_sample_table) return ld3(fp.sample_table + 3 * (l * S + s));
  return area_sample(fp, light_pos(fp, rl, l), s);
}

// ---------------------------------------------------------------------------------------------
// K2: shadow rays for the rays of hit_list.  Job (slot, j) writes vis[slot * J + j]:
//   j < Lmax : gate ray towards light j (lightStrikes of traceRay, src/flyscene.cpp:699);
//   j >= Lmax: sample ray (light (j-Lmax)/S, sample (j-Lmax)%S) of phongShade's lightStrikes (:836).
// In point mode the sample ray of a light IS its gate ray, so only the gate jobs exist (S = 0).
// Rays are shot from the light sample towards the hit point, exactly like the reference
// (origin = sample, direction = hit - sample, occluded iff some face has 1e-5 < t < 0.98).
// Work unit = one warp-round: job j for the 32 consecutive hit slots of one chunk; units are numbered
// j-major (unit = j * n_chunks + chunk).  The 32 lanes of a round are hits of one 8x4 pixel tile
// (hit_list order) looking at the SAME light sample: a pinhole bundle from the sample to a small
// surface patch, far more coherent than the S rays that fan out from one hit point; and at any moment
// all warps of the GPU work on the same one or two samples, which keeps the upper tree levels of that
// bundle in L1.  j is warp-uniform, the per-job set-up is one 16-byte load of the hit point (hit_p).
//
// Area mode runs in two launches.  The reference shoots no sample rays for a hit whose gate failed (SHADOW,
// src/flyscene.cpp:699-710: it returns before phongShade), and on a frame with real shadows a third of the hits are
// dark (C4: 31 %) -- as lanes of the sample units they would idle for the whole unit (ncu on C4: 22.8 of 32 lanes
// held a ray when a unit started).  So pass 1 (`pass` = 1) traces the gate jobs and appends the slots of the hits
// that go on to phongShade to lit_list (ballot + popc, one atomic per warp), and pass 2 traces the sample jobs over
// 32 consecutive entries of that list: full warps again.  Point mode (S = 0) has gate jobs only: one launch, pass 0.
// ---------------------------------------------------------------------------------------------
template <bool STATS, bool PLAIN>
__global__ void __launch_bounds__(128, 7) k_shadow(const DevScene sc, const FrameParams *__restrict__ fpp,
                                               const LevelBufs lv, const int level, const int J, const int Lmax,
                                               const int S, FrameCounts *fc, const int pass) {
  RT_STAGE_FRAME_PARAMS(fpp);
  const int lane = threadIdx.x & 31;
  TravStats st; st.box_tests = 0; st.tri_tests = 0; st.filter_checks = 0; st.filter_slow = 0; st.filter_rejects = 0;
  unsigned traced = 0;
  const unsigned n_slots = pass == 2 ? (unsigned)fc->n_lit[level] : (unsigned)fc->n_hits[level];
  const unsigned n_chunks = (n_slots + 31u) >> 5;
  const unsigned j_first = pass == 2 ? (unsigned)Lmax : 0u;                          // jobs of this pass
  const unsigned j_count = pass == 0 ? (unsigned)J : (pass == 1 ? (unsigned)Lmax : (unsigned)(J - Lmax));
  const unsigned n_units = n_chunks * j_count;  // < 2^27: the host refuses frames with n0 * J >= 2^32
  // (32-bit cursors in the two words of the 64-bit counter -- pass 2 takes the high one --: a warp overshoots
  // n_units by at most one batch)
  unsigned *cursor = reinterpret_cast<unsigned *>(&fc->work_k2[level]) + (pass == 2 ? 1 : 0);
  const unsigned batch = (unsigned)(pool_batch((unsigned long long)n_units * 32ull) >> 5);  // units per cursor update
  unsigned pool_next = 0, pool_end = 0;  // warp-local batch of units (warp-uniform)
  // job and chunk of the current unit, advanced incrementally (one division per batch, not per unit)
  unsigned j = 0, chunk = 0;
  int l = 0, s = -1;  // light and sample of job j; s < 0: gate ray

  Trav<true, STATS, PLAIN> tr;
  int stack[RT_STACK_SIZE];
  tr.idle();
  // Work donation (Trav::run_split): when the launch has only a few rounds per resident warp, its run time is set by
  // the longest rays of the last rounds (C3: 3.3 rounds per warp, SMs idle for 29 % of K2) -- idle lanes then take
  // over pending subtrees of the lanes still traversing (K2 0.340 -> 0.28 ms).  With plenty of rounds per warp the
  // bookkeeping costs more than the tail (C5 K2 2.17 -> 2.25 ms), so it is tied to the amount of work.
  const int split_min = sc.split_min >= 100 ? sc.split_min - 100 : sc.split_min;
  const bool split = !PLAIN && split_min > 0 &&
                     (sc.split_min >= 100 || n_units < 12u * gridDim.x * (blockDim.x >> 5));

  for (;;) {
    bool new_job = false;

    if (pool_next >= pool_end) {
      unsigned base = 0;
      if (lane == 0) base = atomicAdd(cursor, batch);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base >= n_units) break;
      pool_next = base;
      pool_end = base + batch < n_units ? base + batch : n_units;
      const unsigned jj = base / n_chunks;
      j = j_first + jj;
      chunk = base - jj * n_chunks;
      new_job = true;
    } else if (++chunk == n_chunks) {
      chunk = 0; ++j;
      new_job = true;
    }
    pool_next += 1u;
    const unsigned entry = chunk * 32u + (unsigned)lane;  // position in hit_list order (pass 0, 1) or in lit_list (pass 2)
    const bool in_range = entry < n_slots;
    const unsigned slot = !in_range ? 0u : (pass == 2 ? (unsigned)lv.lit_list[entry] : entry);
    if (new_job) {
      l = (int)j; s = -1;
      if (j >= (unsigned)Lmax) {
        const unsigned q = j - (unsigned)Lmax;
        if (Lmax == 1) { l = 0; s = (int)q; }
        else { l = (int)(q / (unsigned)S); s = (int)(q - (unsigned)l * (unsigned)S); }
      }
    }
    bool active = false, have = false;
    uint8_t visible = 0;
    int iw = 0;
    if (in_range) {
      const float4 hp = lv.hit_p[slot];
      const V3 hit = mk(hp);
      iw = __float_as_int(hp.w);
      V3 src;
      if (iw & 1) {
        // mirror child: its light list is the single point it inherited (the parent's hit point)
        have = l == 0;
        const int i = iw >> 1;
        const V3 lp = mk(lv.ray_d[i].w, lv.ray_l[i].x, lv.ray_l[i].y);
        src = s < 0 ? lp : area_sample(fp, lp, s);
      } else {
        have = l < fp.n_lights;
        src = s < 0 ? ld3(fp.lights + 3 * l)
                    : (fp.have_sample_table ? ld3(fp.sample_table + 3 * (l * S + s)) : area_sample(fp, ld3(fp.lights + 3 * l), s));
      }
      if (have) {
        const V3 sd = sub(hit, src);  // :920
        const V3 rdir = recip_dir(sd);
        traced++;
        if (STATS) st.box_tests += 1;  // the bounds test
        visible = 1;  // unless the traversal finds an occluder: t stays FLT_MAX >= 0.98 (:946)
        // a segment that cannot reach the BVH's bounds is unoccluded whatever the reference's root-box test (:924)
        // says, so that (bit-exact, dearer) test is only evaluated for the rays that will be traversed
        if (segment_reaches_bvh(sc, src, rdir, 0.98f)) {
          const bool tri_enabled = ref_box_intersect_quick(sc.root_min, sc.root_max, src, hit, rdir);  // :924
          if (tri_enabled || (!PLAIN && sc.n_spheres > 0)) {
            tr.init(src, sd, hit, tri_enabled, rdir);
            active = true;
          }
        }
      }
    }
    // ---- traverse: the whole warp, lanes without a ray idle (or help, run_split) inside ----
    if (split) {
      if (__any_sync(0xffffffffu, active)) {
        tr.run_split(sc, st, stack, active, split_min);
        if (active) visible = tr.fin_occ ? 0 : 1;
      }
    } else {
      while (__any_sync(0xffffffffu, active)) {
        tr.run(sc, st, stack, 0xffffffffu);
        if (active && tr.finish(sc, st)) {
          active = false;
          visible = tr.occluded ? 0 : 1;
        }
      }
    }
    if (in_range) lv.vis[(size_t)slot * (size_t)J + j] = visible;
    if (pass == 1 && j == 0u) {
      // The hit goes on to phongShade unless its gate failed.  With one light (the scene's only one, or the single
      // inherited point of a mirror child) that is this ray's verdict; with several lights every hit stays in the
      // list (its other gates are other units) and K3 applies the gate.
      const bool lit = in_range && !(((iw & 1) || fp.n_lights == 1) && visible == 0);
      const int pos = warp_append(&fc->n_lit[level], lit);
      if (lit) lv.lit_list[pos] = (int32_t)slot;
    }
  }
  // rays actually traced (work; the census is counted by K3): a few sample rays of dark hits may have
  // been traced before their gate result was visible
  if (STATS) warp_sum_add(&fc->ctr.shadow_rays_traced, traced);
  if (STATS) {
    warp_sum_add(&fc->ctr.box_tests_k2, st.box_tests);
    warp_sum_add(&fc->ctr.tri_tests_k2, st.tri_tests);
    warp_sum_add(&fc->ctr.filter_checks, st.filter_checks);
    warp_sum_add(&fc->ctr.filter_slow, st.filter_slow);
    warp_sum_add(&fc->ctr.filter_rejects, st.filter_rejects);
  }
}


// ---------------------------------------------------------------------------------------------
// shading helpers
// ---------------------------------------------------------------------------------------------
struct Material {
  V3 kd, ks;
  float ns, ni;
  int illum;
};
__device__ __forceinline__ Material load_material(const DevScene &sc, int mid) {
  const float4 a = __ldg(sc.mats + 3 * mid), b = __ldg(sc.mats + 3 * mid + 1), c = __ldg(sc.mats + 3 * mid + 2);
  Material m;
  m.kd = mk(a); m.ns = a.w; m.ks = mk(b); m.ni = b.w; m.illum = __float_as_int(c.x);
  return m;
}

// powf as the reference's libm computes it: glibc's powf is correctly rounded in all but
// astronomically rare cases; CUDA's float powf is not (up to 4 ulp), so evaluate in double and
// round once.  Integer exponents (the usual MTL "Ns 10" / "Ns 50") use binary powering: <= 2*log2(n)
// double multiplications, relative error <= ~n * 1.1e-16, i.e. closer to the true value than a
// general double pow() and an order of magnitude cheaper.  pow(0, y>0) = 0 and pow(1, y) = 1
// short-cut the common cases.
__device__ __forceinline__ float pow_ref(float x, float y) {
  if (x == 1.0f) return 1.0f;
  if (x == 0.0f && y > 0.0f) return 0.0f;
  const int n = (int)y;
  if ((float)n == y && n >= 1 && n <= 256 && x > 0.0f) {
    double b = (double)x, r = 1.0;
    for (int k = n; k != 0; k >>= 1) {
      if (k & 1) r *= b;
      b *= b;
    }
    return (float)r;
  }
  return (float)pow((double)x, (double)y);
}

// Flyscene::fresnel, src/flyscene.cpp:890-910
__device__ __forceinline__ float fresnel_ref(V3 I, V3 N, float ior) {
  float cosi = dot(I, N);
  float etai = 1.f, etat = ior;
  if (cosi > 0.f) { const float tmp = etai; etai = etat; etat = tmp; }
  const float sint = etai / etat * sqrtf(max_std(0.f, 1.f - cosi * cosi));
  if (sint >= 1.f) return 1.f;
  const float cost = sqrtf(max_std(0.f, 1.f - sint * sint));
  cosi = fabsf(cosi);
  const float Rs = ((etat * cosi) - (etai * cost)) / ((etat * cosi) + (etai * cost));
  const float Rp = ((etai * cosi) - (etat * cost)) / ((etai * cosi) + (etat * cost));
  return (Rs * Rs + Rp * Rp) / 2.f;
}

// refraction vector, src/flyscene.cpp:747-749 (float c1, double pow/sqrt for c2)
__device__ __forceinline__ V3 refract_ref(V3 d, V3 n, float ni) {
  const float c1 = fabsf(dot(d, n));
  const float inv = 1.f / ni;
  const double p1 = (double)inv * (double)inv;
  const double p2 = (double)c1 * (double)c1;
  const float c2 = (float)sqrt(1.0 - p1 * (1.0 - p2));
  return add(mul(inv, d), mul(inv * c1 - c2, n));
}

// ppmIO.hpp:145 : min(255, (int)(255*c)); negative / NaN results (printed as negative numbers by
// the reference) are stored as 0 in the packed framebuffer.
__device__ __forceinline__ unsigned char quantize(float c) {
  const float v = 255.f * c;
  int q = (v != v) ? 0 : __float2int_rz(v);
  q = q < 255 ? q : 255;
  return (unsigned char)(q < 0 ? 0 : q);
}
__device__ __forceinline__ uchar4 pack_pixel(V3 c) { return make_uchar4(quantize(c.x), quantize(c.y), quantize(c.z), 255); }

__device__ __forceinline__ V3 blend(RecType ty, V3 P, float f, V3 child) {
  V3 c;
  switch (ty) {
    case REC_MIRROR: c = add(mul(0.15f, P), mul(0.85f, child)); break;
    case REC_MIRROR_FRESNEL: c = mul(f, add(mul(0.15f, P), mul(0.85f, child))); break;
    case REC_GLASS9: c = add(mul(0.10f, P), mul(0.90f, child)); break;
    case REC_REFRACT6: c = add(mul(0.2f, P), mul(0.8f, child)); break;
    default: c = P; break;
  }
  return c;
}

// surface data at a hit: face normal, un-normalised interpolated normal, material id
// (getInterpolatedNormal, src/flyscene.cpp:864-888; for spheres the radial direction)
__device__ __forceinline__ void surface_at(const DevScene &sc, int face, V3 hit, V3 &fn, V3 &nrm_in, int &mid) {
  if (face < sc.n_faces) {
    const float4 *sp = sc.shade + (size_t)face * 7;
    const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1), s2 = __ldg(sp + 2);
    const float4 s3 = __ldg(sp + 3), s4 = __ldg(sp + 4), s5 = __ldg(sp + 5), s6 = __ldg(sp + 6);
    mid = __float_as_int(s0.w);
    fn = mk(s6);
    const V3 a = mk(s0), b = mk(s1), c = mk(s2);
    const V3 v0 = sub(b, a), v1 = sub(c, a), v2 = sub(hit, a);
    const float d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
    const float denom = d00 * d11 - d01 * d01;
    const float bv = (d11 * d20 - d01 * d21) / denom;
    const float bw = (d00 * d21 - d01 * d20) / denom;
    const float bu = 1.0f - bv - bw;
    nrm_in = add(add(mul(bu, mk(s3)), mul(bv, mk(s4))), mul(bw, mk(s5)));
  } else {
    const int si = face - sc.n_faces;
    const float4 cr = __ldg(sc.spheres + si);
    mid = __ldg(sc.sphere_mat + si);
    nrm_in = sub(hit, mk(cr));
    fn = normalized(nrm_in);
  }
}

// one visible light sample of phongShade's inner loop (src/flyscene.cpp:844-854): diffuse + specular
__device__ __forceinline__ V3 phong_sample(V3 Ikd, V3 Iks, float shininess, V3 hit, V3 spos, V3 normal, V3 eye) {
  const V3 ldir = normalized(sub(spos, hit));
  const float costheta = max_std(0.0f, dot(ldir, normal));
  const V3 diffuse = mul(costheta, Ikd);
  const V3 refl = normalized(sub(ldir, mul(2.f * dot(ldir, normal), normal)));
  const float cosphi = max_std(0.0f, dot(eye, mul(-1.f, refl)));
  const V3 specular = mul(pow_ref(cosphi, shininess), Iks);
  return add(diffuse, specular);
}

// ---------------------------------------------------------------------------------------------
// K3: shade one bounce level.  One thread per ray that hit (hit_list slot); warps stay converged
// around the queue append (ballot + popc, one atomic per warp).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 8) k_shade(const DevScene sc, const FrameParams *__restrict__ fpp,
                                              const LevelBufs lv, const LevelBufs nx, const int level, const int J,
                                              const int Lmax, const int S, FrameCounts *fc,
                                              const cudaGraphConditionalHandle next_level_cond) {
  RT_STAGE_FRAME_PARAMS(fpp);
  uchar4 *fb = level == 0 ? fp.out_rgba : nullptr;
  float *rgb_f32 = level == 0 ? fp.out_rgbf : nullptr;
  const int n_hits = fc->n_hits[level];
  const int n_round = (n_hits + 31) & ~31;
  unsigned samples_shaded = 0, shadow_asked = 0;
  for (int slot_in = blockIdx.x * blockDim.x + threadIdx.x; slot_in < n_round; slot_in += gridDim.x * blockDim.x) {
    const bool valid = slot_in < n_hits;
    const int i = valid ? lv.hit_list[slot_in] : 0;
    bool spawn = false;
    V3 colour = mk(1.f, 1.f, 1.f);  // BACKGROUND, src/flyscene.cpp:12
    RecType ty = REC_TERMINAL;
    float fres = 1.f;
    V3 child_o = mk(0, 0, 0), child_d = mk(0, 0, 0), child_lp = mk(0, 0, 0);
    int child_flags = 0;
    const int face = valid ? lv.hit_face[i] : -1;
    if (face >= 0) {
      const float4 ro = lv.ray_o[i], rd = lv.ray_d[i];
      const float2 rl2 = lv.ray_l[i];
      const RayLights rl = ray_lights(fp, ro, rd, rl2);
      const V3 o = mk(ro), d = mk(rd);
      const V3 hit = add(o, mul(lv.hit_t[i], d));
      const uint8_t *vis = lv.vis + (size_t)slot_in * J;
      bool any = false;
      for (int l = 0; l < rl.n; ++l) any = any || (vis[l] != 0);
      // ray census with the reference's semantics: L gate rays per hit, L*S sample rays only if the gate
      // passed (src/flyscene.cpp:699-710,836); in point mode (S = 0) the sample ray IS the gate ray
      shadow_asked += (unsigned)(rl.n + (any ? rl.n * S : 0));
      if (!any) {
        colour = mk(0.f, 0.f, 0.f);  // SHADOW, :699-710
      } else {
        // ---- surface data ----
        V3 fn, nrm_in;
        int mid;
        surface_at(sc, face, hit, fn, nrm_in, mid);
        const Material m = load_material(sc, mid);
        // ---- phongShade, :822-859 ----
        const V3 I = ld3(fp.light_color);
        const V3 normal = normalized(affine_point(sc.model, nrm_in));
        const V3 eye = normalized(mul(-1.f, sub(hit, o)));
        const V3 Ikd = cmul(I, m.kd), Iks = cmul(I, m.ks);
        V3 P = mk(0.f, 0.f, 0.f);
        const int ns = fp.point_light ? 1 : S;
        for (int l = 0; l < rl.n; ++l) {
          float sum = 0.f;
          V3 acc = mk(0.f, 0.f, 0.f);
          const V3 lpos = light_pos(fp, rl, l);
          for (int s = 0; s < ns; ++s) {
            const bool v = fp.point_light ? (vis[l] != 0) : (vis[Lmax + l * S + s] != 0);
            if (!v) continue;
            sum += 1.f;
            const V3 spos = fp.point_light ? lpos : light_sample(fp, rl, l, s, S);
            acc = add(acc, phong_sample(Ikd, Iks, m.ns, hit, spos, normal, eye));
            samples_shaded++;
          }
          const float fa = sum / (float)ns, fb2 = 1.3f / (float)ns;
          P = add(P, mul(fb2, mul(fa, acc)));
        }
        // ---- material switch, :712-760 ----
        int imodel = m.illum;
        if (fp.max_depth >= 0 && level >= fp.max_depth) imodel = 2;
        colour = P;
        if (imodel == 9) {
          ty = REC_GLASS9; child_d = d;
          child_flags = rl.single ? 1 : 0; child_lp = rl.lp;
        } else if (imodel == 6) {
          ty = REC_REFRACT6; child_d = refract_ref(d, fn, m.ni);
          child_flags = rl.single ? 1 : 0; child_lp = rl.lp;
        } else if (imodel > 2 && imodel < 6) {
          child_d = sub(d, mul(2.f * dot(d, fn), fn));  // :734
          ty = REC_MIRROR;
          if (imodel == 5) { ty = REC_MIRROR_FRESNEL; fres = fresnel_ref(child_d, fn, m.ni); }
          child_flags = 1; child_lp = hit;  // reflectedLights = { hitPoint }, :735-736
        }
        // illum 7: both child traces are multiplied by (1 - fresnelIndex) = 0 -> Phong (:726,751,758)
        if (ty != REC_TERMINAL) {
          if (level >= fp.guard_depth) {
            // recursion guard of unbounded mode: the child returns BACKGROUND untraced
            colour = blend(ty, P, fres, mk(1.f, 1.f, 1.f));
            ty = REC_TERMINAL;
          } else {
            spawn = true; child_o = hit;
          }
        }
      }
    }
    const int slot = warp_append(&fc->n_rays[level + 1], spawn);
    if (!valid) continue;
    if (spawn) {
      nx.ray_o[slot] = make_float4(child_o.x, child_o.y, child_o.z, __int_as_float(child_flags));
      nx.ray_d[slot] = make_float4(child_d.x, child_d.y, child_d.z, child_lp.x);
      nx.ray_l[slot] = make_float2(child_lp.y, child_lp.z);
    }
    lv.rec[i] = make_float4(colour.x, colour.y, colour.z, fres);
    lv.child[i] = spawn ? slot : -1;
    lv.type[i] = (uint8_t)ty;
    if (level == 0 && ty == REC_TERMINAL) {
      if (fb) fb[fb_index(fp, i)] = pack_pixel(colour);
      if (rgb_f32) { rgb_f32[3 * (size_t)i] = colour.x; rgb_f32[3 * (size_t)i + 1] = colour.y; rgb_f32[3 * (size_t)i + 2] = colour.z; }
    }
  }
  // counters: warp shuffle -> shared -> one global atomic per CTA and counter
  __shared__ unsigned s_count[2];
  if (threadIdx.x < 2) s_count[threadIdx.x] = 0u;
  __syncthreads();
  for (int off = 16; off > 0; off >>= 1) {
    samples_shaded += __shfl_down_sync(0xffffffffu, samples_shaded, off);
    shadow_asked += __shfl_down_sync(0xffffffffu, shadow_asked, off);
  }
  if ((threadIdx.x & 31) == 0) {
    if (samples_shaded) atomicAdd(&s_count[0], samples_shaded);
    if (shadow_asked) atomicAdd(&s_count[1], shadow_asked);
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_count[0]) atomicAdd(&fc->ctr.shade_samples, (unsigned long long)s_count[0]);
  if (threadIdx.x == 1 && s_count[1]) atomicAdd(&fc->ctr.shadow_rays, (unsigned long long)s_count[1]);
  // CUDA-graph replay: the last CTA to finish tells the graph whether the next bounce level has any
  // rays; if not, the conditional node that holds that level's kernels (and everything deeper) is skipped
  if (next_level_cond != 0) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned prev = atomicAdd(&fc->k3_done[level], 1u);
      if (prev == gridDim.x - 1) {
        __threadfence();
        const int spawned = atomicAdd(&fc->n_rays[level + 1], 0);
        cudaGraphSetConditional(next_level_cond, spawned > 0 ? 1u : 0u);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K3b: fold level k from level k+1 (deepest first); level 0 writes the framebuffer.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fold(const FrameParams *__restrict__ fpp, const LevelBufs lv, const LevelBufs nx,
                                             const int level, const int n0, const FrameCounts *fc) {
  RT_STAGE_FRAME_PARAMS(fpp);
  uchar4 *fb = level == 0 ? fp.out_rgba : nullptr;
  float *rgb_f32 = level == 0 ? fp.out_rgbf : nullptr;
  const int n = n0 >= 0 ? n0 : fc->n_rays[level];
  if (fc->n_rays[level + 1] == 0) return;  // nothing was spawned below this level
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const RecType ty = (RecType)lv.type[i];
    if (ty == REC_TERMINAL) continue;
    const float4 r = lv.rec[i];
    const float4 c = nx.rec[lv.child[i]];
    const V3 col = blend(ty, mk(r), r.w, mk(c));
    lv.rec[i] = make_float4(col.x, col.y, col.z, 1.f);
    if (level == 0) {
      if (fb) fb[fb_index(fp, i)] = pack_pixel(col);
      if (rgb_f32) { rgb_f32[3 * (size_t)i] = col.x; rgb_f32[3 * (size_t)i + 1] = col.y; rgb_f32[3 * (size_t)i + 2] = col.z; }
    }
  }
}

// writes the frame parameters (passed by value at launch) into their device buffer
__global__ void k_set_frame(const FrameParams fp, FrameParams *dst) {
  const uint32_t *src = reinterpret_cast<const uint32_t *>(&fp);
  uint32_t *d = reinterpret_cast<uint32_t *>(dst);
  for (int w = threadIdx.x; w < (int)(sizeof(FrameParams) / 4); w += blockDim.x) d[w] = src[w];
}

// ---------------------------------------------------------------------------------------------
// small batched entry points
// ---------------------------------------------------------------------------------------------
__global__ void k_box_intersect(const DevScene sc, const int64_t n, const float *o, const float *dst, uint8_t *out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = ref_box_intersect(sc.root_min, sc.root_max, ld3(o + 3 * i), ld3(dst + 3 * i)) ? 1 : 0;
}

__global__ void k_screen_to_world(const FrameParams fp, const int64_t n, const float *pix, float *out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const V3 w = screen_to_world(fp, pix[2 * i], pix[2 * i + 1]);
    out[3 * i] = w.x; out[3 * i + 1] = w.y; out[3 * i + 2] = w.z;
  }
}

// lightStrikes for explicit hit points: job = point * L + light
__global__ void __launch_bounds__(128) k_light_strikes(const DevScene sc, const FrameParams fp, const int64_t n,
                                                      const float *hits, uint8_t *out) {
  const int L = fp.n_lights;
  TravStats st; st.box_tests = 0; st.tri_tests = 0; st.filter_checks = 0; st.filter_slow = 0; st.filter_rejects = 0;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n * L; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = g / L;
    const int l = (int)(g - i * L);
    const V3 hit = ld3(hits + 3 * i), src = ld3(fp.lights + 3 * l);
    const bool tri_enabled = ref_box_intersect(sc.root_min, sc.root_max, src, hit);
    bool occ = false;
    if (tri_enabled || sc.n_spheres > 0) {
      float bt = RT_NO_HIT_T; int bi = -1;
      occ = traverse<true, false>(sc, src, sub(hit, src), hit, tri_enabled, bt, bi, st);
    }
    out[g] = occ ? 0 : 1;
  }
}

// ---------------------------------------------------------------------------------------------
// Batched per-function entry points (one thread per query; not on the frame path)
// ---------------------------------------------------------------------------------------------
// BoundingBox::boxIntersect against an arbitrary box
__global__ void k_box_intersect_box(const float3 mn, const float3 mx, const int64_t n, const float *o, const float *dst,
                                    uint8_t *out) {
  const float bmn[3] = {mn.x, mn.y, mn.z}, bmx[3] = {mx.x, mx.y, mx.z};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = ref_box_intersect(bmn, bmx, ld3(o + 3 * i), ld3(dst + 3 * i)) ? 1 : 0;
}

// Flyscene::rayTriangleIntersection (src/flyscene.cpp:787-819) for (ray, face) pairs; -72 = miss
__global__ void k_ray_triangle(const DevScene sc, const int64_t n, const float *o, const float *d, const int32_t *face,
                               float *t_out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int f = face[i];
    float res = -72.f;
    if (f >= 0 && f < sc.n_faces) {
      const float4 *sp = sc.shade + (size_t)f * 7;
      const V3 a = mk(__ldg(sp)), b = mk(__ldg(sp + 1)), c = mk(__ldg(sp + 2)), nrm = mk(__ldg(sp + 6));
      const V3 ro = ld3(o + 3 * i), rd = ld3(d + 3 * i);
      const float den = dot(rd, nrm);
      if (den != 0.f) {
        const float t = (dot(nrm, a) - dot(ro, nrm)) / den;
        const V3 P = add(ro, mul(t, rd));
        const V3 v0 = sub(c, a), v1 = sub(b, a), v2 = sub(P, a);
        const float d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d02 = dot(v0, v2), d12 = dot(v1, v2);
        const float inv = 1.f / (d00 * d11 - d01 * d01);
        const float u = (d11 * d02 - d01 * d12) * inv, v = (d00 * d12 - d01 * d02) * inv;
        if ((u >= 0.f) && (v >= 0.f) && (u + v < 1.f)) res = t;
      }
    }
    t_out[i] = res;
  }
}

// BoxTree::intersect (src/boxTree.cpp:150-173): flag[f] = 1 iff the reference octree offers face f
// for the query (origin, dest).  One thread per face.
__global__ void k_octree_candidates(const DevScene sc, const float3 o3, const float3 d3, uint8_t *flag) {
  const V3 o = mk(o3.x, o3.y, o3.z), dest = mk(d3.x, d3.y, d3.z);
  const bool root = ref_box_intersect(sc.root_min, sc.root_max, o, dest);
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < sc.n_faces; f += gridDim.x * blockDim.x)
    flag[f] = (root && (sc.oct_box == nullptr || ref_candidate(sc, f, o, dest))) ? 1 : 0;
}

// Flyscene::phongShade (src/flyscene.cpp:822-859) for explicit (origin, hit point, face) triples with
// the frame's light list; shadow rays for every sample are traced inline.
__global__ void __launch_bounds__(128) k_phong_shade(const DevScene sc, const FrameParams fp, const int64_t n,
                                                    const float *origins, const float *hits, const int32_t *faces,
                                                    float *rgb_out) {
  TravStats st; st.box_tests = 0; st.tri_tests = 0; st.filter_checks = 0; st.filter_slow = 0; st.filter_rejects = 0;
  const int S = fp.point_light ? 1 : fp.usteps * fp.vsteps;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int face = faces[i];
    V3 P = mk(0.f, 0.f, 0.f);
    if (face >= 0 && face < sc.n_faces + sc.n_spheres) {
      const V3 o = ld3(origins + 3 * i), hit = ld3(hits + 3 * i);
      V3 fn, nrm_in;
      int mid;
      surface_at(sc, face, hit, fn, nrm_in, mid);
      const Material m = load_material(sc, mid);
      const V3 I = ld3(fp.light_color);
      const V3 normal = normalized(affine_point(sc.model, nrm_in));
      const V3 eye = normalized(mul(-1.f, sub(hit, o)));
      const V3 Ikd = cmul(I, m.kd), Iks = cmul(I, m.ks);
      for (int l = 0; l < fp.n_lights; ++l) {
        float sum = 0.f;
        V3 acc = mk(0.f, 0.f, 0.f);
        const V3 lpos = ld3(fp.lights + 3 * l);
        for (int s = 0; s < S; ++s) {
          const V3 spos = fp.point_light ? lpos : area_sample(fp, lpos, s);
          const bool tri_enabled = ref_box_intersect(sc.root_min, sc.root_max, spos, hit);
          bool occ = false;
          if (tri_enabled || sc.n_spheres > 0) {
            float bt = RT_NO_HIT_T; int bi = -1;
            occ = traverse<true, false>(sc, spos, sub(hit, spos), hit, tri_enabled, bt, bi, st);
          }
          if (occ) continue;
          sum += 1.f;
          acc = add(acc, phong_sample(Ikd, Iks, m.ns, hit, spos, normal, eye));
        }
        const float fa = sum / (float)S, fb2 = 1.3f / (float)S;
        P = add(P, mul(fb2, mul(fa, acc)));
      }
    }
    rgb_out[3 * i] = P.x; rgb_out[3 * i + 1] = P.y; rgb_out[3 * i + 2] = P.z;
  }
}

// self-test of div3 (rt_device.cuh) against IEEE division: trial k draws a divisor and three numerators from
// a counter hash -- uniformly random bit patterns, exponents confined to the given range around 1 for the
// "typical" trials, unrestricted patterns (incl. zero, denormal, inf, NaN) otherwise
__device__ __forceinline__ unsigned selftest_hash(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__global__ void k_selftest_div3(const unsigned long long n_trials, const unsigned seed, const int exp_range,
                                unsigned long long *mismatches) {
  unsigned long long bad = 0;
  for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; k < n_trials;
       k += (unsigned long long)gridDim.x * blockDim.x) {
    float v[4];
    for (int c = 0; c < 4; ++c) {
      unsigned h = selftest_hash(seed ^ selftest_hash((unsigned)(k >> 32) * 0x9E3779B9U + (unsigned)k * 4u + (unsigned)c));
      if (exp_range > 0) {
        const unsigned ex = 127u - (unsigned)exp_range + (selftest_hash(h) % (2u * (unsigned)exp_range + 1u));
        h = (h & 0x807fffffu) | (ex << 23);
        if ((k & 15ull) == (unsigned long long)c) h &= 0x80000000u;  // sprinkle signed zeros over the numerators
      }
      v[c] = __int_as_float((int)h);
    }
    if (exp_range > 0 && v[0] == 0.f) v[0] = 1.5f;
    const V3 q = div3(mk(v[1], v[2], v[3]), v[0]);
    const float ex0 = v[1] / v[0], ex1 = v[2] / v[0], ex2 = v[3] / v[0];
    // NaN results must be NaN on both sides; everything else bit-identical (signed zeros included)
    const bool ok0 = (ex0 != ex0) ? (q.x != q.x) : (__float_as_int(q.x) == __float_as_int(ex0));
    const bool ok1 = (ex1 != ex1) ? (q.y != q.y) : (__float_as_int(q.y) == __float_as_int(ex1));
    const bool ok2 = (ex2 != ex2) ? (q.z != q.z) : (__float_as_int(q.z) == __float_as_int(ex2));
    if (!(ok0 && ok1 && ok2)) ++bad;
  }
  if (bad) atomicAdd(mismatches, bad);
}

}  // namespace rtd
