// rt_frame.cuh -- the whole frame as ONE persistent kernel (sm_100a).
//
//   k_frame   raytraceScene's pixel loop (src/flyscene.cpp:573-625) with traceRay's complete recursion
//             (:651-771) inside: nearest hit -> lightStrikes gate + area-light samples -> phongShade ->
//             material switch -> child ray, level after level, until every ray chain of the frame has ended
//             and its pixel is written as packed uchar4.
// (citations relative to /root/reference).  Compiled with -fmad=false, see rt_device.cuh.
//
// Why one kernel.  The wavefront pipeline of rt_kernels.cuh (K1 -> K2 -> K3 per bounce level, K3b folds) pays
// for its generality on every frame: ten dependent launches, a hit list, a 16-byte hit record and a
// visibility byte per shadow ray, a 37-byte record per ray and level that a fold kernel reads again.  On the
// bundled scenes that bookkeeping is most of the frame (profiles/r01_c2_ncu_final.txt: 363 warp-instructions
// per 32 shadow rays that test two boxes each).  Here a warp owns 32 rays (an 8x4 pixel tile, or 32 queued
// rays of a deeper level) from generation to pixel:
//   * the nearest-hit result, the hit point and the visibility bits of the <= 64 shadow jobs of a hit stay in
//     registers; nothing is queued between "trace", "shadow" and "shade";
//   * shadow job j (gate ray of light j, or area-light sample s of light l) is traced for all hit lanes of
//     the warp at once: j, the light and the sample position are warp-uniform, the lanes are neighbouring
//     surface points -- the same pinhole bundle as a work unit of k_shadow; the gate result is known to the
//     warp before any sample ray is shot, so sample rays of hits whose lights are all occluded are skipped
//     exactly (the reference returns SHADOW before phongShade, :699-710) instead of through a racy flag;
//   * a ray that spawns a child (mirror / glass / refraction) leaves a 24-byte chain record (Phong term,
//     blend type, Fresnel factor, parent record); the ray whose colour is final -- the LAST of its chain --
//     walks the records back to the pixel with the reference's own expressions (0.15f*P + 0.85f*child, ...)
//     and stores the uchar4.  The reference's recursion is a chain (one child per ray), so exactly one lane
//     ends every chain: no fold kernels, no synchronisation;
//   * child rays continue IN THE WARP when at least `cont_min` lanes spawned one (a mirror wall: the whole
//     tile goes on), otherwise they are compacted (ballot + popc, one atomic per warp) into the queue of
//     the next level, which all warps drain together after a grid-wide barrier -- sparse reflections
//     (a few mirror spheres) are traced with full warps again.
// The frame is: one memset of the counters + one cooperative launch.  No CUDA graph, no per-level launches,
// no host read-back.
#pragma once

#include "rt_kernels.cuh"

namespace rtd {

#ifndef RT_SHADE_UNROLL
#define RT_SHADE_UNROLL 2
#endif
// Resident CTAs per SM the frame kernel is compiled for (register budget 65536 / (128 * MINB)).  The PLAIN
// instantiation (the bundled scenes) runs best with 7 CTAs = 72 registers: its stalls are fixed-latency dependencies
// (`wait`), which a seventh warp per scheduler hides -- C2 0.272 -> 0.254 ms; 8 CTAs (64 registers) spill too much
// (0.257 ms), 5 CTAs 0.293 ms (A/B of builds on one box, tools/ab_gpu.sh).
#ifndef RT_FRAME_MINB
#define RT_FRAME_MINB 6
#endif
#ifndef RT_FRAME_MINB_PLAIN
#define RT_FRAME_MINB_PLAIN 7
#endif
#define RT_FUSED_MAX_LEVELS 9   // levels 0..8: max_depth <= 8 (deeper caps take the wavefront path)
#define RT_FUSED_MAX_JOBS 64    // shadow jobs per hit held as a bit mask

struct FusedBufs {
  // rays deferred to level k live in slab k (capacity n_cap each): 3 x float4 per ray
  float4 *q_o;   // (o.xyz, bits(flags))       flags bit0: light list = the single inherited point lp
  float4 *q_d;   // (d.xyz, lp.x)
  float4 *q_x;   // (lp.y, lp.z, bits(parent record), bits(pixel))
  // chain records of the rays of level k that spawned a child live in slab k
  float4 *rec_a;  // (Phong term P.xyz, fresnel factor)
  int2 *rec_b;    // (RecType, parent record id or -1)
  int32_t n_cap;  // rays per slab (= rays of level 0)
};

struct FusedCounts {
  unsigned int cursor[RT_FUSED_MAX_LEVELS];        // persistent work cursor of each phase
  unsigned int remaining[RT_FUSED_MAX_LEVELS + 1]; // rays queued at levels >= k when the barrier before phase k opened
  int32_t n_queue[RT_FUSED_MAX_LEVELS + 1];        // rays deferred to level k
  int32_t n_rec[RT_FUSED_MAX_LEVELS + 1];          // chain records written by level k
  int32_t n_spawn[RT_FUSED_MAX_LEVELS + 1];        // rays that exist at level k (k >= 1), in-warp or queued
  unsigned int bar_count, bar_gen;                 // grid barrier
  int32_t overflow;                                // unbounded depth: a ray wanted to go deeper than the slabs allow
  int32_t pad;
  unsigned long long phase_clk[4];                 // stats builds: warp-cycles spent in trace / shadow / shade / rest
  Counters ctr;
};

// Grid-wide barrier before phase `phase` (>= 1).  The kernel is launched cooperatively (all CTAs co-resident);
// thread 0 of every CTA arrives with a release fence.  The last one to arrive -- at that moment no warp of the
// grid is working, so every queue is quiescent -- publishes how many rays are still queued at levels >= phase
// and then opens the next generation.  Every CTA bases "is the frame finished?" on that one published number:
// reading the live counters after the barrier would not do, because a CTA that is already working on this
// phase may be appending to deeper queues.
__device__ __forceinline__ void grid_barrier(FusedCounts *fc, const int phase, const int depth_cap) {
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile unsigned int *gen = &fc->bar_gen;
    const unsigned int g = *gen;
    __threadfence();
    if (atomicAdd(&fc->bar_count, 1u) == gridDim.x - 1) {
      unsigned int rem = 0;
      for (int k = phase; k <= depth_cap; ++k) rem += (unsigned int)*reinterpret_cast<volatile int32_t *>(&fc->n_queue[k]);
      fc->remaining[phase] = rem;
      fc->bar_count = 0u;
      __threadfence();
      atomicAdd(&fc->bar_gen, 1u);
    } else {
      while (*gen == g) __nanosleep(64);
    }
    __threadfence();
  }
  __syncthreads();
}

// The ray whose colour is final walks its chain of records back to the pixel: the unwinding of traceRay's
// recursion (src/flyscene.cpp:718,738-743,754) in the reference's own evaluation order.
__device__ __forceinline__ V3 fold_chain(const FusedBufs &fb, int parent, V3 c) {
  while (parent >= 0) {
    const float4 a = fb.rec_a[parent];
    const int2 b = fb.rec_b[parent];
    c = blend((RecType)b.x, mk(a), a.w, c);
    parent = b.y;
  }
  return c;
}

template <bool STATS, bool PLAIN>
__global__ void __launch_bounds__(128, PLAIN ? RT_FRAME_MINB_PLAIN : RT_FRAME_MINB) k_frame(const DevScene sc, const FrameParams fp, const FusedBufs fb,
                                                  FusedCounts *fc, const int n0, const int explicit0, const int J,
                                                  const int Lmax, const int S, const int depth_cap, const int cont_min) {
  const int lane = threadIdx.x & 31;
  TravStats st; st.box_tests = 0; st.tri_tests = 0; st.filter_checks = 0; st.filter_slow = 0; st.filter_rejects = 0;
  unsigned samples_shaded = 0, shadow_asked = 0, traced = 0, box_k2 = 0, tri_k2 = 0;
  long long clk_trace = 0, clk_shadow = 0, clk_shade = 0, clk_all = 0;
  const long long clk_start = STATS ? clock64() : 0;
  int stack[RT_STACK_SIZE];
  const int twl = fp.tile_w_log2, tile_w = 1 << twl, tile_h = 32 >> twl;
  const int tiles_x = (fp.width + tile_w - 1) >> twl;
  const size_t cap = (size_t)fb.n_cap;

  for (int phase = 0; phase <= depth_cap; ++phase) {
    // ---- work of this phase: pixel tiles (phase 0 of a camera frame) or the rays deferred to this level ----
    const bool primary = phase == 0 && !explicit0;
    unsigned int n_items;
    if (primary) n_items = (unsigned int)tiles_x * (unsigned int)((fp.local_rows + tile_h - 1) / tile_h) * 32u;
    else if (phase == 0) n_items = (unsigned int)n0;
    else {
      grid_barrier(fc, phase, depth_cap);
      if (*reinterpret_cast<volatile unsigned int *>(&fc->remaining[phase]) == 0u) break;  // the frame is complete
      n_items = (unsigned int)*reinterpret_cast<volatile int32_t *>(&fc->n_queue[phase]);  // final since the barrier
      if (n_items == 0u) continue;
    }
    unsigned int *cursor = &fc->cursor[phase];
    const unsigned int batch = (unsigned int)pool_batch(n_items);
    unsigned int pool_next = 0, pool_end = 0;  // warp-local batch of work items (warp-uniform)
    const float4 *qo = fb.q_o + (size_t)phase * cap, *qd = fb.q_d + (size_t)phase * cap, *qx = fb.q_x + (size_t)phase * cap;

    for (;;) {
      if (pool_next >= pool_end) {
        // (32-bit cursor: n_items < 2^31 and every warp overshoots it by at most one batch)
        unsigned int base = 0;
        if (lane == 0) base = atomicAdd(cursor, batch);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_items) break;
        pool_next = base;
        pool_end = base + batch < n_items ? base + batch : n_items;
      }
      const unsigned int item = pool_next + (unsigned int)lane;
      pool_next += 32u;

      // ---- the lane's ray ----
      V3 o = mk(0.f, 0.f, 0.f), d = mk(0.f, 0.f, 0.f), lp = mk(0.f, 0.f, 0.f), screen = mk(0.f, 0.f, 0.f);
      bool valid = false, single = false;
      int pix = 0, parent = -1;
      if (item < pool_end) {
        if (primary) {
          const int tile = (int)(item >> 5), in_tile = (int)(item & 31);
          int tx, ty;
          tile_xy(fp, tile, tiles_x, (fp.local_rows + tile_h - 1) / tile_h, tx, ty);
          const int px = (tx << twl) + (in_tile & (tile_w - 1)), py = ty * tile_h + (in_tile >> twl);
          valid = px < fp.width && py < fp.local_rows;
          pix = py * fp.width + px;
          o = ld3(fp.eye);
          if (tile_outside(fp, tx, ty)) {  // (warp-uniform) the whole tile is BACKGROUND, src/flyscene.cpp:658-665
            if (valid) {
              if (fp.out_face) fp.out_face[pix] = -1;
              if (fp.out_t) fp.out_t[pix] = RT_NO_HIT_T;
              if (fp.out_rgba) fp.out_rgba[fb_index(fp, pix)] = make_uchar4(255, 255, 255, 255);
              if (fp.out_rgbf) { fp.out_rgbf[3 * (size_t)pix] = 1.f; fp.out_rgbf[3 * (size_t)pix + 1] = 1.f; fp.out_rgbf[3 * (size_t)pix + 2] = 1.f; }
            }
            continue;
          }
          if (valid) {
            screen = screen_to_world(fp, (float)px, (float)global_row(fp, py));
            d = sub(screen, o);  // :619, not normalised
          }
        } else {
          valid = true;
          const float4 a = qo[item], b = qd[item], c = qx[item];
          o = mk(a); d = mk(b);
          single = (__float_as_int(a.w) & 1) != 0;
          lp = mk(b.w, c.x, c.y);
          parent = __float_as_int(c.z);
          pix = __float_as_int(c.w);
        }
      }

      // ---- the rays of this warp, level after level ----
      for (int level = phase;; ++level) {
        const bool first = primary && level == 0;
        long long c0 = STATS ? clock64() : 0;
        // ================= nearest hit (traceRay :655-695) =================
        float hit_t = RT_NO_HIT_T;
        int hit_face = -1;
        {
          Trav<false, STATS, PLAIN> tr;
          tr.idle();
          bool active = false;
          if (valid) {
            const V3 rdir = recip_dir(d);
            bool tri_enabled = true;
            // raytraceScene's root-box pre-cull on (origin, screen), :576 (skipped with analytic spheres, like rt_oracle.c)
            if (first) tri_enabled = ref_box_intersect_quick(sc.root_min, sc.root_max, o, screen, rdir) || (!PLAIN && sc.n_spheres > 0);
            // traceRay's own root test on (origin, origin + direction), :655
            const V3 dest = add(o, d);
            tri_enabled = tri_enabled && ref_box_intersect_quick(sc.root_min, sc.root_max, o, dest, recip_dir(sub(dest, o)));
            if (tri_enabled || (!PLAIN && sc.n_spheres > 0)) {
              tr.init(o, d, dest, tri_enabled, rdir);
              active = true;
            }
          }
          while (__any_sync(0xffffffffu, active)) {
            tr.run(sc, st, stack, 0xffffffffu);
            if (active && tr.finish(sc, st)) {
              active = false;
              hit_t = tr.best_t;
              hit_face = tr.best_id;
            }
          }
        }
        const bool hit = valid && hit_face >= 0;
        if (level == 0 && valid) {
          if (fp.out_face) fp.out_face[pix] = hit_face;
          if (fp.out_t) fp.out_t[pix] = hit_t;
        }
        if (STATS) { const long long c1 = clock64(); clk_trace += c1 - c0; c0 = c1; }

        // ================= shadow jobs (lightStrikes :912-954 for the gate and every sample) =================
        const V3 P = add(o, mul(hit_t, d));  // :695
        unsigned long long vis = 0ull;
        bool any = false;
        const int n_l = single ? 1 : fp.n_lights;
        if (__any_sync(0xffffffffu, hit)) {
          unsigned box0 = 0, tri0 = 0;
          if (STATS) { box0 = st.box_tests; tri0 = st.tri_tests; }
          for (int j = 0; j < J; ++j) {
            // light and sample of job j (warp-uniform); s < 0: gate ray
            int l = j, s = -1;
            if (j >= Lmax) {
              const int q = j - Lmax;
              l = q / S; s = q - l * S;
              // the gates are all known: hits with no visible light are SHADOW, their samples are never shot
              if (q == 0) any = (vis & ((1ull << Lmax) - 1ull)) != 0ull;
            }
            const bool want = hit && l < n_l && (s < 0 || any);
            if (!__any_sync(0xffffffffu, want)) continue;
            Trav<true, STATS, PLAIN> tr;
            tr.idle();
            bool active = false;
            if (want) {
              V3 src;
              if (single) src = s < 0 ? lp : area_sample(fp, lp, s);
              else src = s < 0 ? ld3(fp.lights + 3 * l)
                               : (fp.have_sample_table ? ld3(fp.sample_table + 3 * (l * S + s)) : area_sample(fp, ld3(fp.lights + 3 * l), s));
              const V3 sd = sub(P, src);  // :920
              const V3 rdir = recip_dir(sd);
              traced++;
              if (STATS) st.box_tests += 1;  // the bounds test
              vis |= 1ull << j;  // visible unless the traversal finds an occluder: t stays FLT_MAX >= 0.98 (:946)
              // a segment that cannot reach the BVH's bounds is unoccluded whatever the reference's root-box test
              // (:924) says: that bit-exact, dearer test is only evaluated for the rays that will be traversed
              if (segment_reaches_bvh(sc, src, rdir, 0.98f)) {
                const bool tri_enabled = ref_box_intersect_quick(sc.root_min, sc.root_max, src, P, rdir);  // :924
                if (tri_enabled || (!PLAIN && sc.n_spheres > 0)) {
                  tr.init(src, sd, P, tri_enabled, rdir);
                  active = true;
                }
              }
            }
            while (__any_sync(0xffffffffu, active)) {
              tr.run(sc, st, stack, 0xffffffffu);
              if (active && tr.finish(sc, st)) {
                active = false;
                if (tr.occluded) vis &= ~(1ull << j);
              }
            }
          }
          if (S == 0) any = (vis & ((1ull << Lmax) - 1ull)) != 0ull;  // point mode: the gate ray is the sample ray
          if (STATS) { box_k2 += st.box_tests - box0; tri_k2 += st.tri_tests - tri0; st.box_tests = box0; st.tri_tests = tri0; }
        }
        if (STATS) { const long long c1 = clock64(); clk_shadow += c1 - c0; c0 = c1; }

        // ================= shading and the material switch (:697-760) =================
        bool spawn = false;
        V3 colour = mk(1.f, 1.f, 1.f);  // BACKGROUND, :12 (misses: :658-665 / :684-691)
        RecType ty = REC_TERMINAL;
        float fres = 1.f;
        V3 child_d = mk(0.f, 0.f, 0.f), child_lp = lp;
        bool child_single = single;
        if (hit) {
          // ray census with the reference's semantics: L gate rays per hit, L*S sample rays only if a gate passed
          shadow_asked += (unsigned)(n_l + (any ? n_l * S : 0));
          if (!any) {
            colour = mk(0.f, 0.f, 0.f);  // SHADOW, :699-710
          } else {
            V3 fn, nrm_in;
            int mid;
            surface_at(sc, hit_face, P, fn, nrm_in, mid);
            const Material m = load_material(sc, mid);
            // ---- phongShade, :822-859 ----
            const V3 I = ld3(fp.light_color);
            const V3 normal = normalized(affine_point(sc.model, nrm_in));
            const V3 eye = normalized(mul(-1.f, sub(P, o)));
            const V3 Ikd = cmul(I, m.kd), Iks = cmul(I, m.ks);
            V3 Ph = mk(0.f, 0.f, 0.f);
            const int ns = fp.point_light ? 1 : S;
            for (int l = 0; l < n_l; ++l) {
              float sum = 0.f;
              V3 acc = mk(0.f, 0.f, 0.f);
              const V3 lpos = single ? lp : ld3(fp.lights + 3 * l);
#if RT_SHADE_UNROLL > 1
              // RT_SHADE_UNROLL samples per iteration: their (independent) normalisation / pow chains overlap in
              // the pipeline; the sums are taken in sample order, so the result is that of the one-by-one loop
              // (C1 0.119 -> 0.115 ms, C2 +-0; 4 at a time: C2 +3 %)
              int s = 0;
              for (; s + RT_SHADE_UNROLL <= ns; s += RT_SHADE_UNROLL) {
                const int bit = Lmax + l * S + s;  // (ns > 1 only in area mode)
                const unsigned vbits = (unsigned)(vis >> bit) & ((1u << RT_SHADE_UNROLL) - 1u);
                if (vbits == 0u) continue;
                V3 ph[RT_SHADE_UNROLL];
#pragma unroll
                for (int u = 0; u < RT_SHADE_UNROLL; ++u) {
                  const V3 sp = (!single && fp.have_sample_table) ? ld3(fp.sample_table + 3 * (l * S + s + u)) : area_sample(fp, lpos, s + u);
                  ph[u] = phong_sample(Ikd, Iks, m.ns, P, sp, normal, eye);
                }
#pragma unroll
                for (int u = 0; u < RT_SHADE_UNROLL; ++u)
                  if ((vbits >> u) & 1u) { sum += 1.f; acc = add(acc, ph[u]); samples_shaded++; }
              }
              for (; s < ns; ++s) {
#else
              for (int s = 0; s < ns; ++s) {
#endif
                const int bit = fp.point_light ? l : Lmax + l * S + s;
                if (!((vis >> bit) & 1ull)) continue;
                sum += 1.f;
                V3 spos = lpos;
                if (!fp.point_light)
                  spos = (!single && fp.have_sample_table) ? ld3(fp.sample_table + 3 * (l * S + s)) : area_sample(fp, lpos, s);
                acc = add(acc, phong_sample(Ikd, Iks, m.ns, P, spos, normal, eye));
                samples_shaded++;
              }
              const float fa = sum / (float)ns, fb2 = 1.3f / (float)ns;
              Ph = add(Ph, mul(fb2, mul(fa, acc)));
            }
            // ---- material switch, :712-760 ----
            int imodel = m.illum;
            if (fp.max_depth >= 0 && level >= fp.max_depth) imodel = 2;
            colour = Ph;
            if (imodel == 9) {
              ty = REC_GLASS9; child_d = d;
            } else if (imodel == 6) {
              ty = REC_REFRACT6; child_d = refract_ref(d, fn, m.ni);
            } else if (imodel > 2 && imodel < 6) {
              child_d = sub(d, mul(2.f * dot(d, fn), fn));  // :734
              ty = REC_MIRROR;
              if (imodel == 5) { ty = REC_MIRROR_FRESNEL; fres = fresnel_ref(child_d, fn, m.ni); }
              child_single = true; child_lp = P;  // reflectedLights = { hitPoint }, :735-736
            }
            // illum 7: both child traces are multiplied by (1 - fresnelIndex) = 0 -> Phong (:726,751,758)
            if (ty != REC_TERMINAL) {
              if (level >= depth_cap) {
                // unbounded depth only (a bounded cap shades as Phong above): the slabs end here; the frame is
                // flagged and the caller renders it again through the wavefront path
                atomicExch(&fc->overflow, 1);
                ty = REC_TERMINAL;
              } else {
                spawn = true;
              }
            }
          }
        }
        if (STATS) { const long long c1 = clock64(); clk_shade += c1 - c0; c0 = c1; }

        // ================= rays that end here: fold the chain, write the pixel =================
        if (valid && !spawn) {
          const V3 c = fold_chain(fb, parent, colour);
          if (fp.out_rgba) fp.out_rgba[fb_index(fp, pix)] = pack_pixel(c);
          if (fp.out_rgbf) { fp.out_rgbf[3 * (size_t)pix] = c.x; fp.out_rgbf[3 * (size_t)pix + 1] = c.y; fp.out_rgbf[3 * (size_t)pix + 2] = c.z; }
        }

        // ================= rays that go on: chain record, then in-warp or queued =================
        const unsigned spawn_mask = __ballot_sync(0xffffffffu, spawn);
        if (spawn_mask == 0u) break;
        const int n_spawn = __popc(spawn_mask);
        const int leader = __ffs(spawn_mask) - 1;
        const int rank = __popc(spawn_mask & ((1u << lane) - 1u));
        const bool go_on = n_spawn >= cont_min;
        int base_rec = 0, base_q = 0;
        if (lane == leader) {
          base_rec = atomicAdd(&fc->n_rec[level], n_spawn);
          atomicAdd(&fc->n_spawn[level + 1], n_spawn);
          if (!go_on) base_q = atomicAdd(&fc->n_queue[level + 1], n_spawn);
        }
        base_rec = __shfl_sync(0xffffffffu, base_rec, leader);
        base_q = __shfl_sync(0xffffffffu, base_q, leader);
        if (spawn) {
          const int rec = level * fb.n_cap + base_rec + rank;
          fb.rec_a[rec] = make_float4(colour.x, colour.y, colour.z, fres);
          fb.rec_b[rec] = make_int2((int)ty, parent);
          parent = rec;
          if (!go_on) {
            const size_t slot = (size_t)(level + 1) * cap + (size_t)(base_q + rank);
            fb.q_o[slot] = make_float4(P.x, P.y, P.z, __int_as_float(child_single ? 1 : 0));
            fb.q_d[slot] = make_float4(child_d.x, child_d.y, child_d.z, child_lp.x);
            fb.q_x[slot] = make_float4(child_lp.y, child_lp.z, __int_as_float(rec), __int_as_float(pix));
          }
        }
        if (!go_on) break;
        // the warp goes on with the child rays of the lanes that spawned one
        valid = spawn;
        o = P; d = child_d; single = child_single; lp = child_lp;
      }
    }
  }

  // ---- counters: one atomic per warp and counter ----
  warp_sum_add(&fc->ctr.shade_samples, samples_shaded);
  warp_sum_add(&fc->ctr.shadow_rays, shadow_asked);
  if (STATS) {
    clk_all = clock64() - clk_start;
    warp_sum_add(&fc->ctr.shadow_rays_traced, traced);
    warp_sum_add(&fc->ctr.box_tests, st.box_tests);
    warp_sum_add(&fc->ctr.tri_tests, st.tri_tests);
    warp_sum_add(&fc->ctr.box_tests_k2, box_k2);
    warp_sum_add(&fc->ctr.tri_tests_k2, tri_k2);
    warp_sum_add(&fc->ctr.filter_checks, st.filter_checks);
    warp_sum_add(&fc->ctr.filter_slow, st.filter_slow);
    warp_sum_add(&fc->ctr.filter_rejects, st.filter_rejects);
    if (lane == 0) {
      atomicAdd(&fc->phase_clk[0], (unsigned long long)clk_trace);
      atomicAdd(&fc->phase_clk[1], (unsigned long long)clk_shadow);
      atomicAdd(&fc->phase_clk[2], (unsigned long long)clk_shade);
      atomicAdd(&fc->phase_clk[3], (unsigned long long)clk_all);
    }
  }
}

}  // namespace rtd
