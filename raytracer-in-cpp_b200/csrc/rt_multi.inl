// rt_multi.inl -- the frame on several GPUs of one box (included at the end of rt_api.cu).
//
// The reference spreads raytraceScene's pixels over the workers of a ThreadPool that all write one shared
// pixel_data array (src/flyscene.cpp:558-629, src/ThreadPool.h:38-120).  Here the workers are GPUs: the image is
// cut into interleaved row bands (band b -> device b mod N), every device holds the whole scene and renders its
// bands, and the frame is assembled without a separate gather pass:
//   * host frame   : every device copies ITS OWN bands straight into the caller's host frame (one strided 2-D copy
//                    per device, N PCIe links in parallel instead of one);
//   * device frame : the kernels of every device store their pixels into ONE frame on device 0 through NVLink peer
//                    access (RtParams.out_full_frame), completion by CUDA events.
// Two ways to drive it, same kernels:
//   rt_multi_*                          one host process, one worker thread per device   (C++ hosts: Flyscene, rt_cli)
//   rt_shared_frame_* / rt_host_frame_* one process per device (torchrun): the frame lives in CUDA-IPC memory of
//                                       rank 0 (device frame) or in POSIX shared memory (host frame); ranks signal
//                                       through flags, no collective is needed
#include <atomic>
#include <condition_variable>
#include <fcntl.h>
#include <mutex>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

// copy the packed local bands of one rank into their global rows of a full H x W frame (host or device memory)
int copy_bands_into_frame(const RtParams &p, const void *d_local, uint8_t *frame, cudaStream_t st) {
  const size_t row_bytes = (size_t)p.width * 4;
  if (p.band_world <= 1) {
    CUDA_TRY(cudaMemcpyAsync(frame, d_local, row_bytes * (size_t)p.height, cudaMemcpyDefault, st));
    return RT_OK;
  }
  const int B = std::max(1, p.band_rows);
  const int bands = (p.height + B - 1) / B;
  // bands owned by this rank: rank, rank + world, ...; all but possibly the last are B rows tall
  int n_mine = 0, last_rows = 0;
  for (int b = p.band_rank; b < bands; b += p.band_world) { ++n_mine; last_rows = std::min(B, p.height - b * B); }
  if (n_mine == 0) return RT_OK;
  const int n_full = last_rows == B ? n_mine : n_mine - 1;
  const size_t band_bytes = row_bytes * (size_t)B;
  if (n_full > 0)
    CUDA_TRY(cudaMemcpy2DAsync(frame + (size_t)p.band_rank * band_bytes, (size_t)p.band_world * band_bytes, d_local, band_bytes,
                               band_bytes, (size_t)n_full, cudaMemcpyDefault, st));
  if (n_full < n_mine) {
    const size_t gb = (size_t)p.band_rank + (size_t)n_full * (size_t)p.band_world;  // global index of the partial band
    CUDA_TRY(cudaMemcpyAsync(frame + gb * band_bytes, (const uint8_t *)d_local + (size_t)n_full * band_bytes,
                             row_bytes * (size_t)last_rows, cudaMemcpyDefault, st));
  }
  return RT_OK;
}

int auto_band_rows(int height, int world) { return std::max(8, (height / (std::max(world, 1) * 8)) & ~7); }

}  // namespace

// rt_render for one rank of a band-sharded frame whose HOST frame is shared by all ranks: renders this rank's bands
// and copies them to their global rows of `rgba_full` ([H][W][4], page-locked for speed).  Blocking.
extern "C" int rt_render_into_frame(RtScene *sc, const RtCamera *cam, const RtLights *lights, const RtParams *p,
                                    uint8_t *rgba_full) {
  if (!sc || !cam || !lights || !p || !rgba_full) return fail(RT_ERR_INVALID, "null argument");
  int rc = use_device(sc->device);
  if (rc) return rc;
  const size_t n = (size_t)rt_local_rows(p) * (size_t)std::max(0, p->width);
  if (n == 0) return RT_OK;
  RtParams lp = *p;
  if (g_opt_host_direct && (g_opt_host_direct == 2 || n * 4 <= (size_t)g_opt_direct_max_mb << 20)) {
    // a page-locked frame is mapped into this device's address space: the kernels store this rank's pixels at their
    // global rows themselves (see rt_render), no staging buffer and no strided copy
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, rgba_full) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer != nullptr) {
      lp.out_full_frame = 1;
      g_tile_override = g_opt_direct_tile_w_log2;
      rc = rt_render_device(sc, cam, lights, &lp, at.devicePointer, nullptr, nullptr, nullptr, sc->stream, nullptr);
      g_tile_override = 0;
      if (rc) return rc;
      CUDA_TRY(cudaStreamSynchronize(sc->stream));
      return RT_OK;
    }
    cudaGetLastError();
  }
  if ((rc = sc->out_rgba.reserve(n * 4))) return rc;
  lp.out_full_frame = 0;
  if ((rc = rt_render_device(sc, cam, lights, &lp, sc->out_rgba.p, nullptr, nullptr, nullptr, sc->stream, nullptr))) return rc;
  if ((rc = copy_bands_into_frame(lp, sc->out_rgba.p, rgba_full, sc->stream))) return rc;
  CUDA_TRY(cudaStreamSynchronize(sc->stream));
  return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// one process, N devices
// ---------------------------------------------------------------------------------------------
struct RtMulti {
  int n = 0;
  std::vector<int> devices;
  std::vector<RtScene *> scenes;
  // device-resident frame on devices[0], written by every device over peer access
  DevBuf frame0;
  std::vector<cudaEvent_t> ev_begin, ev_end;
  // host frame registration cache
  void *reg_ptr = nullptr;
  size_t reg_bytes = 0;
  // worker threads (one per device) and the job they share
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv;
  std::atomic<unsigned> job_seq{0};
  std::atomic<int> done{0};
  bool quit = false;
  struct Job {
    const RtCamera *cam = nullptr;
    const RtLights *lights = nullptr;
    RtParams params{};
    uint8_t *host_frame = nullptr;  // host mode
    void *dev_frame = nullptr;      // device mode (on devices[0])
    bool want_stats = false;
  } job;
  std::vector<int> rc;
  std::vector<std::string> err;
  std::vector<RtStats> stats;
  std::vector<float> ms;
};

namespace {

void multi_run_device(RtMulti *m, int k) {
  const RtMulti::Job &j = m->job;
  RtScene *sc = m->scenes[(size_t)k];
  RtParams p = j.params;
  p.band_rank = k;
  p.band_world = m->n;
  int rc = use_device(sc->device);
  if (rc == RT_OK) {
    if (j.host_frame) {
      if (j.want_stats) {
        // diagnostic: census and per-kernel times of this device's share (blocking call)
        const size_t n = (size_t)rt_local_rows(&p) * (size_t)p.width;
        rc = sc->out_rgba.reserve(std::max<size_t>(n, 1) * 4);
        p.out_full_frame = 0;
        if (rc == RT_OK) rc = rt_render_device(sc, j.cam, j.lights, &p, sc->out_rgba.p, nullptr, nullptr, nullptr, sc->stream, &m->stats[(size_t)k]);
        if (rc == RT_OK && n) rc = copy_bands_into_frame(p, sc->out_rgba.p, j.host_frame, sc->stream);
        if (rc == RT_OK && cudaStreamSynchronize(sc->stream) != cudaSuccess) rc = fail(RT_ERR_CUDA, "stream synchronize failed");
      } else {
        rc = rt_render_into_frame(sc, j.cam, j.lights, &p, j.host_frame);
      }
    } else {
      p.out_full_frame = 1;
      cudaEventRecord(m->ev_begin[(size_t)k], sc->stream);
      rc = rt_local_rows(&p) > 0 ? rt_render_device(sc, j.cam, j.lights, &p, j.dev_frame, nullptr, nullptr, nullptr, sc->stream, nullptr) : RT_OK;
      cudaEventRecord(m->ev_end[(size_t)k], sc->stream);
      if (rc == RT_OK && cudaStreamSynchronize(sc->stream) != cudaSuccess) rc = fail(RT_ERR_CUDA, "stream synchronize failed");
      if (rc == RT_OK) cudaEventElapsedTime(&m->ms[(size_t)k], m->ev_begin[(size_t)k], m->ev_end[(size_t)k]);
    }
  }
  m->rc[(size_t)k] = rc;
  if (rc) m->err[(size_t)k] = g_err;
}

void multi_worker(RtMulti *m, int k) {
  unsigned seen = 0;
  for (;;) {
    // spin briefly (frames are sub-millisecond), then sleep on the condition variable
    unsigned seq = m->job_seq.load(std::memory_order_acquire);
    for (int spin = 0; seq == seen && spin < 20000; ++spin) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
      seq = m->job_seq.load(std::memory_order_acquire);
    }
    if (seq == seen) {
      std::unique_lock<std::mutex> lk(m->mu);
      m->cv.wait(lk, [&] { return m->quit || m->job_seq.load(std::memory_order_acquire) != seen; });
      if (m->quit) return;
      seq = m->job_seq.load(std::memory_order_acquire);
    }
    {
      std::lock_guard<std::mutex> lk(m->mu);
      if (m->quit) return;
    }
    seen = seq;
    multi_run_device(m, k);
    m->done.fetch_add(1, std::memory_order_release);
  }
}

int multi_dispatch(RtMulti *m) {
  m->done.store(0, std::memory_order_relaxed);
  {
    std::lock_guard<std::mutex> lk(m->mu);
    m->job_seq.fetch_add(1, std::memory_order_release);
  }
  m->cv.notify_all();
  while (m->done.load(std::memory_order_acquire) < m->n) {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  for (int k = 0; k < m->n; ++k)
    if (m->rc[(size_t)k]) return fail(m->rc[(size_t)k], "device %d: %s", m->devices[(size_t)k], m->err[(size_t)k].c_str());
  return RT_OK;
}

}  // namespace

extern "C" int rt_multi_create(const RtSceneDesc *desc, int n_devices, const int *devices, RtMulti **out) {
  if (!desc || !out || n_devices <= 0 || !devices) return fail(RT_ERR_INVALID, "bad argument");
  int rc = validate_scene_desc(desc);
  if (rc) return rc;
  {
    // (the same device may be listed more than once -- two scenes on one GPU, useful for testing the band assembly
    // on a single-GPU box; rt_init_devices wants each device once)
    std::vector<int> uniq;
    for (int k = 0; k < n_devices; ++k)
      if (std::find(uniq.begin(), uniq.end(), devices[k]) == uniq.end()) uniq.push_back(devices[k]);
    if ((rc = rt_init_devices((int)uniq.size(), uniq.data()))) return rc;
  }
  // Large scenes are built on every device by the device itself (rt_gpu_build.inl); otherwise the host work (BVH,
  // octree filter, soup) is done once for all devices.
  bool gpu_build = wants_gpu_build(desc);
  HostBake hb;
  if (!gpu_build) bake_scene(desc, hb);
  RtMulti *m = new RtMulti();
  m->n = n_devices;
  m->devices.assign(devices, devices + n_devices);
  m->rc.assign((size_t)n_devices, 0);
  m->err.resize((size_t)n_devices);
  m->stats.resize((size_t)n_devices);
  m->ms.assign((size_t)n_devices, 0.f);
  std::vector<RtScene *> built((size_t)n_devices, nullptr);
  if (gpu_build) {
    // every device builds its own copy of the scene, all of them at the same time (one host thread per device:
    // 8 GPUs take the 55 ms of one build instead of 8 x 55 ms)
    std::vector<int> brc((size_t)n_devices, RT_OK);
    std::vector<char> deep((size_t)n_devices, 0);
    std::vector<std::string> berr((size_t)n_devices);
    std::vector<std::thread> builders;
    for (int k = 0; k < n_devices; ++k)
      builders.emplace_back([&, k] {
        bool too_deep = false;
        int r = use_device(devices[k]);
        if (r == RT_OK) r = gpu_build_scene(desc, &built[(size_t)k], &too_deep);
        brc[(size_t)k] = r;
        deep[(size_t)k] = too_deep ? 1 : 0;
        if (r) berr[(size_t)k] = g_err;
      });
    for (auto &t : builders) t.join();
    for (int k = 0; k < n_devices; ++k) {
      if (brc[(size_t)k] == RT_OK && !deep[(size_t)k]) continue;
      // a failed build fails the call; a tree too deep for the traversal stack (never seen) sends every device to the host bake
      for (RtScene *sc : built) if (sc) { use_device(sc->device); rt_scene_destroy(sc); }
      std::fill(built.begin(), built.end(), nullptr);
      if (brc[(size_t)k] != RT_OK) { rt_multi_destroy(m); return fail(brc[(size_t)k], "device %d: %s", devices[k], berr[(size_t)k].c_str()); }
      gpu_build = false;
      bake_scene(desc, hb);
      break;
    }
  }
  for (int k = 0; k < n_devices; ++k) {
    RtScene *sc = built[(size_t)k];
    if ((rc = use_device(devices[k]))) { rt_multi_destroy(m); return rc; }
    if (!gpu_build && (rc = upload_scene(hb, &sc))) { rt_multi_destroy(m); return rc; }
    m->scenes.push_back(sc);
    cudaEvent_t a = nullptr, b = nullptr;
    cudaEventCreate(&a); cudaEventCreate(&b);
    m->ev_begin.push_back(a); m->ev_end.push_back(b);
  }
  for (int k = 0; k < n_devices; ++k) m->workers.emplace_back(multi_worker, m, k);
  use_device(devices[0]);
  *out = m;
  return RT_OK;
}

extern "C" void rt_multi_destroy(RtMulti *m) {
  if (!m) return;
  {
    std::lock_guard<std::mutex> lk(m->mu);
    m->quit = true;
    m->job_seq.fetch_add(1, std::memory_order_release);
  }
  m->cv.notify_all();
  for (auto &t : m->workers) t.join();
  if (m->reg_ptr) cudaHostUnregister(m->reg_ptr);
  for (size_t k = 0; k < m->scenes.size(); ++k) {
    use_device(m->scenes[k]->device);
    if (k < m->ev_begin.size()) { cudaEventDestroy(m->ev_begin[k]); cudaEventDestroy(m->ev_end[k]); }
    rt_scene_destroy(m->scenes[k]);
  }
  if (!m->devices.empty()) { use_device(m->devices[0]); m->frame0.release(); }
  delete m;
}

extern "C" int rt_multi_device_count(const RtMulti *m) { return m ? m->n : 0; }

extern "C" int rt_multi_scene(const RtMulti *m, int k, RtScene **out) {
  if (!m || !out || k < 0 || k >= m->n) return fail(RT_ERR_INVALID, "bad argument");
  *out = m->scenes[(size_t)k];
  return RT_OK;
}

// Flyscene::raytraceScene on N GPUs: every device renders its interleaved bands and copies them into rgba_out.
extern "C" int rt_multi_render(RtMulti *m, const RtCamera *cam, const RtLights *lights, const RtParams *p, uint8_t *rgba_out,
                               RtStats *stats) {
  if (!m || !cam || !lights || !p || !rgba_out) return fail(RT_ERR_INVALID, "null argument");
  if (p->width <= 0 || p->height <= 0) return fail(RT_ERR_INVALID, "bad image size %dx%d", p->width, p->height);
  const size_t bytes = (size_t)p->width * (size_t)p->height * 4;
  if (m->reg_ptr != rgba_out || m->reg_bytes != bytes) {
    // page-lock the caller's frame once (device -> host copies of all devices then run concurrently at full PCIe
    // speed); memory that is already page-locked (cudaMallocHost, a pinned tensor) reports "already registered"
    if (m->reg_ptr) { cudaHostUnregister(m->reg_ptr); m->reg_ptr = nullptr; }
    if (cudaHostRegister(rgba_out, bytes, cudaHostRegisterPortable) == cudaSuccess) { m->reg_ptr = rgba_out; m->reg_bytes = bytes; }
    cudaGetLastError();
  }
  m->job.cam = cam; m->job.lights = lights; m->job.params = *p;
  if (p->band_rows <= 0 || p->band_world <= 1) m->job.params.band_rows = auto_band_rows(p->height, m->n);
  m->job.host_frame = rgba_out; m->job.dev_frame = nullptr; m->job.want_stats = stats != nullptr;
  int rc = multi_dispatch(m);
  if (rc) return rc;
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    for (int k = 0; k < m->n; ++k) {
      const RtStats &s = m->stats[(size_t)k];
      stats->rays_primary += s.rays_primary; stats->rays_shadow += s.rays_shadow; stats->rays_secondary += s.rays_secondary;
      stats->pixels += s.pixels; stats->levels = std::max(stats->levels, s.levels);
      stats->ms_total = std::max(stats->ms_total, s.ms_total); stats->ms_trace = std::max(stats->ms_trace, s.ms_trace);
      stats->ms_shadow = std::max(stats->ms_shadow, s.ms_shadow); stats->ms_shade = std::max(stats->ms_shade, s.ms_shade);
      stats->kernel_launches += s.kernel_launches;
      stats->box_tests += s.box_tests; stats->tri_tests += s.tri_tests; stats->shade_samples += s.shade_samples;
      stats->box_tests_shadow += s.box_tests_shadow; stats->tri_tests_shadow += s.tri_tests_shadow;
      stats->filter_checks += s.filter_checks; stats->filter_slow += s.filter_slow; stats->filter_rejects += s.filter_rejects;
      stats->shadow_rays_traced += s.shadow_rays_traced; stats->fused = s.fused;
    }
  }
  return RT_OK;
}

// Device-resident frame: the kernels of every device store their pixels straight into one [H][W][4] frame on
// devices[0] over NVLink peer access.  *d_frame stays valid until the next call; *ms (optional) = device time of the
// slowest device (CUDA events around its frame).
extern "C" int rt_multi_render_device(RtMulti *m, const RtCamera *cam, const RtLights *lights, const RtParams *p,
                                      void **d_frame, float *ms) {
  if (!m || !cam || !lights || !p || !d_frame) return fail(RT_ERR_INVALID, "null argument");
  if (p->width <= 0 || p->height <= 0) return fail(RT_ERR_INVALID, "bad image size %dx%d", p->width, p->height);
  int rc = use_device(m->devices[0]);
  if (rc) return rc;
  if ((rc = m->frame0.reserve((size_t)p->width * (size_t)p->height * 4))) return rc;
  m->job.cam = cam; m->job.lights = lights; m->job.params = *p;
  if (p->band_rows <= 0 || p->band_world <= 1) m->job.params.band_rows = auto_band_rows(p->height, m->n);
  m->job.host_frame = nullptr; m->job.dev_frame = m->frame0.p; m->job.want_stats = false;
  if ((rc = multi_dispatch(m))) return rc;
  *d_frame = m->frame0.p;
  if (ms) { *ms = 0.f; for (int k = 0; k < m->n; ++k) *ms = std::max(*ms, m->ms[(size_t)k]); }
  use_device(m->devices[0]);
  return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// one process per device: completion flags in the shared device frame (no collective)
// ---------------------------------------------------------------------------------------------
// The allocation made by rt_shared_frame_create is the frame followed by kFlagBytes of flags; flag r (128 bytes
// apart) holds the sequence number of the last frame rank r has finished storing.
namespace {
constexpr size_t kFlagBytes = 4096;
constexpr size_t kFlagStride = 128;

__global__ void k_frame_signal(unsigned int *flag, unsigned int seq) {
  __threadfence_system();  // everything this device stored before (stream order) is visible before the flag is
  *reinterpret_cast<volatile unsigned int *>(flag) = seq;
}

// one lane per rank polls that rank's flag; gives up after ~2 s of GPU clock so that a dead rank cannot hang the GPU
__global__ void k_frame_wait(const unsigned int *flags, int world, unsigned int seq, unsigned int *timeout_flag) {
  const int r = threadIdx.x;
  if (r < world) {
    const volatile unsigned int *f = reinterpret_cast<const volatile unsigned int *>(reinterpret_cast<const char *>(flags) + (size_t)r * kFlagStride);
    const long long t0 = clock64();
    while ((int)(*f - seq) < 0) {
      __nanosleep(100);
      if (clock64() - t0 > 4000000000LL) { atomicExch(timeout_flag, 1u); break; }
    }
  }
  __threadfence_system();
}
}  // namespace

extern "C" int rt_shared_frame_signal(void *d_frame, size_t frame_bytes, int rank, unsigned int seq, void *stream) {
  if (!d_frame || rank < 0 || (size_t)rank * kFlagStride >= kFlagBytes - kFlagStride) return fail(RT_ERR_INVALID, "bad argument");
  unsigned int *flag = reinterpret_cast<unsigned int *>(reinterpret_cast<char *>(d_frame) + ((frame_bytes + 255) & ~(size_t)255) + (size_t)rank * kFlagStride);
  k_frame_signal<<<1, 1, 0, (cudaStream_t)stream>>>(flag, seq);
  CUDA_TRY(cudaGetLastError());
  return RT_OK;
}

extern "C" int rt_shared_frame_wait(void *d_frame, size_t frame_bytes, int world, unsigned int seq, void *stream) {
  if (!d_frame || world <= 0 || world > 31) return fail(RT_ERR_INVALID, "bad argument");
  unsigned int *flags = reinterpret_cast<unsigned int *>(reinterpret_cast<char *>(d_frame) + ((frame_bytes + 255) & ~(size_t)255));
  unsigned int *timeout_flag = reinterpret_cast<unsigned int *>(reinterpret_cast<char *>(flags) + kFlagBytes - kFlagStride);
  k_frame_wait<<<1, 32, 0, (cudaStream_t)stream>>>(flags, world, seq, timeout_flag);
  CUDA_TRY(cudaGetLastError());
  return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// one process per device: a HOST frame in POSIX shared memory, page-locked in every process, plus a barrier
// ---------------------------------------------------------------------------------------------
struct RtHostFrame {
  std::string name;
  size_t bytes = 0, map_bytes = 0;
  void *base = nullptr;
  bool owner = false, registered = false;
  unsigned local_sense = 0;
};
namespace {
struct HostFrameHeader {  // lives in the first 256 bytes of the mapping
  std::atomic<unsigned> count;
  std::atomic<unsigned> sense;
};
}  // namespace

extern "C" int rt_host_frame_open(const char *name, size_t bytes, int create, RtHostFrame **out) {
  if (!name || !out || bytes == 0) return fail(RT_ERR_INVALID, "bad argument");
  const size_t map_bytes = ((bytes + 4095) & ~(size_t)4095) + 4096;
  int fd = shm_open(name, create ? (O_CREAT | O_RDWR | O_TRUNC) : O_RDWR, 0600);
  if (fd < 0) return fail(RT_ERR_IO, "shm_open(%s) failed", name);
  if (create && ftruncate(fd, (off_t)map_bytes) != 0) { close(fd); shm_unlink(name); return fail(RT_ERR_IO, "ftruncate(%s) failed", name); }
  void *base = mmap(nullptr, map_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (base == MAP_FAILED) { if (create) shm_unlink(name); return fail(RT_ERR_IO, "mmap(%s) failed", name); }
  RtHostFrame *h = new RtHostFrame();
  h->name = name; h->bytes = bytes; h->map_bytes = map_bytes; h->base = base; h->owner = create != 0;
  if (create) {
    HostFrameHeader *hd = new (base) HostFrameHeader();
    hd->count.store(0); hd->sense.store(0);
  }
  // page-lock it for this process's device copies (skipped without a device: the mapping still works as plain memory)
  if (g_default_device >= 0 || ensure_device() == RT_OK) {
    h->registered = cudaHostRegister(base, map_bytes, cudaHostRegisterPortable) == cudaSuccess;
    cudaGetLastError();
  }
  *out = h;
  return RT_OK;
}

extern "C" void *rt_host_frame_ptr(RtHostFrame *h) { return h ? (void *)((char *)h->base + 4096) : nullptr; }

// sense-reversing barrier of `world` processes on the shared header (spins; microseconds)
extern "C" int rt_host_frame_barrier(RtHostFrame *h, int world) {
  if (!h || world <= 0) return fail(RT_ERR_INVALID, "bad argument");
  HostFrameHeader *hd = reinterpret_cast<HostFrameHeader *>(h->base);
  h->local_sense ^= 1u;
  if (hd->count.fetch_add(1, std::memory_order_acq_rel) == (unsigned)world - 1) {
    hd->count.store(0, std::memory_order_relaxed);
    hd->sense.store(h->local_sense, std::memory_order_release);
  } else {
    const auto t0 = std::chrono::steady_clock::now();
    unsigned spins = 0;
    while (hd->sense.load(std::memory_order_acquire) != h->local_sense) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
      if ((++spins & 0xfffffu) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(60))
        return fail(RT_ERR_IO, "host frame barrier timed out (a rank is missing)");
    }
  }
  return RT_OK;
}

extern "C" void rt_host_frame_close(RtHostFrame *h) {
  if (!h) return;
  if (h->registered) cudaHostUnregister(h->base);
  munmap(h->base, h->map_bytes);
  if (h->owner) shm_unlink(h->name.c_str());
  delete h;
}
