/*
 * rt_api.h -- the drop-in boundary of the B200 render path (C ABI, no torch / C++ types).
 *
 * The reference (Sh-Anand/Raytracer-in-CPP) has no FFI: its hot path is a set of C++ member
 * functions of Flyscene / BoxTree / BoundingBox / arealight.  Each entry point
 * below names the reference interface it replaces (file:line relative to /root/reference); the
 * C++ facade in raytracer-in-cpp_b200/host/flyscene.hpp keeps the reference's class and method
 * names and forwards to these functions, and INTEGRATION.md shows the binding a maintainer of the
 * reference would add inside src/flyscene.cpp.
 *
 * Conventions (all taken from the reference):
 *   - ray directions are NOT normalised; t is in units of |d|
 *   - pixel (i, j): i = column, j = row, row 0 is the top of the image, integer corner (no +0.5)
 *   - colours are float RGB in [0,1]; 8-bit = min(255, (int)(255*c))   (tucano/utils/ppmIO.hpp:145)
 *   - BACKGROUND = (1,1,1), SHADOW = (0,0,0)                            (src/flyscene.cpp:12-13)
 *   - "no hit" face id = -1, t = FLT_MAX
 *
 * Every function returns 0 on success or a negative RtStatus; rt_last_error() gives the message
 * (thread-local).  There is NO CPU fallback: without a CUDA device every compute entry point fails
 * with RT_ERR_NO_DEVICE.  All calls are blocking unless a stream is passed.
 *
 * Threading: like the reference, whose raytraceScene() is called from the main thread and blocks
 * (SURVEY.md 8b), the library is driven by ONE host thread per process; one process per GPU
 * (rt_init selects the device).  Options (rt_set_option) and the per-scene workspaces are not locked.
 */
#ifndef RT_API_H
#define RT_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_API_VERSION 3
#define RT_MAX_LIGHTS 25   /* bool visibleLights[25], src/flyscene.cpp:699,835 */
#define RT_MAX_SAMPLES 25

typedef enum {
  RT_OK = 0,
  RT_ERR_INVALID = -1,    /* bad argument */
  RT_ERR_NO_DEVICE = -2,  /* no CUDA device / CUDA runtime failure at init */
  RT_ERR_CUDA = -3,       /* CUDA call failed; see rt_last_error() */
  RT_ERR_IO = -4,         /* file could not be read / written */
  RT_ERR_LIMIT = -5       /* more than RT_MAX_LIGHTS lights or RT_MAX_SAMPLES samples */
} RtStatus;

/* Tucano::Material::Mtl as read by the hot path (tucano/materials/mtl.hpp:108-114) */
typedef struct {
  float kd[3];
  float ks[3];
  float ns;      /* shininess */
  float ni;      /* optical density */
  int32_t illum; /* illumination model, drives the material switch src/flyscene.cpp:712-760 */
} RtMaterial;

/* The baked scene: exactly the values the reference's inner loops read per face
 * (src/flyscene.cpp:788-792, 867-872, 712-713).  Caller keeps ownership; rt_scene_create copies. */
typedef struct {
  int32_t n_faces;
  const float *verts;           /* [T][3][3] world space: (mesh.getShapeModelMatrix()*v).head<3>() */
  const float *face_normals;    /* [T][3]    Tucano::Face::normal */
  const float *vertex_normals;  /* [T][3][3] mesh.getNormal(face.vertex_ids[k]) */
  const int32_t *material_id;   /* [T]       Tucano::Face::material_id */
  int32_t n_materials;
  const RtMaterial *materials;
  float model_matrix[12];       /* 3x4 row-major mesh.getModelMatrix(), applied to the shading normal (:829) */
  int32_t n_spheres;            /* analytic spheres: NOT in the reference (BASELINE config 4), may be 0 */
  const float *spheres;         /* [S][4] centre, radius */
  const int32_t *sphere_material; /* [S] */
} RtSceneDesc;

/* Tucano::Flycamera state read by raytraceScene (src/flyscene.cpp:551,575; tucano/camera.hpp:115-173) */
typedef struct {
  float eye[3];        /* flycamera.getCenter() */
  float view_inv[12];  /* flycamera.getViewMatrix().inverse(), 3x4 row-major */
  float viewport[4];   /* (0, 0, w, h) */
  float fovy;          /* degrees (60) */
  float aspect;        /* w / (float)h */
} RtCamera;

/* Flyscene::lights + lightrep colour (src/flyscene.cpp:68,72) */
typedef struct {
  int32_t n;           /* <= RT_MAX_LIGHTS */
  const float *pos;    /* [n][3] */
  float color[3];      /* (1,1,0) in the reference */
} RtLights;

typedef struct {
  int32_t width, height;
  int32_t area_light;        /* stdin flag, src/flyscene.cpp:31-32 */
  int32_t point_light;       /* stdin flag, src/flyscene.cpp:33-34 (wins over area_light, :964) */
  int32_t max_depth;         /* rays at level >= max_depth shade as plain Phong; < 0 = reference
                                behaviour (unbounded; guarded at 64 levels) */
  int32_t usteps, vsteps;    /* area-light grid, reference 5 x 5 (src/flyscene.cpp:971) */
  float area_len_x, area_len_y; /* 0.3, 0.15 */
  /* multi-GPU row-band sharding: this call renders bands b with b % band_world == band_rank,
   * band = band_rows full-width rows; output buffers then hold only the local rows, packed in
   * ascending row order.  band_world <= 1 renders the whole image. */
  int32_t band_rows, band_rank, band_world;
  /* rt_render_device only: 1 = d_rgba is the FULL height x width frame and every rendered row is
   * stored at its global position (used to write bands straight into rank 0's framebuffer over an
   * NVLink peer mapping, see rt_shared_frame_*); 0 = d_rgba holds the local rows, packed. */
  int32_t out_full_frame;
  /* area_light = point_light = 0: the reference's "spherical" light (src/flyscene.cpp:974-993), 25 points
   * light + R*(sin(phi)cos(theta), sin(phi)sin(theta), cos(phi))/5, theta = 2*pi*u, phi = acos(2u-1).
   * The reference draws u from a freshly seeded mt19937(random_device) for every point of every shading
   * call, which no two runs reproduce; here the 25 values of u are fixed per frame:
   *   u_k = (lowbias32(sphere_seed * 0x9E3779B9u + k) >> 8) / 2^24,  k = 0..24
   * (lowbias32: x^=x>>16; x*=0x7feb352d; x^=x>>15; x*=0x846ca68b; x^=x>>16), so a frame is a pure function
   * of its inputs.  sphere_radius = lightrep.getBoundingSphereRadius() (1.0000001f in the reference). */
  uint32_t sphere_seed;
  float sphere_radius;
} RtParams;

typedef struct {
  int64_t rays_primary;     /* one per rendered pixel */
  int64_t rays_shadow;      /* any-hit queries of the reference's census: L gate rays per hit + L*S sample rays
                               for the hits whose gate passed (src/flyscene.cpp:699-710,836); merged in point mode */
  int64_t rays_secondary;   /* reflection / refraction / pass-through rays */
  int64_t pixels;           /* pixels rendered by this call */
  int32_t levels;           /* bounce levels executed */
  float ms_total;           /* device time of the whole call (CUDA events) */
  float ms_trace;           /* nearest-hit kernels (K1) */
  float ms_shadow;          /* any-hit kernels (K2) */
  float ms_shade;           /* shade / compaction / fold / framebuffer kernels (K3) */
  int32_t kernel_launches;  /* kernels launched by this call */
  /* traversal work counters, filled only when rt_set_option("stats", 1):
   * ray-AABB and ray-triangle tests of the nearest-hit kernels (K1) and of the shadow kernels (K2) */
  int64_t box_tests, tri_tests, shade_samples;
  int64_t box_tests_shadow, tri_tests_shadow;
  /* reference-candidate filter: winners checked, checks that needed the exact octree walk, rejections */
  int64_t filter_checks, filter_slow, filter_rejects;
  /* any-hit queries the shadow kernels actually traced ("stats" option): equal to rays_shadow -- the sample rays of a
   * hit whose gate failed are neither asked for by the reference nor traced here */
  int64_t shadow_rays_traced;
  /* which frame path ran: 0 = wavefront kernels replayed as one CUDA graph (depth caps 0..8), 1 = ONE persistent kernel
   * (csrc/rt_frame.cuh; ms_trace / ms_shadow / ms_shade are then that kernel's duration split by the warp-cycles its
   * phases took ("stats" option), not separate launches), 2 = wavefront kernels launched level by level with a host
   * read-back of the next level's ray count after each (unbounded depth, or a depth cap above 8) */
  int32_t fused;
} RtStats;

typedef struct RtScene RtScene;   /* device-resident BVH + triangle soup + shading tables */
typedef struct RtMesh RtMesh;     /* host-side loaded + baked OBJ/MTL */

/* ---- lifecycle ----------------------------------------------------------------------------- */
int rt_api_version(void);
/* Select the CUDA device for this process (one process per GPU).  Replaces nothing in the
 * reference (CPU only); ThreadPool construction src/flyscene.cpp:609 is the closest analogue. */
int rt_init(int device);
/* Several GPUs of one box (the surveyed rt_init(device_count, devices)): validates the list, makes devices[0] the
 * default device and enables peer access between the devices, so that kernels of one GPU can store pixels into a
 * framebuffer on another over NVLink.  Used by rt_multi_create; one host process. */
int rt_init_devices(int count, const int *devices);
/* Releases what the library allocated lazily for rendering (ray queues, frame graphs, staging buffers) on every
 * live scene and forgets the device selection; scenes stay valid (their workspace is re-created on demand). */
void rt_shutdown(void);
const char *rt_last_error(void);
int rt_device_name(char *buf, size_t n);
/* runtime knobs: "stats" (0/1 traversal counters), "leaf_size", "persistent_ctas_per_sm",
 * "reference_candidates" (default 1: scenes created afterwards filter BVH hits through the
 * reference's octree candidate sets so that the image matches the reference bit for bit; 0: plain
 * BVH = exact nearest hit over all faces), "graph_conditionals" (1, default: empty bounce levels are skipped inside the frame's CUDA graph),
 * "fused_frame" (the frame as one persistent kernel, csrc/rt_frame.cuh: 0 never, 1 whenever eligible, 2 = default:
 *   for scenes without spheres and without an octree filter (<= 1000 triangles, e.g. the bundled cube), for frames of
 *   at most "fused_max_kpixels" thousand rays (default 1200) and for unbounded depth; other frames take the
 *   per-level wavefront kernels of csrc/rt_kernels.cuh -- both paths give bit-identical frames),
 * "host_direct" (rt_render: when rgba_out is page-locked host memory -- cudaHostAlloc, cudaHostRegister, a pinned tensor --
 *   the kernels store the finished pixels straight into it and no device->host copy follows; 1 = default: frames of at
 *   most "host_direct_max_mb" MB (default 16), 2: frames of any size, 0: never; such frames are traced with
 *   2^k x (32 >> k) pixel tiles per warp, k = "host_direct_tile_w_log2" (default 5: 32 x 1, one 128-byte store per warp);
 *   "tile_w_log2" (default 3: 8 x 4) is the tile of every other frame),
 * "tile_order" (1 = default: the pixel tiles inside the screen rectangle of the scene's bounds are handed out first, so
 *   that a launch ends on cheap background tiles -- except for frames written straight to a host frame, whose pixel
 *   stores must stay spread over the frame time; 0 = row-major order; 2 = always first),
 * "tile_cull" (1 = default: pixel tiles outside that rectangle -- widened by one tile -- cannot see the scene: their
 *   pixels are BACKGROUND without a ray being generated; 0 = every pixel runs the reference's two root-box tests),
 * "render_chunks" (rt_render with a staged copy -- pageable or large host frames: a scene that takes the fused kernel at
 *   any size is rendered in this many row chunks, 1..4, default 2, and the device->host copy of a chunk overlaps the
 *   rendering of the next; frames of 32 MB and more are rendered in 4 chunks on any scene),
 * "continue_min_lanes" (fused frame: child rays stay in their warp when at least this many of its lanes spawned one; default 8),
 * "donate_min_lanes" (scenes with an octree filter or spheres, shadow kernel of launches with few rays per resident warp:
 *   lanes whose ray is finished take over pending subtrees of the lanes still traversing once at least this many lanes
 *   are idle; 0 = never; default 12; 100 + n = in every nearest-hit and shadow launch),
 * "gpu_build" / "gpu_build_min_prims" (see rt_scene_build_info) */
int rt_set_option(const char *key, int value);
void rt_default_params(RtParams *p);

/* ---- scene bake (host) --------------------------------------------------------------------- */
/* Tucano::MeshImporter::loadObjFile + loadMTL + Mesh::normalizeModelMatrix
 * (tucano/utils/objimporter.hpp:83-284, mtlIO.hpp:45-125, model.hpp:169-173; src/flyscene.cpp:50-56) */
int rt_mesh_load_obj(const char *obj_path, RtMesh **out);
/* fills desc with pointers into the mesh (valid until rt_mesh_destroy) */
int rt_mesh_desc(const RtMesh *mesh, RtSceneDesc *desc);
/* normalisation data: centroid[3], radius, scale  (tucano/mesh.hpp:621-642) */
int rt_mesh_info(const RtMesh *mesh, float centroid[3], float *radius, float *scale, int32_t *n_vertices);
void rt_mesh_destroy(RtMesh *mesh);

/* ---- acceleration structure ---------------------------------------------------------------- */
/* BoxTree::BoxTree(Mesh&, capacity) + BoundingBox::BoundingBox(Mesh&)
 * (src/boxTree.cpp:11-31,88-147; src/boundingBox.cpp:14-43; called at src/flyscene.cpp:93).
 * Builds the flattened BVH on the host and uploads everything once. */
int rt_scene_create(const RtSceneDesc *desc, RtScene **out);
void rt_scene_destroy(RtScene *scene);
/* BoxTree::box (root AABB, reference semantics incl. the FLT_MIN max initialiser) */
int rt_scene_root_box(const RtScene *scene, float mn[3], float mx[3]);
/* node / leaf / triangle counts, bytes resident on the device, build milliseconds */
int rt_scene_info(const RtScene *scene, int64_t *n_nodes, int64_t *n_leaves, int64_t *n_tris,
                  int64_t *device_bytes, float *build_ms);
/* Shape of the reference octree for this scene (BoxTree, capacity 1000, depth 15): out[4] =
 * reachable leaves, inner nodes, face references, largest leaf.  Host only (no GPU needed). */
int rt_ref_octree_stats(const RtSceneDesc *desc, int32_t capacity, int64_t out[4]);
/* Host only (no GPU needed): builds the BVH rt_scene_create would build over the triangles of desc with
 * at most leaf_size triangles per leaf and verifies what the device traversal relies on -- every face in
 * exactly one leaf, every leaf box containing its faces, every child box inside its parent's, every node
 * reachable once, depth within the traversal stack.  out[6] = nodes, leaves, depth, largest leaf, faces
 * referenced, 1000 x SAH cost.  RT_ERR_INVALID (+ message) names the first violated invariant. */
int rt_bvh_check(const RtSceneDesc *desc, int32_t leaf_size, int64_t out[6]);
/* Which builder made the scene's acceleration structures and what came out.  Scenes of at least
 * "gpu_build_min_prims" primitives (rt_set_option, default 20000; option "gpu_build": 0 never, 1 from 64 primitives on,
 * 2 = that threshold) are built ON THE DEVICE: the reference octree (BoxTree::BoxTree / split / clasifyFace,
 * src/boxTree.cpp:11-31, 88-147, 203-336) level by level, the BVH top down by binned surface-area-heuristic
 * splits over a Morton order (one launch per tree level), the primitive soup and the shading table by one thread per primitive (csrc/rt_build.cuh).  out[16] =
 * built on the GPU (0/1), BVH pair nodes, leaves, depth, 1000 x SAH cost (the definition of rt_bvh_check),
 * octree leaves / inner nodes / face references / largest leaf (the numbers of rt_ref_octree_stats),
 * build microseconds: total, input upload, octree, Morton sort, BVH splits, emit + bake; 1000 x octree levels +
 * BVH levels. */
int rt_scene_build_info(const RtScene *scene, int64_t out[16]);
/* debug / test access to the flattened BVH (host copies): nodes [n_nodes][16] floats as uploaded,
 * tri_face [n_tris] original face id of each soup slot */
int rt_scene_debug_bvh(const RtScene *scene, float *nodes, int64_t nodes_cap, int32_t *tri_face, int64_t tri_cap);

/* ---- the frame ----------------------------------------------------------------------------- */
/* Flyscene::raytraceScene (src/flyscene.cpp:519-648) minus the PPM write: host output buffers.
 *   rgba_out : [rows][W][4] uint8 (A = 255), required
 *   face_out : [rows][W] int32 primary-hit face id (-1 = none), optional
 *   t_out    : [rows][W] float primary-hit t, optional
 *   rgb_f32_out : [rows][W][3] float colour before quantisation, optional
 * rows = H, or the local row count under band sharding (rt_local_rows). */
int rt_render(RtScene *scene, const RtCamera *cam, const RtLights *lights, const RtParams *params,
              uint8_t *rgba_out, int32_t *face_out, float *t_out, float *rgb_f32_out, RtStats *stats);
/* Same, outputs are DEVICE pointers (may be NULL except d_rgba); stream = cudaStream_t or NULL.
 * Asynchronous with respect to the host when stats == NULL. */
int rt_render_device(RtScene *scene, const RtCamera *cam, const RtLights *lights, const RtParams *params,
                     void *d_rgba, int32_t *d_face, float *d_t, float *d_rgb_f32, void *stream, RtStats *stats);
/* Frame sequences -- the reference re-renders on every camera key (src/main.cpp:59-92) and overlapping
 * the output of frame k with the tracing of frame k+1 is SURVEY.md 8(f)2.  rt_render_submit queues one
 * frame plus the device->host copy of its packed RGBA and returns without waiting; *ticket names it.
 * rt_render_wait blocks until that frame is complete in its rgba_out.  At most two frames are in
 * flight (double-buffered device frame, copy on its own stream): submitting a third before waiting for
 * the oldest is RT_ERR_INVALID.  rgba_out should be page-locked; it must stay valid until the wait. */
int rt_render_submit(RtScene *scene, const RtCamera *cam, const RtLights *lights, const RtParams *params,
                     uint8_t *rgba_out, int *ticket);
int rt_render_wait(RtScene *scene, int ticket);
int rt_local_rows(const RtParams *params);
/* global row index of each local row (ascending); rows_out has rt_local_rows entries */
int rt_local_row_map(const RtParams *params, int32_t *rows_out);

/* ---- multi-GPU: a framebuffer shared between the per-GPU processes (CUDA IPC over NVLink) ------------
 * Replaces the reference's shared pixel_data array written by all pool workers (src/flyscene.cpp:620). */
int rt_shared_frame_create(size_t bytes, void **d_ptr, unsigned char handle[64]); /* owner (rank 0) */
int rt_shared_frame_open(const unsigned char handle[64], void **d_ptr);           /* other ranks */
int rt_shared_frame_close(void *d_ptr, int owner);
int rt_device_copy_to_host(void *host, const void *d_ptr, size_t bytes);
/* Completion without a collective: the shared allocation carries one flag per rank behind the frame.  After its
 * frame (rt_render_device with out_full_frame = 1 on `stream`) every rank queues rt_shared_frame_signal(seq) on the
 * same stream; rank 0 queues rt_shared_frame_wait(seq), a one-warp kernel that polls the flags in its own memory
 * (it gives up after ~2 s so that a dead rank cannot hang the GPU).  seq must increase from frame to frame.
 * Replaces pool.enqueue / result.get() of the reference's ThreadPool (src/flyscene.cpp:613-629). */
int rt_shared_frame_signal(void *d_frame, size_t frame_bytes, int rank, unsigned int seq, void *stream);
int rt_shared_frame_wait(void *d_frame, size_t frame_bytes, int world, unsigned int seq, void *stream);

/* A HOST frame shared by the per-GPU processes: POSIX shared memory (shm name, e.g. "/rt_frame_<pid>"), page-locked
 * in every process that opens it.  create != 0 on exactly one rank, before the others open it.  Every rank copies
 * its own bands into it (rt_render_into_frame): N PCIe links carry the frame instead of rank 0's alone.
 * rt_host_frame_barrier is a spinning barrier of `world` processes on the shared header (frame complete / frame
 * consumed); it times out after 60 s. */
typedef struct RtHostFrame RtHostFrame;
int rt_host_frame_open(const char *name, size_t bytes, int create, RtHostFrame **out);
void *rt_host_frame_ptr(RtHostFrame *frame);
int rt_host_frame_barrier(RtHostFrame *frame, int world);
void rt_host_frame_close(RtHostFrame *frame);
/* rt_render for one rank of a band-sharded frame (params->band_*): renders this rank's bands and copies them to
 * their global rows of rgba_full, the full [H][W][4] host frame all ranks share.  Blocking.  band_world <= 1: = rt_render. */
int rt_render_into_frame(RtScene *scene, const RtCamera *cam, const RtLights *lights, const RtParams *params,
                         uint8_t *rgba_full);

/* ---- multi-GPU from ONE host process --------------------------------------------------------------------------
 * Flyscene::raytraceScene's ThreadPool fan-out (src/flyscene.cpp:558-629) with GPUs as the workers: the scene is
 * baked once on the host and uploaded to every device; device k renders the interleaved row bands k, k+N, ...
 * (band height params->band_rows, 0 = chosen by the library; band_rank / band_world are set by the library) and
 * one worker thread per device submits its share, so the devices start within microseconds of each other. */
typedef struct RtMulti RtMulti;
int rt_multi_create(const RtSceneDesc *desc, int n_devices, const int *devices, RtMulti **out);
void rt_multi_destroy(RtMulti *multi);
int rt_multi_device_count(const RtMulti *multi);
int rt_multi_scene(const RtMulti *multi, int k, RtScene **scene_out); /* device k's scene, for the per-function calls */
/* rgba_out: host [H][W][4]; every device copies its own bands straight into it.  stats (optional, diagnostic, slower):
 * census and work counters summed over the devices, times = max over the devices. */
int rt_multi_render(RtMulti *multi, const RtCamera *cam, const RtLights *lights, const RtParams *params,
                    uint8_t *rgba_out, RtStats *stats);
/* The frame stays on the GPU: the kernels of every device store their pixels into one [H][W][4] frame on devices[0]
 * over NVLink peer access.  *d_frame: that frame (valid until the next call).  *ms (optional): device time of the
 * slowest device (CUDA events around its share). */
int rt_multi_render_device(RtMulti *multi, const RtCamera *cam, const RtLights *lights, const RtParams *params,
                           void **d_frame, float *ms);

/* ---- per-function entry points (batched) --------------------------------------------------- */
/* Flyscene::traceRay (src/flyscene.cpp:651-771) for n arbitrary rays: rgb_out [n][3] float,
 * face_out / t_out optional.  level-0 semantics (lights = the scene lights). Host pointers. */
int rt_trace_rays(RtScene *scene, int64_t n, const float *origins, const float *dirs, const RtLights *lights,
                  const RtParams *params, float *rgb_out, int32_t *face_out, float *t_out);
/* Flyscene::lightStrikes (src/flyscene.cpp:912-954): n hit points x lights->n lights,
 * visible_out [n][lights->n] uint8 */
int rt_light_strikes(RtScene *scene, int64_t n, const float *hit_points, const RtLights *lights,
                     uint8_t *visible_out);
/* BoundingBox::boxIntersect (src/boundingBox.cpp:48-83) against the scene root box for n segments
 * origin -> dest; hit_out [n] uint8 */
int rt_box_intersect(RtScene *scene, int64_t n, const float *origins, const float *dests, uint8_t *hit_out);
/* BoundingBox::boxIntersect (src/boundingBox.cpp:48-83) for an arbitrary box (BoundingBox(min,max)) */
int rt_box_intersect_box(const float mn[3], const float mx[3], int64_t n, const float *origins, const float *dests,
                         uint8_t *hit_out);
/* Flyscene::rayTriangleIntersection (src/flyscene.cpp:787-819) for n (ray, face id) pairs;
 * t_out[i] = t or the reference's miss sentinel -72 */
int rt_ray_triangle(RtScene *scene, int64_t n, const float *origins, const float *dirs, const int32_t *faces,
                    float *t_out);
/* BoxTree::intersect (src/boxTree.cpp:150-173): the candidate face ids the reference's octree offers
 * for the query origin -> dest, ascending (the std::set order).  Returns the count (may exceed cap;
 * only the first cap ids are written) or < 0. */
int rt_octree_candidates(RtScene *scene, const float origin[3], const float dest[3], int32_t *ids, int32_t cap);
/* Flyscene::phongShade (src/flyscene.cpp:822-859) for n (ray origin, hit point, face id) triples;
 * shadow rays for every light sample are traced as lightStrikes does.  rgb_out [n][3] */
int rt_phong_shade(RtScene *scene, int64_t n, const float *origins, const float *hit_points, const int32_t *faces,
                   const RtLights *lights, const RtParams *params, float *rgb_out);
/* Camera::screenToWorld (tucano/camera.hpp:155-173) for n pixel coordinates; out [n][3] */
int rt_screen_to_world(const RtCamera *cam, int64_t n, const float *pixels_xy, float *out);
/* Flyscene::createSpherePoint / arealight::getPointLights (src/flyscene.cpp:962-972,
 * arealight.hpp:15-25): host-side, returns the sample count (<= RT_MAX_SAMPLES) or < 0 */
int rt_light_samples(const RtParams *params, const float light[3], float *out /*[25][3]*/);

/* ---- output -------------------------------------------------------------------------------- */
/* Device self-test of the shading kernels' shared-reciprocal vector division (x/n, y/n, z/n with one
 * reciprocal; must equal IEEE division bit for bit): n_trials random operand sets, exponents within
 * +-exp_range of 1.0 (exp_range <= 0: unrestricted bit patterns incl. zero, denormal, inf, NaN).
 * *mismatches receives the number of trials whose three quotients are not all identical. */
int rt_selftest_div3(int64_t n_trials, uint32_t seed, int32_t exp_range, int64_t *mismatches);

/* Tucano::ImageImporter::writePPMImage (tucano/utils/ppmIO.hpp:130-151): ASCII P3, "r g b " per
 * pixel, one text line per image row.  binary != 0 writes P6 instead. */
int rt_write_ppm(const char *path, const uint8_t *rgba, int32_t width, int32_t height, int32_t binary);

#ifdef __cplusplus
}
#endif
#endif /* RT_API_H */
