/*
 * ref_driver.cpp -- TEST INFRASTRUCTURE (oracle/_ref build only).
 *
 * Headless driver around the UNMODIFIED reference ray tracer.  It is compiled against the
 * reference's own translation units where they lie under /root/reference (see oracle/Makefile)
 * and calls the reference's own public entry points:
 *
 *   Flyscene::initialize            /root/reference/src/flyscene.cpp:29
 *   Flycamera::screenToWorld        dependencies/tucano/tucano/camera.hpp:155
 *   BoundingBox::boxIntersect       src/boundingBox.cpp:48
 *   BoxTree::intersect              src/boxTree.cpp:150
 *   Flyscene::rayTriangleIntersection  src/flyscene.cpp:787
 *   Flyscene::traceRay              src/flyscene.cpp:651
 *   Flyscene::raytraceScene         src/flyscene.cpp:519   (--mode rts)
 *
 * The per-pixel loop below is the reference's own loop (src/flyscene.cpp:573-598, 613-625)
 * restated so that it also works for W != H (the reference indexes pixel_data[x][y] on a
 * [H][W] array and crashes) and so that float RGB, primary face id and t can be captured.
 *
 * Outputs (little-endian binary, documented in oracle/FORMATS.md):
 *   --dump-scene FILE   the baked scene the reference actually traces (world-space vertices,
 *                       face normals, per-corner vertex normals, materials, camera, lights)
 *   --out FILE          per-pixel float RGB + primary hit (face id, t) for the sampled pixels
 *
 * Compiled with -fno-access-control to read Flyscene's private members.
 */
#include "flyscene.hpp"

#include <atomic>
#include <csignal>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <sys/types.h>
#include <dirent.h>
#include <unistd.h>

#ifdef RT_PATCHED
// knobs read by the sed-patched flyscene TU (see oracle/Makefile)
int rt_oracle_max_depth = 1 << 30;
int rt_oracle_usteps = 5;
int rt_oracle_vsteps = 5;
#endif

static void die(const char *msg) {
  fprintf(stderr, "ref_driver: %s\n", msg);
  exit(2);
}

static std::string abspath(const std::string &p) {
  char buf[4096];
  if (!realpath(p.c_str(), buf)) die(("cannot resolve path " + p).c_str());
  return buf;
}

// The reference hard-codes "resources/models/cube.obj" relative to the CWD
// (src/flyscene.cpp:50-51): build a scratch CWD whose cube.obj is the wanted OBJ.
static std::string make_scratch_cwd(const std::string &obj_abs) {
  char tmpl[] = "/tmp/rt_ref_XXXXXX";
  char *d = mkdtemp(tmpl);
  if (!d) die("mkdtemp failed");
  std::string root = d;
  mkdir((root + "/resources").c_str(), 0755);
  std::string models = root + "/resources/models";
  mkdir(models.c_str(), 0755);
  std::string src_dir = obj_abs.substr(0, obj_abs.find_last_of('/'));
  DIR *dir = opendir(src_dir.c_str());
  if (!dir) die("cannot open OBJ directory");
  while (dirent *e = readdir(dir)) {
    std::string n = e->d_name;
    if (n == "." || n == ".." || n == "cube.obj") continue;
    symlink((src_dir + "/" + n).c_str(), (models + "/" + n).c_str());
  }
  closedir(dir);
  if (symlink(obj_abs.c_str(), (models + "/cube.obj").c_str()) != 0) die("symlink cube.obj failed");
  return root;
}

template <class T> static void wr(FILE *f, const T &v) { fwrite(&v, sizeof(T), 1, f); }
static void wr3(FILE *f, const Eigen::Vector3f &v) { float a[3] = {v[0], v[1], v[2]}; fwrite(a, 4, 3, f); }

static bool g_rts_mode = false;
static void on_segv(int) {
  // Flyscene::raytraceScene destroys its ThreadPool twice (src/flyscene.cpp:634) and
  // crashes at scope exit, after result.ppm is complete.
  if (g_rts_mode) _exit(0);
  _exit(139);
}

int main(int argc, char **argv) {
  std::string obj, out_path, scene_path, mode = "trace", lights_arg;
  int W = 1000, H = 1000, area = 0, point = 1, stride = 1, threads = 0, off_x = 0, off_y = 0;
  float cam_rx = 0.f, cam_ry = 0.f, cam_tx = 0.f, cam_ty = 0.f, cam_tz = 0.f;
  bool primary_only = false, verbose = false, capture_hits = true;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto next = [&]() -> const char * { if (i + 1 >= argc) die("missing value"); return argv[++i]; };
    if (a == "--scene") obj = next();
    else if (a == "--w") W = atoi(next());
    else if (a == "--h") H = atoi(next());
    else if (a == "--area") area = atoi(next());
    else if (a == "--point") point = atoi(next());
    else if (a == "--stride") stride = atoi(next());
    else if (a == "--offx") off_x = atoi(next());
    else if (a == "--offy") off_y = atoi(next());
    else if (a == "--threads") threads = atoi(next());
    else if (a == "--mode") mode = next();
    else if (a == "--out") out_path = next();
    else if (a == "--dump-scene") scene_path = next();
    else if (a == "--lights") lights_arg = next();  // "x,y,z;x,y,z;..."
    else if (a == "--cam-rot") { cam_rx = atof(next()); cam_ry = atof(next()); }
    else if (a == "--cam-trans") { cam_tx = atof(next()); cam_ty = atof(next()); cam_tz = atof(next()); }
    else if (a == "--primary-only") primary_only = true;
    else if (a == "--no-capture") capture_hits = false;  // timing runs: skip the untimed face / t pass
    else if (a == "--verbose") verbose = true;
#ifdef RT_PATCHED
    else if (a == "--max-depth") rt_oracle_max_depth = atoi(next());
    else if (a == "--grid") { rt_oracle_usteps = atoi(next()); rt_oracle_vsteps = atoi(next()); }
#endif
    else die(("unknown argument " + a).c_str());
  }
  if (obj.empty()) die("--scene <file.obj> required");
  if (!out_path.empty()) out_path = (out_path[0] == '/') ? out_path : abspath(".") + "/" + out_path;
  if (!scene_path.empty()) scene_path = (scene_path[0] == '/') ? scene_path : abspath(".") + "/" + scene_path;
  std::string scratch = make_scratch_cwd(abspath(obj));
  if (chdir(scratch.c_str()) != 0) die("chdir failed");

  // stdin answers for Flyscene::initialize (src/flyscene.cpp:31-34) and a quiet stdout
  std::istringstream fake_in(std::to_string(area) + " " + std::to_string(point) + "\n");
  std::streambuf *old_in = std::cin.rdbuf(fake_in.rdbuf());
  std::ofstream devnull("/dev/null");
  std::streambuf *old_out = std::cout.rdbuf();
  if (!verbose) std::cout.rdbuf(devnull.rdbuf());

  signal(SIGSEGV, on_segv);

  Flyscene *fs = new Flyscene();
  auto t0 = std::chrono::high_resolution_clock::now();
  fs->initialize(W, H);
  double init_s = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
  std::cin.rdbuf(old_in);

  // octree build time alone (src/flyscene.cpp:91-102 times exactly this call)
  double build_s = 0.0;
  {
    auto tb = std::chrono::high_resolution_clock::now();
    BoxTree probe(fs->getMesh(), 1000);
    build_s = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - tb).count();
  }

  if (!lights_arg.empty()) {
    fs->lights.clear();
    std::stringstream ss(lights_arg);
    std::string item;
    while (std::getline(ss, item, ';')) {
      float x, y, z;
      if (sscanf(item.c_str(), "%f,%f,%f", &x, &y, &z) != 3) die("bad --lights");
      fs->lights.push_back(Eigen::Vector3f(x, y, z));
    }
  }
  if (cam_rx != 0.f || cam_ry != 0.f || cam_tx != 0.f || cam_ty != 0.f || cam_tz != 0.f) {
    fs->flycamera.rotation_X_axis = cam_rx;
    fs->flycamera.rotation_Y_axis = cam_ry;
    fs->flycamera.translation_vector = Eigen::Vector3f(cam_tx, cam_ty, cam_tz);
    fs->flycamera.updateViewMatrix();
  }

  Tucano::Mesh &mesh = fs->getMesh();
  const int T = mesh.getNumberOfFaces();

  if (!scene_path.empty()) {
    FILE *f = fopen(scene_path.c_str(), "wb");
    if (!f) die("cannot open scene dump");
    const int M = (int)fs->materials.size();
    const int L = (int)fs->lights.size();
    const int NV = (int)mesh.getNumberOfVertices();
    fwrite("RTSC", 1, 4, f);
    wr<int32_t>(f, 2);  // version
    wr<int32_t>(f, T); wr<int32_t>(f, M); wr<int32_t>(f, L); wr<int32_t>(f, NV);
    wr<int32_t>(f, W); wr<int32_t>(f, H); wr<int32_t>(f, area); wr<int32_t>(f, point);
    // camera (the inputs of screenToWorld / getCenter)
    wr3(f, fs->flycamera.getCenter());
    Eigen::Affine3f vinv = fs->flycamera.getViewMatrix().inverse();
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) wr<float>(f, vinv.matrix()(r, c));
    Eigen::Affine3f view = fs->flycamera.getViewMatrix();
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) wr<float>(f, view.matrix()(r, c));
    Eigen::Vector4f vp = fs->flycamera.getViewport();
    for (int k = 0; k < 4; ++k) wr<float>(f, vp[k]);
    wr<float>(f, fs->flycamera.fovy);
    wr<float>(f, fs->flycamera.aspect_ratio);
    wr<float>(f, fs->flycamera.getPerspectiveScale());
    // light model
    Eigen::Vector4f lc = fs->lightrep.getColor();
    wr<float>(f, lc[0]); wr<float>(f, lc[1]); wr<float>(f, lc[2]);
    wr<float>(f, fs->lightrep.getBoundingSphereRadius());
    // root box + mesh normalisation data
    wr3(f, fs->octree.box.getMin()); wr3(f, fs->octree.box.getMax());
    wr3(f, mesh.getCentroid());
    wr<float>(f, mesh.getBoundingSphereRadius());
    wr<float>(f, mesh.getNormalizationScale());
    // model matrix applied to normals in phongShade (src/flyscene.cpp:829)
    Eigen::Affine3f mm = mesh.getModelMatrix();
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) wr<float>(f, mm.matrix()(r, c));
    for (int l = 0; l < L; ++l) wr3(f, fs->lights[l]);
    for (int m = 0; m < M; ++m) {
      Tucano::Material::Mtl &mt = fs->materials[m];
      wr3(f, mt.getDiffuse()); wr3(f, mt.getSpecular());
      wr<float>(f, mt.getShininess()); wr<float>(f, mt.getOpticalDensity());
      wr<int32_t>(f, (int)mt.getIlluminationModel());
    }
    // per face: the exact values rayTriangleIntersection / getInterpolatedNormal read
    Eigen::Affine3f sm = mesh.getShapeModelMatrix();
    for (int i = 0; i < T; ++i) {
      Tucano::Face &face = mesh.getFace(i);
      for (int k = 0; k < 3; ++k) {
        Eigen::Vector3f v = (sm * mesh.getVertex(face.vertex_ids[k])).head<3>();
        wr3(f, v);
      }
    }
    for (int i = 0; i < T; ++i) wr3(f, mesh.getFace(i).normal);
    for (int i = 0; i < T; ++i)
      for (int k = 0; k < 3; ++k) wr3(f, mesh.getNormal(mesh.getFace(i).vertex_ids[k]));
    for (int i = 0; i < T; ++i) wr<int32_t>(f, mesh.getFace(i).material_id);
    for (int i = 0; i < T; ++i)
      for (int k = 0; k < 3; ++k) wr<int32_t>(f, (int32_t)mesh.getFace(i).vertex_ids[k]);
    // raw object-space vertices + stored vertex normals (for loader parity tests)
    for (int v = 0; v < NV; ++v) { Eigen::Vector4f p = mesh.getVertex(v); wr<float>(f, p[0]); wr<float>(f, p[1]); wr<float>(f, p[2]); }
    // octree statistics (src/boxTree.cpp): leaves, inner nodes, face references
    {
      long leaves = 0, inner = 0, refs = 0, maxleaf = 0;
      std::vector<const BoxTree *> st; st.push_back(&fs->octree);
      while (!st.empty()) {
        const BoxTree *n = st.back(); st.pop_back();
        if (n->isLeaf && !n->isEmpty) { leaves++; refs += (long)n->faces.size(); if ((long)n->faces.size() > maxleaf) maxleaf = (long)n->faces.size(); }
        else if (!n->isEmpty) { inner++; for (const BoxTree &c : n->children) st.push_back(&c); }
      }
      wr<int64_t>(f, leaves); wr<int64_t>(f, inner); wr<int64_t>(f, refs); wr<int64_t>(f, maxleaf);
    }
    fclose(f);
  }

  if (mode == "rts") {
    // the reference's own frame driver; writes result.ppm into the scratch CWD
    g_rts_mode = true;
    std::cout.rdbuf(old_out);
    fprintf(stderr, "scratch_cwd %s\n", scratch.c_str());
    fs->raytraceScene();
    _exit(0);
  }

  if (out_path.empty()) { std::cout.rdbuf(old_out); return 0; }

  // sampled pixel list
  std::vector<int> px, py;
  for (int i = off_x; i < W; i += stride)
    for (int j = off_y; j < H; j += stride) { px.push_back(i); py.push_back(j); }
  const size_t N = px.size();
  std::vector<float> rgb(N * 3), tt(N);
  std::vector<int32_t> fid(N);

  Eigen::Vector3f origin = fs->flycamera.getCenter();
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency() - 1;  // src/flyscene.cpp:558
  if (threads < 1) threads = 1;

  // Pass 1 (TIMED): exactly the reference's per-pixel work -- screenToWorld, the root-box cull and traceRay
  // (src/flyscene.cpp:573-598, 613-625) -- and nothing else.
  // Pass 2 (untimed, skipped by --no-capture): the primary hit (face id, t) of the same pixels, recomputed with
  // the same public calls as src/flyscene.cpp:672-683.  It used to run inside the timed loop, where it added one
  // nearest-hit query per hit pixel to the reference's time.
  auto run_pool = [&](const std::function<void(size_t)> &body) {
    std::atomic<size_t> cursor(0);
    auto worker = [&]() {
      for (;;) {
        size_t k0 = cursor.fetch_add(64);
        if (k0 >= N) break;
        size_t k1 = std::min(N, k0 + 64);
        for (size_t k = k0; k < k1; ++k) body(k);
      }
    };
    std::vector<std::thread> pool;
    for (int i = 0; i < threads; ++i) pool.emplace_back(worker);
    for (auto &th : pool) th.join();
  };
  auto t1 = std::chrono::high_resolution_clock::now();
  run_pool([&](size_t k) {
    Eigen::Vector3f o = origin;
    Eigen::Vector3f screen = fs->flycamera.screenToWorld(Eigen::Vector2f(px[k], py[k]));
    bool hitBox = fs->octree.box.boxIntersect(o, screen);  // src/flyscene.cpp:576
    Eigen::Vector3f colour(1.f, 1.f, 1.f);                 // BACKGROUND, src/flyscene.cpp:12
    if (hitBox && !primary_only) {
      Eigen::Vector3f direction = screen - o;  // src/flyscene.cpp:619 (not normalised)
      colour = fs->traceRay(o, direction, 0, fs->lights, false);
    }
    rgb[3 * k + 0] = colour[0]; rgb[3 * k + 1] = colour[1]; rgb[3 * k + 2] = colour[2];
  });
  double render_s = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t1).count();
  auto t2 = std::chrono::high_resolution_clock::now();
  for (size_t k = 0; k < N; ++k) { fid[k] = -1; tt[k] = std::numeric_limits<float>::max(); }
  if (capture_hits) {
    run_pool([&](size_t k) {
      Eigen::Vector3f o = origin;
      Eigen::Vector3f screen = fs->flycamera.screenToWorld(Eigen::Vector2f(px[k], py[k]));
      if (!fs->octree.box.boxIntersect(o, screen)) return;
      Eigen::Vector3f direction = screen - o;
      int best = -1;
      float t = std::numeric_limits<float>::max();
      std::set<int> faces = fs->octree.intersect(o, direction + o);
      for (int fi : faces) {
        Tucano::Face tri = mesh.getFace(fi);
        float is = fs->rayTriangleIntersection(o, direction, tri);
        if (is != -72 && is < t && is > 0.00001f) { t = is; best = fi; }
      }
      fid[k] = best; tt[k] = t;
    });
  }
  double capture_s = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t2).count();

  FILE *f = fopen(out_path.c_str(), "wb");
  if (!f) die("cannot open --out");
  fwrite("RTOR", 1, 4, f);
  wr<int32_t>(f, 1);
  wr<int32_t>(f, W); wr<int32_t>(f, H); wr<int32_t>(f, stride); wr<int32_t>(f, threads);
  wr<int64_t>(f, (int64_t)N);
  wr<double>(f, render_s); wr<double>(f, build_s); wr<double>(f, init_s);
  std::vector<int32_t> pxy(N * 2);
  for (size_t k = 0; k < N; ++k) { pxy[2 * k] = px[k]; pxy[2 * k + 1] = py[k]; }
  fwrite(pxy.data(), 4, N * 2, f);
  fwrite(rgb.data(), 4, N * 3, f);
  fwrite(fid.data(), 4, N, f);
  fwrite(tt.data(), 4, N, f);
  fclose(f);

  std::cout.rdbuf(old_out);
  printf("{\"pixels\": %zu, \"threads\": %d, \"render_s\": %.6f, \"capture_s\": %.6f, \"octree_build_s\": %.6f, \"init_s\": %.6f, \"faces\": %d}\n",
         N, threads, render_s, capture_s, build_s, init_s, T);
  fflush(stdout);
  _exit(0);  // skip GL-object destructors
}
