// facade_probe.cpp -- exercises every render-path member of the C++ facade (host/flyscene.hpp) once and
// prints the results as JSON; tests/test_gpu_cli.py compares them with the oracle.  Mirrors the calls the
// reference's own debug-ray tool makes (src/flyscene.cpp:241-300: screenToWorld, boxIntersect,
// octree.intersect, rayTriangleIntersection, lightStrikes, traceRay).
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "flyscene.hpp"

static void pv(const char *k, rt::Vector3f v, bool comma = true) {
  printf("\"%s\": [%.9g, %.9g, %.9g]%s\n", k, v.x, v.y, v.z, comma ? "," : "");
}

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: facade_probe scene.obj [px py]\n"); return 2; }
  const float px = argc > 3 ? atof(argv[2]) : 300.f, py = argc > 3 ? atof(argv[3]) : 260.f;
  try {
    if (rt_init(0) < 0) throw std::runtime_error(rt_last_error());
    rt::Flyscene fs;
    fs.setModelPath(argv[1]);
    fs.setLightMode(/*area=*/true, /*point=*/false);
    fs.setMaxDepth(2);
    fs.initialize(640, 480);
    fs.getLights().push_back(rt::Vector3f(1.5f, 1.0f, 1.0f));
    rt::Flycamera *cam = fs.getCamera();
    rt::Vector3f origin = cam->getCenter();
    rt::Vector3f screen = cam->screenToWorld(px, py);
    rt::Vector3f dir = screen - origin;
    printf("{\n");
    pv("origin", origin);
    pv("screen", screen);
    printf("\"faces\": %d,\n", fs.getNumberOfFaces());
    printf("\"root_hit\": %d,\n", (int)fs.octree.box.boxIntersect(origin, screen));
    pv("root_min", fs.octree.box.getMin());
    pv("root_max", fs.octree.box.getMax());
    std::set<int> cand = fs.octree.intersect(origin, dir + origin);
    printf("\"candidates\": %zu,\n\"candidate_sum\": %lld,\n", cand.size(), [&] { long long s = 0; for (int c : cand) s += c; return s; }());
    // nearest hit over the candidates, exactly the loop of src/flyscene.cpp:675-683
    float t = 3.402823466e+38f;
    int best = -1;
    for (int f : cand) {
      rt::Face face = fs.getFace(f);
      float is = fs.rayTriangleIntersection(origin, dir, face);
      if (is != -72 && is < t && is > 0.00001f) { t = is; best = f; }
    }
    printf("\"best\": %d,\n\"t\": %.9g,\n", best, t);
    rt::Vector3f colour = fs.traceRay(origin, dir, 0, fs.getLights(), false);
    pv("colour", colour);
    if (best >= 0) {
      rt::Vector3f hit = origin + t * dir;
      bool vis[25];
      bool any = fs.lightStrikes(hit, fs.getLights(), vis);
      printf("\"light_any\": %d,\n\"light_vis\": [%d, %d],\n", (int)any, (int)vis[0], (int)vis[1]);
      rt::Face face = fs.getFace(best);
      pv("phong", fs.phongShade(origin, hit, face, fs.getLights()));
    }
    std::vector<rt::Vector3f> samples = fs.createSpherePoint(fs.getLights()[0]);
    printf("\"n_samples\": %zu,\n", samples.size());
    pv("sample0", samples[0]);
    pv("sample24", samples.back());
    rt::BoundingBox bb(rt::Vector3f(-0.25f, -0.25f, -0.25f), rt::Vector3f(0.25f, 0.25f, 0.25f));
    printf("\"bb_hit\": %d,\n\"bb_miss\": %d,\n", (int)bb.boxIntersect(origin, rt::Vector3f(0, 0, 0)),
           (int)bb.boxIntersect(origin, rt::Vector3f(3, 3, 0)));
    // the headless debug ray (recursiveDebugRay, src/flyscene.cpp:241-430): chain of mirror bounces
    std::vector<rt::DebugRayLevel> dbg = fs.debugRay(px, py, 3, false);
    printf("\"debug_ray\": [");
    for (size_t k = 0; k < dbg.size(); ++k) {
      const rt::DebugRayLevel &lv = dbg[k];
      printf("%s{\"level\": %d, \"face\": %d, \"t\": %.9g, \"origin\": [%.9g, %.9g, %.9g], \"direction\": [%.9g, %.9g, %.9g], "
             "\"colour\": [%.9g, %.9g, %.9g], \"visible\": [%d, %d]}",
             k ? ", " : "", lv.level, lv.face, lv.t, lv.origin.x, lv.origin.y, lv.origin.z, lv.direction.x, lv.direction.y,
             lv.direction.z, lv.colour.x, lv.colour.y, lv.colour.z, lv.face >= 0 ? (int)lv.light_visible[0] : -1,
             lv.face >= 0 ? (int)lv.light_visible[1] : -1);
    }
    printf("],\n");
    std::vector<int64_t> st = fs.octree.stats();
    printf("\"octree\": [%lld, %lld, %lld, %lld]\n}\n", (long long)st[0], (long long)st[1], (long long)st[2], (long long)st[3]);
  } catch (const std::exception &e) {
    fprintf(stderr, "facade_probe: %s\n", e.what());
    return 1;
  }
  return 0;
}
