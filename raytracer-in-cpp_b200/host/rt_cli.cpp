// rt_cli.cpp -- headless replacement for the reference's GLFW shell (src/main.cpp): where the
// reference opens a 1000x1000 window and renders on key 'T' (main.cpp:8-9,69-70), this program
// initialises a Flyscene, calls raytraceScene() once and exits, leaving result.ppm in the CWD.
//
//   rt_cli [--scene file.obj] [--width W] [--height H] [--area 0|1] [--point 0|1]
//          [--max-depth D] [--grid U V] [--light x y z]... [--cam-rot rx ry] [--cam-trans x y z]
//          [--frames N] [--device k] [--gpus N]      (--gpus N: devices 0..N-1 share every frame, rt_multi_*)
// Without --area/--point the two flags are read from stdin exactly like the reference.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <string>

#include "flyscene.hpp"

int main(int argc, char **argv) {
  std::string scene = "resources/models/cube.obj";
  int W = 1000, H = 1000, area = -1, point = -1, depth = -1, gu = 5, gv = 5, frames = 1, device = 0, gpus = 1;
  float rx = 0, ry = 0, tx = 0, ty = 0, tz = 0;
  uint32_t sphere_seed = 1;
  std::vector<rt::Vector3f> extra_lights;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto next = [&]() -> const char * {
      if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); }
      return argv[++i];
    };
    if (a == "--scene") scene = next();
    else if (a == "--width") W = atoi(next());
    else if (a == "--height") H = atoi(next());
    else if (a == "--area") area = atoi(next());
    else if (a == "--point") point = atoi(next());
    else if (a == "--max-depth") depth = atoi(next());
    else if (a == "--grid") { gu = atoi(next()); gv = atoi(next()); }
    else if (a == "--light") { float x = atof(next()), y = atof(next()), z = atof(next()); extra_lights.emplace_back(x, y, z); }
    else if (a == "--cam-rot") { rx = atof(next()); ry = atof(next()); }
    else if (a == "--cam-trans") { tx = atof(next()); ty = atof(next()); tz = atof(next()); }
    else if (a == "--frames") frames = atoi(next());
    else if (a == "--device") device = atoi(next());
    else if (a == "--gpus") gpus = atoi(next());
    else if (a == "--sphere-seed") sphere_seed = (uint32_t)strtoul(next(), nullptr, 10);
    else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
  }
  try {
    if (rt_init(device) < 0) throw std::runtime_error(rt_last_error());
    rt::Flyscene fs;
    fs.setModelPath(scene);
    if (area >= 0 || point >= 0) fs.setLightMode(area > 0, point != 0);
    fs.setMaxDepth(depth);
    fs.setAreaGrid(gu, gv);
    fs.setSphereSeed(sphere_seed);
    if (gpus > 1) {
      std::vector<int> devs;
      for (int k = 0; k < gpus; ++k) devs.push_back(device + k);
      fs.setDevices(devs);
    }
    fs.initialize(W, H);
    for (const auto &l : extra_lights) fs.getLights().push_back(l);
    if (rx != 0 || ry != 0) fs.getCamera()->setRotation(rx, ry);
    if (tx != 0 || ty != 0 || tz != 0) fs.getCamera()->translate(tx, ty, tz);
    for (int f = 0; f < frames; ++f) fs.raytraceScene();
    RtStats st;
    fs.render(W, H, &st);
    printf("{\"faces\": %d, \"gpus\": %d, \"octree_build_s\": %.6f, \"frame_ms\": %.4f, \"rays\": %lld, \"mrays_per_s\": %.2f}\n",
           fs.getNumberOfFaces(), gpus, fs.octree_seconds, st.ms_total,
           (long long)(st.rays_primary + st.rays_shadow + st.rays_secondary),
           (st.rays_primary + st.rays_shadow + st.rays_secondary) / (st.ms_total * 1e3));
  } catch (const std::exception &e) {
    fprintf(stderr, "rt_cli: %s\n", e.what());
    return 1;
  }
  return 0;
}
