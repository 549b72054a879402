#!/usr/bin/env python
"""Developer tool: render a few frames of a bench workload (for ncu captures).
    python tools/one_frame.py c4 [frames] [key=value ...]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
capi = importlib.import_module("raytracer-in-cpp_b200").capi
capi.init(0)
w = sys.argv[1]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 4
for kv in sys.argv[3:]:
    k, v = kv.split("="); capi.set_option(k, int(v))
wl = bench.WORKLOADS[w]
arrs, sp, sm = bench.workload_arrays(wl)
scene = capi.Scene(*arrs, None, sp, sm)
W, H = wl["w"], wl["h"]
cam = capi.default_camera(W, H)
lights = capi.Lights(np.array([[-1, 1, 1]], np.float32))
params = capi.make_params(W, H, wl["area"], wl["point"], wl["max_depth"], wl["grid"])
for _ in range(frames):
    fr = scene.render(cam, lights, params, want_face=False, want_t=False, want_rgb=False)
print(w, fr.stats["ms_total"], "ms")
