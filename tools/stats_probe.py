#!/usr/bin/env python
"""Developer tool: one stats frame per workload and option set -- work counters, phase shares, ms."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
capi = importlib.import_module("raytracer-in-cpp_b200").capi
capi.init(0)
wls = sys.argv[1].split(",")
optsets = sys.argv[2:] or [""]
for w in wls:
    wl = bench.WORKLOADS[w]
    arrs, sp, sm = bench.workload_arrays(wl)
    scene = None
    W, H = wl["w"], wl["h"]
    cam = capi.default_camera(W, H)
    lights = capi.Lights(np.array([[-1, 1, 1]], np.float32))
    params = capi.make_params(W, H, wl["area"], wl["point"], wl["max_depth"], wl["grid"])
    for o in optsets:
        for kv in o.split():
            k, v = kv.split("="); capi.set_option(k, int(v))
        if scene is None or "build" in o or "leaf" in o:   # options that act at scene creation
            scene = capi.Scene(*arrs, None, sp, sm)
            print(w, f"[{o}] build", scene.build_info(), flush=True)
        capi.set_option("stats", 1)
        for _ in range(3):
            fr = scene.render(cam, lights, params, want_face=False, want_t=False, want_rgb=False)
        capi.set_option("stats", 0)
        ms = []
        for _ in range(5):
            f2 = scene.render(cam, lights, params, want_face=False, want_t=False, want_rgb=False)
            ms.append(f2.stats["ms_total"])
        s = fr.stats
        print(f"{w} [{o}] ms {np.median(ms):.4f} (stats build {s['ms_total']:.4f}: trace {s['ms_trace']:.4f} shadow {s['ms_shadow']:.4f} shade {s['ms_shade']:.4f}) "
              f"box {s['box_tests']} tri {s['tri_tests']} box_s {s['box_tests_shadow']} tri_s {s['tri_tests_shadow']} samples {s['shade_samples']} "
              f"traced {s['shadow_rays_traced']} shadow {s['rays_shadow']} sec {s['rays_secondary']} levels {s['levels']} launches {s['kernel_launches']}", flush=True)
