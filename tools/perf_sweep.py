#!/usr/bin/env python
"""Developer tool: time the frame kernels of a bench workload for several values of a runtime option.
    python tools/perf_sweep.py --workload c3 --option leaf_size --values 1 2 4
Prints per-category kernel milliseconds (CUDA events inside the library), median of --frames frames."""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--option", default="leaf_size")
    ap.add_argument("--values", type=int, nargs="+", default=[0])
    ap.add_argument("--frames", type=int, default=7)
    ap.add_argument("--recreate", action="store_true", help="re-create the scene per value (build-time options)")
    a = ap.parse_args()
    capi = importlib.import_module("raytracer-in-cpp_b200").capi
    capi.init(0)
    wl = bench.WORKLOADS[a.workload]
    arrs, spheres, smat = bench.workload_arrays(wl)
    W, H = wl["w"], wl["h"]
    cam = capi.default_camera(W, H)
    lights = capi.Lights(np.array([[-1.0, 1.0, 1.0]], np.float32))
    params = capi.make_params(W, H, wl["area"], wl["point"], wl["max_depth"], wl["grid"])
    scene = None
    for v in a.values:
        capi.set_option(a.option, v)
        if scene is None or a.recreate:
            scene = capi.Scene(*arrs, None, spheres, smat)
        rows = []
        for _ in range(a.frames + 2):
            fr = scene.render(cam, lights, params, want_face=False, want_t=False, want_rgb=False)
            rows.append((fr.stats["ms_trace"], fr.stats["ms_shadow"], fr.stats["ms_shade"], fr.stats["ms_total"]))
        m = np.median(np.array(rows[2:]), 0)
        capi.set_option("stats", 1)
        st = scene.render(cam, lights, params, want_face=False, want_t=False, want_rgb=False).stats
        capi.set_option("stats", 0)
        print("   work:", {k: st[k] for k in ("box_tests", "tri_tests", "box_tests_shadow", "tri_tests_shadow", "filter_checks",
                                             "filter_slow", "filter_rejects", "rays_shadow")})
        print(f"{a.workload} {a.option}={v}: trace {m[0]:.3f} shadow {m[1]:.3f} shade {m[2]:.3f} total {m[3]:.3f} ms  "
              f"bvh {scene.info()['nodes']} nodes", flush=True)


if __name__ == "__main__":
    main()
