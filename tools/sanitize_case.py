#!/usr/bin/env python
"""Small render used under compute-sanitizer (memcheck): gallery scene with every material branch,
area light, depth 3, plus spheres and band sharding; exercises K1/K2/K3/K3b and the batched entry points."""
import importlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("raytracer-in-cpp_b200")
capi = pkg.capi
capi.init(0)
d = tempfile.mkdtemp()
obj = os.path.join(d, "g.obj")
pkg.scenes.write_gallery(obj, 3)
mesh = capi.Mesh(obj)
arrs = list(mesh.arrays())
spheres = pkg.scenes.sphere_cloud(20)
spheres[:, 3] *= 4
scene = capi.Scene(*arrs, None, spheres, np.zeros(len(spheres), np.int32) + 3)
W, H = 160, 90
cam = capi.default_camera(W, H)
lights = capi.Lights(np.array([[-1, 1.2, 1.5], [1.5, 1, 1]], np.float32))
for kw in (dict(area=1, point=0, max_depth=3, grid=(3, 3)), dict(area=0, point=1, max_depth=-1),
           dict(area=1, point=0, max_depth=2, grid=(2, 2), band_rows=8, band_rank=1, band_world=3)):
    fr = scene.render(cam, lights, capi.make_params(W, H, **kw))
    print(kw, fr.rgba.shape, fr.stats["rays_shadow"], fr.stats["rays_secondary"], fr.stats["levels"])
o = np.random.default_rng(0).uniform(-1, 1, (500, 3)).astype(np.float32)
print(scene.trace_rays(o, -o, lights, capi.make_params(8, 8, 1, 0, 2, (2, 2)))[0].mean())
print(scene.light_strikes(o, lights).mean(), scene.box_intersect(o, -o).mean(), len(scene.octree_candidates((0, 0, 2), (0, 0, 1))))
print("SANITIZE_CASE_OK")
