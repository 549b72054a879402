#!/usr/bin/env python
"""Developer tool: time-to-first-frame pieces for the 1 M-triangle scene:
OBJ/MTL bake, BVH + reference-octree build + upload (rt_scene_create), first frame; and, when the
headless reference binary is present, the reference's own loader + octree build."""
import importlib, json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("raytracer-in-cpp_b200")
capi = pkg.capi
capi.init(0)
out = {}
for name in sys.argv[1:] or ["c3"]:
    wl = bench.WORKLOADS[name]
    obj = bench.workload_obj(wl)
    t0 = time.perf_counter(); mesh = capi.Mesh(obj); t1 = time.perf_counter()
    arrs = mesh.arrays(); t2 = time.perf_counter()
    scene = capi.Scene(*arrs); t3 = time.perf_counter()
    W, H = wl["w"], wl["h"]
    fr = scene.render(capi.default_camera(W, H), capi.Lights(np.array([[-1, 1, 1]], np.float32)),
                      capi.make_params(W, H, wl["area"], wl["point"], wl["max_depth"], wl["grid"]), want_face=False, want_t=False, want_rgb=False)
    t4 = time.perf_counter()
    rec = {"faces": int(arrs[0].shape[0]), "load_obj_s": t1 - t0, "scene_create_s": t3 - t2, "scene_info": scene.info(),
           "first_frame_s": t4 - t3, "build_info_first_call": scene.build_info()}
    # the same scene again (context, allocator and kernels warm), and with the host builders
    t5 = time.perf_counter(); s2 = capi.Scene(*arrs); t6 = time.perf_counter()
    rec["scene_create_warm_s"] = t6 - t5
    rec["build_info_warm"] = s2.build_info()
    s2.close()
    capi.set_option("gpu_build", 0)
    t7 = time.perf_counter(); s3 = capi.Scene(*arrs); t8 = time.perf_counter()
    capi.set_option("gpu_build", 2)
    rec["scene_create_host_builders_s"] = t8 - t7
    rec["build_info_host_builders"] = s3.build_info()
    s3.close()
    from oracle import oracle as O
    if O.have_ref():
        tmp = os.path.join(tempfile.gettempdir(), "bt.bin")
        js = json.loads(O.run_ref(obj, tmp, W, H, wl["area"], wl["point"], stride=max(W, H), primary_only=True).strip().splitlines()[-1])
        rec["reference"] = {"initialize_s": js["init_s"], "octree_build_s": js["octree_build_s"]}
    out[name] = rec
print(json.dumps(out, indent=1, default=float))
