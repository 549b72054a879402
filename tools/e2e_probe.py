#!/usr/bin/env python
"""Developer tool: blocking rt_render into a pinned host frame vs the device-resident frame, for several depth caps
of the C2 scene (how much of the end-to-end time is the host frame?).   python tools/e2e_probe.py [key=value ...]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
capi = importlib.import_module("raytracer-in-cpp_b200").capi
capi.init(0)
for kv in sys.argv[1:]:
    k, v = kv.split("="); capi.set_option(k, int(v))
wl = bench.WORKLOADS["c2"]
arrs, sp, sm = bench.workload_arrays(wl)
scene = capi.Scene(*arrs)
W, H = 1920, 1080
cam = capi.default_camera(W, H)
lights = capi.Lights(np.array([[-1, 1, 1]], np.float32))
host = torch.zeros((H, W, 4), dtype=torch.uint8).pin_memory().numpy()
dev = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
ws = torch.cuda.Stream(); torch.cuda.set_stream(ws)
for depth in (0, 1, 3):
    params = capi.make_params(W, H, 1, 0, depth, (4, 4))
    for _ in range(20):
        scene.render(cam, lights, params, want_face=False, want_t=False, want_rgb=False, want_stats=False, out_rgba=host)
    t0 = time.perf_counter()
    for _ in range(200):
        scene.render(cam, lights, params, want_face=False, want_t=False, want_rgb=False, want_stats=False, out_rgba=host)
    e2e = (time.perf_counter() - t0) / 200
    for _ in range(20):
        scene.render_device(cam, lights, params, dev.data_ptr(), stream=ws.cuda_stream)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200):
        scene.render_device(cam, lights, params, dev.data_ptr(), stream=ws.cuda_stream)
    b.record(); torch.cuda.synchronize()
    print(f"depth {depth}: host frame {1e3 * e2e:.4f} ms, device frame {a.elapsed_time(b) / 200:.4f} ms (no L2 flush)", flush=True)
