#!/usr/bin/env python
"""Developer tool: fixed per-frame cost of the launch sequence (tiny 64x36 frame, same settings as C2)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
capi = importlib.import_module("raytracer-in-cpp_b200").capi
capi.init(0)
wl = bench.WORKLOADS["c2"]
arrs, sp, sm = bench.workload_arrays(wl)
scene = capi.Scene(*arrs)
for (W, H, depth) in [(64, 36, 3), (64, 36, 1), (64, 36, 0), (1920, 1080, 3), (1920, 1080, 1)]:
    cam = capi.default_camera(W, H)
    lights = capi.Lights(np.array([[-1, 1, 1]], np.float32))
    params = capi.make_params(W, H, 1, 0, depth, (4, 4))
    out = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
    ws = torch.cuda.Stream()
    torch.cuda.set_stream(ws)
    st = ws.cuda_stream
    for _ in range(20):
        scene.render_device(cam, lights, params, out.data_ptr(), stream=st)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(500):
        scene.render_device(cam, lights, params, out.data_ptr(), stream=st)
    b.record()
    cpu = (time.perf_counter() - t0) / 500 * 1e3
    torch.cuda.synchronize()
    print(f"{W}x{H} depth {depth}: gpu {a.elapsed_time(b) / 500:.4f} ms/frame, cpu submit {cpu:.4f} ms/frame")
