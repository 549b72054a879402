#!/usr/bin/env python
"""Developer tool: per-rank compute of an N-way band split on ONE GPU (rank 0's share), for the wavefront and the
fused frame path:  python tools/share_probe.py c2 [key=value ...]"""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
capi = importlib.import_module("raytracer-in-cpp_b200").capi
capi.init(0)
w = sys.argv[1]
for kv in sys.argv[2:]:
    k, v = kv.split("="); capi.set_option(k, int(v))
wl = bench.WORKLOADS[w]
arrs, spheres, smat = bench.workload_arrays(wl)
W, H = wl["w"], wl["h"]
cam = capi.default_camera(W, H)
lights = capi.Lights(np.array([[-1.0, 1.0, 1.0]], np.float32))
scene = capi.Scene(*arrs, None, spheres, smat)
stream = torch.cuda.Stream()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for world in (1, 2, 4, 8):
    br = bench.band_rows_for(H, world)
    for mode in (0, 1):
        capi.set_option("fused_frame", mode)
        params = capi.make_params(W, H, wl["area"], wl["point"], wl["max_depth"], wl["grid"], band_rows=br, band_rank=0, band_world=world)
        rows = capi.lib().rt_local_rows(params)
        out = torch.empty((rows, W, 4), dtype=torch.uint8, device="cuda")
        ts = []
        with torch.cuda.stream(stream):
            for k in range(43):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                scene.render_device(cam, lights, params, out.data_ptr(), stream=stream.cuda_stream)
                e1.record(stream)
                stream.synchronize()
                if k >= 3:
                    ts.append(e0.elapsed_time(e1))
        print(f"{w} world {world} rank-0 share ({rows} rows, {br}-row bands) {'fused    ' if mode else 'wavefront'}: {np.median(ts):.4f} ms", flush=True)
