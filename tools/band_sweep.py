#!/usr/bin/env python
"""Developer tool: what one rank of an N-way band split costs as a function of the band height.
Renders rank 0's share (band_world = N) of a bench workload on ONE GPU for several band_rows and prints
the frame time (CUDA events around graph replays, L2 flushed) -- the per-rank compute of the N-GPU run
without the peer stores.
    python tools/band_sweep.py --workload c5 --world 8 --rows 4 8 16 32 64 128"""
import argparse
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c5")
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--rows", type=int, nargs="+", default=[4, 8, 16, 32, 64, 128])
    ap.add_argument("--frames", type=int, default=20)
    a = ap.parse_args()
    capi = importlib.import_module("raytracer-in-cpp_b200").capi
    capi.init(0)
    wl = bench.WORKLOADS[a.workload]
    arrs, spheres, smat = bench.workload_arrays(wl)
    W, H = wl["w"], wl["h"]
    cam = capi.default_camera(W, H)
    lights = capi.Lights(np.array([[-1.0, 1.0, 1.0]], np.float32))
    scene = capi.Scene(*arrs, None, spheres, smat)
    stream = torch.cuda.Stream()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for world in (1, a.world):
        for br in (a.rows if world > 1 else [8]):
            params = capi.make_params(W, H, wl["area"], wl["point"], wl["max_depth"], wl["grid"], band_rows=br,
                                      band_rank=0, band_world=world)
            rows = capi.lib().rt_local_rows(params)
            out = torch.empty((rows, W, 4), dtype=torch.uint8, device="cuda")
            ts = []
            with torch.cuda.stream(stream):
                for k in range(a.frames + 3):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    scene.render_device(cam, lights, params, out.data_ptr(), stream=stream.cuda_stream)
                    e1.record(stream)
                    stream.synchronize()
                    if k >= 3:
                        ts.append(e0.elapsed_time(e1))
            print(f"{a.workload} world {world} band_rows {br:4d}: rows {rows:5d}  {np.median(ts):.4f} ms/frame", flush=True)


if __name__ == "__main__":
    main()
