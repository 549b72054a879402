#!/usr/bin/env python
"""Developer tool: rt_multi_create of a bench workload on N devices (scene build on every device, in parallel).
    python tools/multi_create_time.py c3 4"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
capi = importlib.import_module("raytracer-in-cpp_b200").capi
capi.init(0)
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
arrs, sp, sm = bench.workload_arrays(wl)
for rep in range(4):
    t0 = time.perf_counter()
    m = capi.Multi(list(range(n)), *arrs, None, sp, sm)
    t1 = time.perf_counter()
    print(f"rt_multi_create on {n} devices: {1e3 * (t1 - t0):.1f} ms", flush=True)
    m.close()
