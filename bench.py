#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 render path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c4|c5] [--impl ours|reference]

Metric (BASELINE.json): Mrays/s and ms/frame at 1080p.  One "step" = one frame of the workload:
  c2 (default, the configuration the metric is quoted on): the bundled cube scene at 1920x1080 with
     the AreaLight grid at 4x4 samples and reflection depth 3 (BASELINE.json configs[1]).
  c1: cube 1000x1000 point light, natural depth          (configs[0])
  c3: synthetic 999 698-triangle height field, 1080p, point light, depth 0   (configs[2])
  c4: 100 352 triangles + 1000 analytic spheres, 3840x2160, 4x4 samples, depth 5  (configs[3])
A ray = one nearest-hit or any-hit query; counts come from the kernels' own
queue sizes.  `value` is measured with the scene and the framebuffer resident in HBM (CUDA events on
the launch stream, L2 flushed between frames); `e2e` goes through rt_render() with pinned HOST
buffers, copies inside the timed region.

N > 1 (torchrun, one rank per GPU): the frame is split into interleaved row bands (band b -> rank b mod N), every
rank renders its bands and its kernels store the pixels straight into rank 0's frame over NVLink peer memory; per-rank
flags behind the frame signal completion (no collective in the timed step).  The NCCL gather of per-rank buffers is
timed beside it as the baseline (`nccl_gather_baseline`), and so is the same frame driven from ONE process through
rt_multi_* (`single_process_multi_gpu`).  e2e at N > 1: every rank copies its own bands into a shared pinned host
frame.  Strong scaling: the frame is fixed.

--impl reference: the reference's own CPU implementation (oracle/_ref/ref_oracle[_patched], built
from /root/reference by oracle/Makefile; falls back to the C port if that binary is absent) on a
bounded pixel subset of the same workload, with the reference's own thread count
(hardware_concurrency - 1, src/flyscene.cpp:558).
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Mrays/s at 1080p (primary + shadow + secondary rays per second, reference ray census)"
BAND_ROWS = 8  # minimum band height; see band_rows_for()


def band_rows_for(height, world):
    """Interleaved band height: every rank gets about 8 bands (load balance across the image) that are as
    tall as that allows (a rank then touches ~1/world of the scene instead of all of it; measured with
    tools/band_sweep.py: 8K frame on 8 ranks 0.84 -> 0.76 ms per rank going from 8- to 64-row bands)."""
    return max(BAND_ROWS, (height // (max(world, 1) * 8)) & ~7)

WORKLOADS = {
    "c1": dict(desc="bundled cube.obj 1000x1000, point light, natural depth (BASELINE configs[0])",
               scene=("golden", "cube_point_1000"), w=1000, h=1000, area=0, point=1, max_depth=-1, grid=(5, 5)),
    "c2": dict(desc="bundled cube.obj 1920x1080, AreaLight 4x4 soft shadows, reflection depth 3 (BASELINE configs[1])",
               scene=("golden", "cube_point_1000"), w=1920, h=1080, area=1, point=0, max_depth=3, grid=(4, 4)),
    "c3": dict(desc="synthetic height field 999698 triangles 1920x1080, point light, primary+shadow (BASELINE configs[2])",
               scene=("gen", "write_heightfield", (707,)), w=1920, h=1080, area=0, point=1, max_depth=0, grid=(5, 5)),
    "c4": dict(desc="synthetic 100352 triangles + 1000 spheres 3840x2160, AreaLight 4x4, depth 5 (BASELINE configs[3])",
               scene=("gen", "write_heightfield", (224,)), spheres=1000, w=3840, h=2160, area=1, point=0, max_depth=5,
               grid=(4, 4)),
    "c5": dict(desc="synthetic height field 999698 triangles 7680x4320 (8K), point light, primary+shadow (BASELINE configs[4])",
               scene=("gen", "write_heightfield", (707,)), w=7680, h=4320, area=0, point=1, max_depth=0, grid=(5, 5)),
}


def pkg():
    return importlib.import_module("raytracer-in-cpp_b200")


def scene_cache_dir():
    d = os.path.join(tempfile.gettempdir(), "rt_b200_scenes")
    os.makedirs(d, exist_ok=True)
    return d


def workload_obj(wl):
    """Path of the OBJ for generated workloads (deterministic generators), None for golden scenes."""
    kind = wl["scene"][0]
    if kind != "gen":
        return None
    gen, args = wl["scene"][1], wl["scene"][2]
    path = os.path.join(scene_cache_dir(), f"{gen}_{'_'.join(map(str, args))}.obj")
    if not os.path.exists(path):
        # each rank may race here: write under a private name, then rename the pair into place
        tmp = os.path.join(scene_cache_dir(), f"tmp{os.getpid()}")
        os.makedirs(tmp, exist_ok=True)
        tpath = os.path.join(tmp, os.path.basename(path))
        getattr(pkg().scenes, gen)(tpath, *args)
        os.replace(os.path.splitext(tpath)[0] + ".mtl", os.path.splitext(path)[0] + ".mtl")
        os.replace(tpath, path)
    return path


def workload_arrays(wl):
    """Baked scene arrays (+ optional spheres) for a workload."""
    capi = pkg().capi
    if wl["scene"][0] == "golden":
        g = np.load(os.path.join(ROOT, "tests", "golden", wl["scene"][1] + ".npz"))
        arrs = (g["verts"], g["fnormals"], g["vnormals"], g["mat_id"], g["mats"])
    else:
        mesh = capi.Mesh(workload_obj(wl))
        arrs = mesh.arrays()
        mesh.close()
    spheres = sphere_mat = None
    if wl.get("spheres"):
        spheres = pkg().scenes.sphere_cloud(wl["spheres"])
        mats = np.concatenate([arrs[4], np.array([[0.8, 0.3, 0.2, 1, 1, 1, 30, 0, 2], [0.9, 0.9, 0.9, 1, 1, 1, 60, 0, 4]],
                                                 np.float32)], 0)
        base = arrs[4].shape[0]
        sphere_mat = (base + (np.arange(len(spheres)) % 2)).astype(np.int32)
        arrs = arrs[:4] + (mats,)
    return arrs, spheres, sphere_mat


# ----------------------------------------------------------------------------------------------
# clocks sampler (NVML), runs during the timed region
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
                 0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, nm in names.items():
                    if r & bit and nm != "gpu_idle":
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=1.0)
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------
# reference arm (CPU)
# ----------------------------------------------------------------------------------------------
def reference_sample(wl, stride, threads=0):
    """One bounded sample of the workload on the reference's CPU path.  Returns
    (seconds, pixels, kind, threads_used)."""
    from oracle import oracle as O
    w, h = wl["w"], wl["h"]
    patched = wl["max_depth"] >= 0 or tuple(wl["grid"]) != (5, 5)
    use_ref = O.have_ref() and not wl.get("spheres")
    if use_ref:
        obj = os.path.join(O.REF_SCENES, "cube.obj") if wl["scene"][0] == "golden" else workload_obj(wl)
        out = os.path.join(tempfile.gettempdir(), f"rt_ref_sample_{os.getpid()}.bin")
        O.run_ref(obj, out, w, h, wl["area"], wl["point"], stride, threads=threads,
                  max_depth=(wl["max_depth"] if patched and wl["max_depth"] >= 0 else (1 << 30 if patched else None)),
                  grid=(tuple(wl["grid"]) if patched else None),
                  capture=False)  # timed: only the reference's own per-pixel work (no face / t capture pass)
        r = O.load_render_dump(out)
        os.remove(out)
        return r.render_s, len(r.face), "reference", r.threads
    # C port (kind "port"): used when the reference binary is absent, or for spheres (no reference implementation)
    arrs, spheres, sphere_mat = workload_arrays(wl)
    baked = O.BakedScene(*arrs, spheres=spheres, sphere_mat=sphere_mat)
    orc = O.Oracle(baked, area=wl["area"], point=wl["point"], max_depth=wl["max_depth"], grid=wl["grid"])
    cam = O.Oracle.camera((0, 0, 2), np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 2]], np.float32), (0, 0, w, h),
                          60.0, np.float32(w) / np.float32(h))
    nthr = threads if threads > 0 else max(1, (os.cpu_count() or 2) - 1)
    t0 = time.perf_counter()
    pxy, *_ = orc.render(cam, np.array([[-1, 1, 1]], np.float32), w, h, stride=stride, threads=nthr)
    return time.perf_counter() - t0, len(pxy), "port", nthr


def reference_raytrace_scene_c1():
    """BASELINE configs[0] through the reference's OWN frame driver: Flyscene::raytraceScene() on the bundled cube at
    the reference resolution (1000x1000, point light) -- pixel pre-pass, ThreadPool of hardware_concurrency()-1
    workers, traceRay per pixel, ASCII result.ppm -- timed by the reference's own clock (its "ELAPSED TIME" line,
    src/flyscene.cpp:521,642-647).  The reference cannot run this driver on non-square images, so
    the 1080p workloads are timed through the direct traceRay loop instead; this leg shows what its own driver adds."""
    from oracle import oracle as O
    if not O.have_ref():
        return None
    try:
        s, thr, _ = O.run_ref_raytrace_scene(os.path.join(O.REF_SCENES, "cube.obj"), 1000, 0, 1)
    except Exception as ex:
        return {"unavailable": str(ex)[:200]}
    rays = 1000000 + 2 * 494209  # App. A.8: 494 209 hit pixels x (merged gate/sample ray + mirror child) + primaries
    return {"workload": WORKLOADS["c1"]["desc"], "call": "Flyscene::raytraceScene() (reference frame driver incl. result.ppm)",
            "elapsed_s": s, "threads": thr, "value": rays / s / 1e6, "unit": "Mrays/s", "rays_per_frame": rays}


def census_rays(wl, stride):
    """Rays (reference census, SURVEY.md App. A.8, shadow gate+sample merged in point mode) that the
    sampled pixels generate, counted by the C port on the same pixels.  Not timed."""
    from oracle import oracle as O
    arrs, spheres, sphere_mat = workload_arrays(wl)
    orc = O.Oracle(O.BakedScene(*arrs, spheres=spheres, sphere_mat=sphere_mat), area=wl["area"], point=wl["point"],
                   max_depth=wl["max_depth"], grid=wl["grid"])
    w, h = wl["w"], wl["h"]
    cam = O.Oracle.camera((0, 0, 2), np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 2]], np.float32), (0, 0, w, h),
                          60.0, np.float32(w) / np.float32(h))
    orc.census(reset=True)
    pxy, *_ = orc.render(cam, np.array([[-1, 1, 1]], np.float32), w, h, stride=stride,
                         threads=max(1, (os.cpu_count() or 2) - 1))
    prim, shadow, sec = [int(x) for x in orc.census(reset=True)]
    return prim + shadow + sec, len(pxy)


def run_reference_arm(args, wl, rays_per_pixel):
    """bench.py --impl reference: K timed bounded samples of the reference CPU path."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    stride = {"c1": 3, "c2": 4, "c3": 24, "c4": 48, "c5": 96}[args.workload]
    try:
        rays_sample, pix_sample = census_rays(wl, stride)
        rays_per_pixel = rays_sample / max(1, pix_sample)
    except Exception:
        pass
    for _ in range(args.warmup):
        reference_sample(wl, stride * 2)
    secs, pix, kind, thr = [], 0, "reference", 1
    for _ in range(args.steps):
        s, pix, kind, thr = reference_sample(wl, stride)
        secs.append(s)
    total = float(np.sum(secs))
    rays = pix * rays_per_pixel
    value = rays * len(secs) / total / 1e6
    sample = f"every {stride}th pixel in x and y of the {wl['w']}x{wl['h']} frame ({pix} pixels) per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(secs), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "sample": sample, "rays_per_pixel": rays_per_pixel},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": thr, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ms_per_frame_extrapolated": 1e3 * total / len(secs) * (wl["w"] * wl["h"]) / max(1, pix),
        "reference_frame_driver_c1": reference_raytrace_scene_c1(),
    }
    print(json.dumps(line), flush=True)


def kernel_source_hash():
    """sha256 over the CUDA sources of the library: ties a committed ncu capture to the build it was taken on."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "raytracer-in-cpp_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def ncu_metrics_for(workload, kernel):
    """Measured per-launch counters of `kernel` on `workload` (dram bytes, executed thread-instructions) from the
    committed ncu capture -- only if that capture was taken on the kernel sources of THIS build."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_metrics.json")
    if not os.path.exists(path):
        return None
    d = json.load(open(path))
    if d.get("source_hash") != kernel_source_hash():
        return None
    m = d.get("workloads", {}).get(workload, {}).get(kernel)
    if m:
        m = dict(m)
        m["source"] = "ncu --set full capture on this build (profiles/r02_ncu_metrics.json, source hash %s)" % d["source_hash"]
    return m


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
class BandGather:
    """Interleaved row bands -> rank 0.  Ranks own different row counts (the last band may be partial),
    so every rank sends a buffer padded to the largest count; rank 0 scatters the valid rows of each
    part into the frame with one index_copy per rank.  The collective is dist.gather (NCCL on GPUs,
    gloo in the CPU test)."""

    def __init__(self, rows, width, height, rank, world, dist, torch, dev):
        self.rank, self.world, self.dist, self.torch = rank, world, dist, torch
        self.n_local = len(rows)
        n = torch.tensor([self.n_local], dtype=torch.int64, device=dev)
        counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(counts, n)
        self.counts = [int(c.item()) for c in counts]
        self.max_rows = max(self.counts)
        padded = torch.full((self.max_rows,), -1, dtype=torch.int64, device=dev)
        padded[: self.n_local] = torch.as_tensor(np.asarray(rows), dtype=torch.int64, device=dev)
        maps = [torch.zeros(self.max_rows, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(maps, padded)
        self.maps = [m[: self.counts[r]] for r, m in enumerate(maps)]
        self.frame = torch.zeros((height, width, 4), dtype=torch.uint8, device=dev) if rank == 0 else None
        self.parts = [torch.empty((self.max_rows, width, 4), dtype=torch.uint8, device=dev) for _ in range(world)] \
            if rank == 0 else None

    def __call__(self, local_padded):
        """local_padded: [max_rows, W, 4] uint8 (rows beyond n_local are ignored)."""
        self.dist.gather(local_padded, self.parts, dst=0)
        if self.rank != 0:
            return None
        for r in range(self.world):
            self.frame.index_copy_(0, self.maps[r], self.parts[r][: self.counts[r]])
        return self.frame


def gather_frame(local, rows, width, height, rank, world, dist, torch, dev):
    """One-shot helper (tests): gather `local` [n_local, W, 4] to the full frame on rank 0."""
    g = BandGather(rows, width, height, rank, world, dist, torch, dev)
    padded = torch.zeros((g.max_rows, width, 4), dtype=torch.uint8, device=dev)
    padded[: len(rows)] = local
    return g(padded)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 20)")
    ap.add_argument("--min-seconds", type=float, default=0.5,
                    help="repeat the timed K-step block until this many seconds of frames have been timed")
    ap.add_argument("--band-rows", type=int, default=0, help="multi-GPU band height; 0 = band_rows_for(H, world)")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=INT",
                    help="rt_set_option before the scene is created (developer A/B runs), e.g. --opt fused_frame=0")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    wl = WORKLOADS[args.workload]

    # closed-form ray census per pixel is scene dependent; measured by the kernels below.  The
    # reference arm needs it before any GPU work, so it is recorded in a small side file by our arm.
    census_file = os.path.join(tempfile.gettempdir(), f"rt_b200_census_{args.workload}.json")

    if args.impl == "reference":
        rpp = None
        if os.path.exists(census_file):
            rpp = json.load(open(census_file)).get("rays_per_pixel")
        if rpp is None:
            # cube.obj, default camera: 494 209/1e6 of the pixels hit; fall back to
            # the closed form for c1/c2, else count primary rays only
            frac = 0.494209
            rpp = {"c1": 1 + frac * 3, "c2": 1 + frac * (17 + 1 + 17)}.get(args.workload, 1.0)
        run_reference_arm(args, wl, rpp)
        return

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    capi = pkg().capi
    capi.init(local_rank)
    for kv in args.opt:
        k, v = kv.split("=")
        capi.set_option(k, int(v))

    arrs, spheres, sphere_mat = workload_arrays(wl)
    scene = capi.Scene(*arrs, None, spheres, sphere_mat)
    info = scene.info()
    W, H = wl["w"], wl["h"]
    cam = capi.default_camera(W, H)
    lights = capi.Lights(np.array([[-1.0, 1.0, 1.0]], np.float32))  # src/flyscene.cpp:72
    band_rows = args.band_rows or band_rows_for(H, world)
    params = capi.make_params(W, H, wl["area"], wl["point"], wl["max_depth"], wl["grid"], band_rows, rank, world)
    rows = capi.lib().rt_local_rows(C.byref(params))
    gather = BandGather(capi.local_row_map(params), W, H, rank, world, dist, torch, dev) if world > 1 else None
    max_rows = gather.max_rows if gather else rows
    local = torch.zeros((max_rows, W, 4), dtype=torch.uint8, device=dev)
    # a real (non-default) stream: the library replays a captured CUDA graph on it, and the legacy
    # default stream cannot be captured; everything below (renders, NCCL ops, events) runs on it
    work_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(work_stream)
    stream = work_stream.cuda_stream
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step():
        scene.render_device(cam, lights, params, local.data_ptr(), stream=stream)
        if gather is not None:
            gather(local)

    # Fused alternative for N > 1: every rank's kernels store their pixels straight into rank 0's
    # framebuffer through an NVLink peer mapping (CUDA IPC); a 4-byte all-reduce on the stream is the
    # only collective left (completion signal).  The NCCL gather above is kept as the baseline.
    shared = None
    step_peer = None
    if world > 1:
        try:
            hbuf = torch.zeros(64, dtype=torch.uint8, device=dev)
            if rank == 0:
                shared = capi.SharedFrame(W, H)
                hbuf.copy_(torch.frombuffer(bytearray(shared.handle), dtype=torch.uint8))
            dist.broadcast(hbuf, src=0)
            if rank != 0:
                shared = capi.SharedFrame(W, H, bytes(hbuf.cpu().numpy().tobytes()))
            params_peer = capi.make_params(W, H, wl["area"], wl["point"], wl["max_depth"], wl["grid"], band_rows, rank, world)
            params_peer.out_full_frame = 1
            frame_seq = [0]

            def step_peer():
                # no collective: every rank raises its flag behind the shared frame when its pixels are stored, rank 0
                # queues a one-warp kernel that waits for all flags (rt_shared_frame_signal / _wait)
                frame_seq[0] += 1
                scene.render_device(cam, lights, params_peer, shared.ptr.value, stream=stream)
                shared.signal(rank, frame_seq[0], stream)
                if rank == 0:
                    shared.wait(world, frame_seq[0], stream)
        except Exception as ex:  # peer mapping unavailable: keep the NCCL path only
            print(f"[rank {rank}] peer-store path unavailable: {ex}", file=sys.stderr)
            shared, step_peer = None, None
        ok = torch.tensor([1 if step_peer is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            step_peer = None

    # ---- one untimed stats frame: ray census, per-kernel times, traversal counters ----
    capi.set_option("stats", 1)
    st = capi.RtStats()
    scene.render_device(cam, lights, params, local.data_ptr(), stream=stream, stats=st)
    stats = st.as_dict()
    capi.set_option("stats", 0)
    rays_local = stats["rays_primary"] + stats["rays_shadow"] + stats["rays_secondary"]
    rays_t = torch.tensor([rays_local, stats["rays_primary"], stats["rays_shadow"], stats["rays_secondary"]],
                          dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(rays_t)
    rays_total, n_primary, n_shadow, n_secondary = [float(x) for x in rays_t.tolist()]

    # ---- timed: blocks of exactly K steps, CUDA events on the launch stream, L2 flushed between steps ----
    # One block = the K steps the contract asks for (barrier + synchronize on both sides, max over ranks).  A 20-step
    # block of a 0.3 ms frame is a 7 ms measurement: too short for the clock sampler to see and with unknown noise.
    # So the block is REPEATED until at least `--min-seconds` of frames have been timed (at most 200 blocks); `value`
    # and `ms_per_step` come from the MEDIAN block, the spread over blocks is reported beside them.
    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        sampler_ = ClockSampler(local_rank)
        sampler_.start()
        blocks, wall_total, steps_total = [], 0.0, 0
        while True:
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            w0 = time.perf_counter()
            for a, b in evs:
                flush.zero_()
                a.record()
                fn()
                b.record()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            wall_total += time.perf_counter() - w0
            ms_t = torch.tensor([float(sum(a.elapsed_time(b) for a, b in evs))], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
            blocks.append(float(ms_t.item()))
            steps_total += args.steps
            # every rank must take the same decision: it is taken on the all-reduced block times
            if sum(blocks) * 1e-3 >= args.min_seconds or len(blocks) >= 200:
                break
        clocks_ = sampler_.stop()
        per_step = np.array(blocks) / args.steps
        spread_ = {"blocks": len(blocks), "steps_timed": steps_total, "ms_per_step_median": float(np.median(per_step)),
                   "ms_per_step_min": float(per_step.min()), "ms_per_step_max": float(per_step.max()),
                   "ms_per_step_p10": float(np.percentile(per_step, 10)), "ms_per_step_p90": float(np.percentile(per_step, 90)),
                   "rel_spread_p10_p90": float((np.percentile(per_step, 90) - np.percentile(per_step, 10)) / np.median(per_step))}
        return float(np.median(blocks)), wall_total, clocks_, spread_

    kernel_ms = {"trace": 0.0, "shadow": 0.0, "shade": 0.0}
    ms_total, wall, clocks, spread = timed(step)
    nccl_line = None
    if step_peer is not None:
        ms_peer, wall_peer, clocks_peer, spread_peer = timed(step_peer)
        nccl_line = {"value": rays_total * args.steps / (ms_total * 1e-3) / 1e6, "ms_per_step": ms_total / args.steps,
                     "note": "baseline: bands gathered to rank 0 with torch.distributed.gather (NCCL) + scatter"}
        # correctness of the fused path: rank 0's peer-assembled frame == the NCCL-gathered frame
        match = None
        torch.cuda.synchronize(); dist.barrier()
        if rank == 0:
            match = bool((torch.from_numpy(shared.to_host()).to(dev) == gather.frame).all().item())
        ms_total, wall, clocks, spread = ms_peer, wall_peer, clocks_peer, spread_peer
        nccl_line["frames_match"] = match
    ms_per_step = ms_total / args.steps
    value = rays_total * args.steps / (ms_total * 1e-3) / 1e6

    # ---- per-kernel durations (live, CUDA events inside the library on the same stream) ----
    # Wavefront frames: one CUDA-event span per kernel family.  Fused frames (RtStats.fused): ONE kernel, timed as a
    # whole ("frame"); its split into trace / shadow / shade comes from the stats frame above, where the kernel
    # counted the warp-cycles of its phases.
    reps = min(10, args.steps)
    fused = bool(stats.get("fused"))
    if fused:
        kernel_ms = {"frame": 0.0}
    for _ in range(reps):
        flush.zero_()
        s2 = capi.RtStats()
        scene.render_device(cam, lights, params, local.data_ptr(), stream=stream, stats=s2)
        if fused:
            kernel_ms["frame"] += s2.ms_total / reps
        else:
            kernel_ms["trace"] += s2.ms_trace / reps
            kernel_ms["shadow"] += s2.ms_shadow / reps
            kernel_ms["shade"] += s2.ms_shade / reps
    if fused and stats["ms_total"] > 0:
        kernel_ms["phase_share"] = {k: stats["ms_" + k] / stats["ms_total"] for k in ("trace", "shadow", "shade")}

    # ---- e2e: the user-facing blocking call with pinned HOST buffers, copies inside the timed region ----
    # N = 1: rt_render() (H2D of camera/lights/params, render, D2H of the packed frame).
    # N > 1: every rank renders its bands (fused peer store, else NCCL gather) and rank 0 reads the
    # assembled full frame back to pinned host memory.
    k2 = args.e2e_steps or min(args.steps, 20)
    h2d = C.sizeof(capi.RtCamera) + C.sizeof(capi.RtParams) + 12 * lights.c.n + 16
    if world == 1:
        host = torch.zeros((rows, W, 4), dtype=torch.uint8).pin_memory()
        host_np = host.numpy()

        def e2e_step():
            scene.render(cam, lights, params, want_face=False, want_t=False, want_rgb=False, want_stats=False,
                         out_rgba=host_np)
    else:
        # the full frame lives in POSIX shared memory, page-locked in every rank (rt_host_frame_*); every rank copies
        # its OWN bands into it over its own PCIe link (rt_render_into_frame), so the 8.3 MB do not all leave through
        # rank 0.  Two process barriers of the library per frame: "previous frame consumed" and "frame complete".
        names = [f"/rt_b200_frame_{os.getpid()}" if rank == 0 else None]
        dist.broadcast_object_list(names, src=0)
        hostframe = capi.HostFrame(names[0], W, H, create=True) if rank == 0 else None
        dist.barrier()
        if rank != 0:
            hostframe = capi.HostFrame(names[0], W, H, create=False)
        host_np = hostframe.array

        def e2e_step():
            hostframe.barrier(world)
            scene.render_into_frame(cam, lights, params, hostframe.ptr)
            hostframe.barrier(world)
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(k2):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    dt_t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
    dt = float(dt_t.item())
    e2e = {"value": rays_total * k2 / dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d) * world,
           "d2h_bytes_per_step": int(H * W * 4), "ms_per_step": 1e3 * dt / k2, "steps": k2,
           "call": "rt_render (blocking, pinned host frame)" if world == 1 else
                   "rt_render_into_frame on every rank: own bands -> shared pinned host frame (POSIX shm), library process barrier"}
    if world > 1:
        # the assembled host frame must be the frame: compare it with rank 0's device-assembled one
        hostframe.barrier(world)
        if rank == 0 and shared is not None and step_peer is not None:
            e2e["frame_matches_device_frame"] = bool((host_np == shared.to_host()).all())
        hostframe.barrier(world)
    if world == 1:
        # the same steps through the streaming call: rt_render_submit / rt_render_wait, two frames in
        # flight, every frame still copied to pinned host memory inside the timed region
        hosts = [torch.zeros((rows, W, 4), dtype=torch.uint8).pin_memory().numpy() for _ in range(2)]
        def pipelined(k):
            tk = []
            for i in range(k):
                if i >= 2:
                    scene.wait(tk[i - 2])
                tk.append(scene.submit(cam, lights, params, hosts[i & 1]))
            scene.wait(tk[-1])
        pipelined(4)
        t0 = time.perf_counter()
        pipelined(k2)
        dtp = time.perf_counter() - t0
        assert (hosts[(k2 - 1) & 1] == host_np).all(), "streaming and blocking frames differ"
        e2e["pipelined"] = {"value": rays_total * k2 / dtp / 1e6, "unit": "Mrays/s", "ms_per_step": 1e3 * dtp / k2,
                            "call": "rt_render_submit / rt_render_wait, 2 frames in flight"}

    # ---- the same frame from ONE host process driving all N GPUs (rt_multi_*: what the C++ facade / rt_cli use) ----
    # Measured by rank 0 alone while the other ranks wait on the library's CPU-side barrier (their GPUs are idle).
    single_process = None
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        if rank == 0:
            try:
                multi = capi.Multi(list(range(world)), *arrs, None, spheres, sphere_mat)
                mp = capi.make_params(W, H, wl["area"], wl["point"], wl["max_depth"], wl["grid"], band_rows)
                kk = min(args.steps, 50)
                for _ in range(args.warmup):
                    multi.render_device(cam, lights, mp)
                dev_ms, t0 = [], time.perf_counter()
                for _ in range(kk):
                    dev_ms.append(multi.render_device(cam, lights, mp)[1])
                wall_dev = (time.perf_counter() - t0) / kk
                ptr, _ = multi.render_device(cam, lights, mp)
                fr = np.zeros((H, W, 4), np.uint8)
                capi.lib().rt_device_copy_to_host(fr.ctypes.data, ptr, fr.nbytes)
                for _ in range(args.warmup):
                    multi.render(cam, lights, mp, out=host_np)
                t0 = time.perf_counter()
                for _ in range(kk):
                    multi.render(cam, lights, mp, out=host_np)
                wall_host = (time.perf_counter() - t0) / kk
                single_process = {
                    "call": "rt_multi_render_device / rt_multi_render: one process, one worker thread per GPU",
                    "device_frame": {"ms_per_step_slowest_device": float(np.median(dev_ms)), "ms_per_step_wall": 1e3 * wall_dev,
                                     "value": rays_total / wall_dev / 1e6, "unit": "Mrays/s",
                                     "frame_matches": bool(shared is not None and (fr == shared.to_host()).all())},
                    "host_frame": {"ms_per_step_wall": 1e3 * wall_host, "value": rays_total / wall_host / 1e6, "unit": "Mrays/s",
                                   "frame_matches": bool(shared is not None and (host_np == shared.to_host()).all())},
                    "note": "no L2 flush between these frames; wall = host clock around the blocking call"}
                multi.close()
            except Exception as ex:
                single_process = {"unavailable": str(ex)[:300]}
        hostframe.barrier(world)
        hostframe.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    # The path is SM-issue bound, not HBM bound: the algorithmic work of a launch is
    #     W_alg = 16 * ray-AABB tests + 40 * ray-triangle tests + 120 * shaded light samples   [thread-instructions]
    # (counted by a stats build of the same kernels on the same frame) and the peak is one instruction per lane,
    # scheduler and cycle: 148 SMs x 4 schedulers x 32 lanes x f_SM.  `frac` is that of the dominant kernel over its
    # own CUDA-event duration, `frac_frame` of the whole frame over ms_per_step.  The HBM side is reported beside it
    # from MEASURED DRAM bytes (ncu) when the committed capture was taken on exactly these kernel sources.
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    sm_mhz = float(peaks.get("sm_max_mhz") or clocks.get("sm_max_mhz") or 1965)
    n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
    issue_peak = n_sm * 4 * 32 * sm_mhz * 1e6
    w_kernel = {"trace": 16.0 * stats["box_tests"] + 40.0 * stats["tri_tests"],
                "shadow": 16.0 * stats["box_tests_shadow"] + 40.0 * stats["tri_tests_shadow"],
                "shade": 120.0 * stats["shade_samples"]}
    w_frame = sum(w_kernel.values())
    if fused:
        dom, kname, w_dom, dom_ms = "frame", "k_frame", w_frame, max(kernel_ms["frame"], 1e-6)
    else:
        dom = max(("trace", "shadow", "shade"), key=lambda k: kernel_ms[k])
        kname = {"trace": "k_trace_nearest", "shadow": "k_shadow", "shade": "k_shade"}[dom]
        w_dom, dom_ms = w_kernel[dom], max(kernel_ms[dom], 1e-6)
    achieved = w_dom / (dom_ms * 1e-3)
    measured = ncu_metrics_for(args.workload, kname) if world == 1 else None
    traffic = measured.get("dram_bytes") if measured else None
    scene_bytes = int(info.get("device_bytes", 0))
    roofline = {"bound": "sm_issue", "achieved": achieved, "peak": issue_peak, "unit": "thread-instr/s",
                "frac": achieved / issue_peak, "frac_frame": w_frame / (ms_per_step * 1e-3) / issue_peak,
                "kernel": kname, "kernel_ms": dom_ms, "algorithmic_thread_instr_per_launch": w_dom,
                "algorithmic_thread_instr_per_frame": w_frame,
                "model": "16 per ray-AABB test + 40 per ray-triangle test + 120 per shaded light sample (SURVEY.md 8d); "
                         "peak = %d SMs x 4 schedulers x 32 lanes x %.0f MHz" % (n_sm, sm_mhz),
                "executed_over_algorithmic": (measured["thread_inst_executed"] / w_dom) if measured and w_dom else None,
                "traffic": traffic,
                "hbm": {"bytes_per_launch_measured": traffic,
                        "gbs": (traffic / (dom_ms * 1e-3) / 1e9) if traffic else None,
                        "peak_gbs": hbm_peak, "peak_source": peak_src,
                        "frac": (traffic / (dom_ms * 1e-3) / 1e9 / hbm_peak) if traffic else None,
                        "compulsory_bytes_per_frame": scene_bytes + 4 * W * H,
                        "source": (measured or {}).get("source",
                                                     "no ncu capture of these kernel sources is committed (profiles/r02_ncu_metrics.json "
                                                     "is keyed by a hash of csrc/): not reported rather than replayed from an older build")}}

    json.dump({"rays_per_pixel": rays_total / (W * H)}, open(census_file, "w"))

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            stride = {"c1": 2, "c2": 3, "c3": 20, "c4": 40, "c5": 80}[args.workload]
            s, pix, kind, thr = reference_sample(wl, stride)
            rpp = rays_total / (W * H)
            cpu_baseline = {"value": pix * rpp / s / 1e6, "unit": "Mrays/s", "cores": thr, "kind": kind,
                            "sample": f"every {stride}th pixel in x and y of the {W}x{H} frame ({pix} pixels), "
                                      f"{s:.2f} s; extrapolated full frame {s * W * H / pix:.1f} s",
                            "ms_per_frame_extrapolated": 1e3 * s * W * H / pix}
        except Exception as ex:  # the baseline is reporting only; never fail the bench for it
            cpu_baseline = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "unavailable", "sample": str(ex)[:200]}

    line = {
        "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "image": [W, H], "triangles": int(arrs[0].shape[0]),
                   "spheres": int(0 if spheres is None else len(spheres)), "lights": 1,
                   "samples_per_light": 1 if wl["point"] else wl["grid"][0] * wl["grid"][1],
                   "max_depth": wl["max_depth"], "l2": "flushed between timed frames (512 MiB memset)",
                   "parallelism": (f"{world} x interleaved {band_rows}-row bands; " +
                                   ("kernels store pixels into rank 0's framebuffer over NVLink peer memory (CUDA IPC), "
                                    "completion by per-rank flags behind the frame that a one-warp kernel on rank 0 polls (no collective)" if nccl_line else "NCCL gather to rank 0"))
                   if world > 1 else "1 GPU",
                   "bvh": info, "options": args.opt},
        "rays": {"per_frame": rays_total, "primary": n_primary, "shadow": n_shadow, "secondary": n_secondary,
                 "note": "gate and sample shadow rays are separate queries in area mode (as in the reference); "
                         "in point mode the identical gate/sample ray is traced and counted once"},
        "mpix_per_s": W * H * args.steps / (ms_total * 1e-3) / 1e6,
        "work": {k: stats[k] for k in ("box_tests", "tri_tests", "box_tests_shadow", "tri_tests_shadow", "shade_samples",
                                         "filter_checks", "filter_slow", "filter_rejects", "shadow_rays_traced")},
        "timing": spread,
        "kernel_ms": kernel_ms, "frame_path": "fused (one persistent kernel)" if fused else "wavefront (CUDA graph of per-level kernels)",
        "gpu_launches": int(stats["kernel_launches"] * args.steps),
        "clocks": clocks, "wall_s": wall, "nccl_gather_baseline": nccl_line, "single_process_multi_gpu": single_process, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
