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

N > 1 (torchrun, one rank per GPU): the frame is split into interleaved 8-row bands
(band b -> rank b mod N), each rank renders its bands and the bands are gathered to rank 0 with NCCL;
the timed step includes the gather.  Strong scaling: the frame is fixed.

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


def workload_arrays_host(wl):
    """Same as workload_arrays; named separately to make clear that only the host-side loader of the
    C-ABI library is used (no GPU needed)."""
    return workload_arrays(wl)


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
                  grid=(tuple(wl["grid"]) if patched else None))
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


def census_rays(wl, stride):
    """Rays (reference census, SURVEY.md App. A.8, shadow gate+sample merged in point mode) that the
    sampled pixels generate, counted by the C port on the same pixels.  Not timed."""
    from oracle import oracle as O
    arrs, spheres, sphere_mat = workload_arrays_host(wl)
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
    }
    print(json.dumps(line), flush=True)


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
    ap.add_argument("--band-rows", type=int, default=0, help="multi-GPU band height; 0 = band_rows_for(H, world)")
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
            token = torch.zeros(1, dtype=torch.int32, device=dev)

            def step_peer():
                scene.render_device(cam, lights, params_peer, shared.ptr.value, stream=stream)
                dist.all_reduce(token)
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

    # ---- timed: exactly K steps, CUDA events on the launch stream, L2 flushed between steps ----
    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        sampler_ = ClockSampler(local_rank)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler_.start()
        w0 = time.perf_counter()
        for a, b in evs:
            flush.zero_()
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        wall_ = time.perf_counter() - w0
        clocks_ = sampler_.stop()
        ms_t = torch.tensor([float(sum(a.elapsed_time(b) for a, b in evs))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
        return float(ms_t.item()), wall_, clocks_

    kernel_ms = {"trace": 0.0, "shadow": 0.0, "shade": 0.0}
    ms_total, wall, clocks = timed(step)
    nccl_line = None
    if step_peer is not None:
        ms_peer, wall_peer, clocks_peer = timed(step_peer)
        nccl_line = {"value": rays_total * args.steps / (ms_total * 1e-3) / 1e6, "ms_per_step": ms_total / args.steps,
                     "note": "baseline: bands gathered to rank 0 with torch.distributed.gather (NCCL) + scatter"}
        # correctness of the fused path: rank 0's peer-assembled frame == the NCCL-gathered frame
        match = None
        torch.cuda.synchronize(); dist.barrier()
        if rank == 0:
            match = bool((torch.from_numpy(shared.to_host()).to(dev) == gather.frame).all().item())
        ms_total, wall, clocks = ms_peer, wall_peer, clocks_peer
        nccl_line["frames_match"] = match
    ms_per_step = ms_total / args.steps
    value = rays_total * args.steps / (ms_total * 1e-3) / 1e6

    # ---- per-kernel durations (live, CUDA events inside the library on the same stream) ----
    reps = min(10, args.steps)
    for _ in range(reps):
        flush.zero_()
        s2 = capi.RtStats()
        scene.render_device(cam, lights, params, local.data_ptr(), stream=stream, stats=s2)
        kernel_ms["trace"] += s2.ms_trace / reps
        kernel_ms["shadow"] += s2.ms_shadow / reps
        kernel_ms["shade"] += s2.ms_shade / reps

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
        host = torch.zeros((H, W, 4), dtype=torch.uint8).pin_memory()
        host_np = host.numpy()
        fn = step_peer if step_peer is not None else step

        def e2e_step():
            fn()
            torch.cuda.synchronize()
            if rank == 0:
                if step_peer is not None:
                    capi.lib().rt_device_copy_to_host(host_np.ctypes.data, shared.ptr, host_np.nbytes)
                else:
                    host.copy_(gather.frame)
                    torch.cuda.synchronize()
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
           "call": "rt_render (blocking, pinned host frame)" if world == 1 else "bands + assembly on rank 0 + D2H"}
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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    dom = max(kernel_ms, key=kernel_ms.get)
    if dom == "shadow":
        nb, nt = stats["box_tests_shadow"], stats["tri_tests_shadow"]
        out_bytes = stats["rays_shadow"]  # one visibility byte per shadow job
    elif dom == "trace":
        nb, nt = stats["box_tests"], stats["tri_tests"]
        out_bytes = 8 * (stats["rays_primary"] + stats["rays_secondary"])
    else:
        nb, nt = 0, 0
        out_bytes = 4 * stats["pixels"]
    # 32 B per ray-AABB test (one child box of a pair node), 48 B per ray-triangle test
    alg_bytes = 32.0 * nb + 48.0 * nt + out_bytes
    dom_ms = max(kernel_ms[dom], 1e-6)
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    sm_mhz = clocks.get("sm_max_mhz") or 1965
    issue_peak = 148 * 4 * 32 * sm_mhz * 1e6
    w_alg = 16.0 * (stats["box_tests"] + stats["box_tests_shadow"]) + 40.0 * (stats["tri_tests"] + stats["tri_tests_shadow"]) \
        + 120.0 * stats["shade_samples"]
    frame_ms_1gpu = ms_per_step
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    kname = {"trace": "k_trace_nearest", "shadow": "k_shadow", "shade": "k_shade"}[dom]
    if world == 1 and os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload, {}).get(kname)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "kernel": kname,
                # measured DRAM bytes of that launch over its duration: the HBM bandwidth the kernel really draws
                "dram_gbs": (traffic / (dom_ms * 1e-3) / 1e9) if traffic else None,
                "dram_frac": (traffic / (dom_ms * 1e-3) / 1e9 / hbm_peak) if traffic else None,
                "kernel_ms": dom_ms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "note": "achieved = cache-level ALGORITHMIC bytes (32 B per ray-AABB test, 48 B per ray-triangle test, "
                        "SURVEY.md 8d) of the dominant kernel / its CUDA-event duration; traffic = measured DRAM bytes of "
                        "that kernel (ncu --set full, profiles/r01_traffic.json) and dram_gbs / dram_frac what that is per second "
                        "and against the HBM peak: the working set is served by L1/L2, the path is SM-issue / load-latency "
                        "bound, see sm_issue",
                "sm_issue": {"achieved_thread_instr_per_s": w_alg / (frame_ms_1gpu * 1e-3),
                             "peak_thread_instr_per_s": issue_peak, "frac": w_alg / (frame_ms_1gpu * 1e-3) / issue_peak,
                             "model": "16*N_box + 40*N_tri + 120*N_samples over the whole frame (rank 0's share)"}}

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
                                    "4-byte NCCL all-reduce as completion signal" if nccl_line else "NCCL gather to rank 0"))
                   if world > 1 else "1 GPU",
                   "bvh": info},
        "rays": {"per_frame": rays_total, "primary": n_primary, "shadow": n_shadow, "secondary": n_secondary,
                 "note": "gate and sample shadow rays are separate queries in area mode (as in the reference); "
                         "in point mode the identical gate/sample ray is traced and counted once"},
        "mpix_per_s": W * H * args.steps / (ms_total * 1e-3) / 1e6,
        "work": {k: stats[k] for k in ("box_tests", "tri_tests", "box_tests_shadow", "tri_tests_shadow", "shade_samples",
                                         "filter_checks", "filter_slow", "filter_rejects", "shadow_rays_traced")},
        "kernel_ms": kernel_ms, "gpu_launches": int(stats["kernel_launches"] * args.steps),
        "clocks": clocks, "wall_s": wall, "nccl_gather_baseline": nccl_line, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
