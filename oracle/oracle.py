"""oracle.py -- TEST INFRASTRUCTURE: ctypes binding of oracle/liboracle.so (rt_oracle.c) plus readers
for the binary dumps written by oracle/_ref/ref_oracle (oracle/ref_harness/ref_driver.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_BIN = os.path.join(REF_DIR, "ref_oracle")
REF_BIN_PATCHED = os.path.join(REF_DIR, "ref_oracle_patched")
REF_SCENES = os.path.join(REF_DIR, "scenes")


def build_port(force: bool = False) -> str:
    """Compile the C restatement (gcc, pinned flags in oracle/Makefile)."""
    src = os.path.join(HERE, "rt_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(HERE, "rt_oracle.h"))
    ):
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    return LIB_PATH


def build_ref() -> bool:
    """Compile the reference itself when /root/reference is present (this container only)."""
    if not os.path.isdir("/root/reference/src"):
        return os.path.exists(REF_BIN)
    subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref"])
    return True


def have_ref() -> bool:
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


# ----------------------------------------------------------------------------------------------
# ctypes structures (mirror rt_oracle.h)
# ----------------------------------------------------------------------------------------------
class OrMaterial(C.Structure):
    _fields_ = [("kd", C.c_float * 3), ("ks", C.c_float * 3), ("ns", C.c_float), ("ni", C.c_float),
                ("illum", C.c_int32)]


class OrSceneDesc(C.Structure):
    _fields_ = [("n_faces", C.c_int32), ("verts", C.c_void_p), ("fnormals", C.c_void_p),
                ("vnormals", C.c_void_p), ("mat_id", C.c_void_p), ("n_mats", C.c_int32),
                ("mats", C.c_void_p), ("model_matrix", C.c_float * 12), ("n_spheres", C.c_int32),
                ("spheres", C.c_void_p), ("sphere_mat", C.c_void_p)]


class OrCamera(C.Structure):
    _fields_ = [("eye", C.c_float * 3), ("view_inv", C.c_float * 12), ("viewport", C.c_float * 4),
                ("fovy", C.c_float), ("aspect", C.c_float)]


class OrParams(C.Structure):
    _fields_ = [("area_light", C.c_int32), ("point_light", C.c_int32), ("max_depth", C.c_int32),
                ("usteps", C.c_int32), ("vsteps", C.c_int32), ("area_len_x", C.c_float),
                ("area_len_y", C.c_float), ("light_color", C.c_float * 3), ("capacity", C.c_int32),
                ("candidates", C.c_int32), ("recursion_guard", C.c_int32), ("sphere_seed", C.c_uint32),
                ("sphere_radius", C.c_float)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_port()
        L = C.CDLL(LIB_PATH)
        L.or_scene_create.restype = C.c_void_p
        L.or_scene_create.argtypes = [C.POINTER(OrSceneDesc), C.POINTER(OrParams)]
        L.or_scene_destroy.argtypes = [C.c_void_p]
        L.or_scene_root_box.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.or_scene_octree_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.or_screen_to_world.argtypes = [C.POINTER(OrCamera), C.c_float, C.c_float, C.c_void_p]
        L.or_box_intersect.restype = C.c_int
        L.or_box_intersect.argtypes = [C.c_void_p] * 4
        L.or_octree_candidates.restype = C.c_int
        L.or_octree_candidates.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.or_ray_triangle.restype = C.c_float
        L.or_ray_triangle.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.or_light_samples.restype = C.c_int
        L.or_light_samples.argtypes = [C.POINTER(OrParams), C.c_void_p, C.c_void_p]
        L.or_light_strikes.restype = C.c_int
        L.or_light_strikes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.or_trace_ray.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p]
        L.or_phong_shade.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.or_quantize.restype = C.c_int
        L.or_quantize.argtypes = [C.c_float]
        L.or_render_pixels.argtypes = [C.c_void_p, C.POINTER(OrCamera), C.c_void_p, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int]
        L.or_default_params.argtypes = [C.POINTER(OrParams)]
        L.or_census_get.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ----------------------------------------------------------------------------------------------
# Baked scene (the flat arrays both the oracle and the product C-ABI consume)
# ----------------------------------------------------------------------------------------------
@dataclass
class BakedScene:
    verts: np.ndarray       # [T,3,3] f32 world space
    fnormals: np.ndarray    # [T,3]
    vnormals: np.ndarray    # [T,3,3]
    mat_id: np.ndarray      # [T] i32
    mats: np.ndarray        # [M,9] f32: kd3 ks3 ns ni illum
    model_matrix: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32)[:3].copy())
    spheres: np.ndarray | None = None      # [S,4]
    sphere_mat: np.ndarray | None = None   # [S]
    # camera / lights as dumped (optional)
    eye: np.ndarray | None = None
    view_inv: np.ndarray | None = None     # [3,4]
    view: np.ndarray | None = None
    viewport: np.ndarray | None = None
    fovy: float = 60.0
    aspect: float = 1.0
    pscale: float = 0.0
    lights: np.ndarray | None = None       # [L,3]
    light_color: np.ndarray | None = None
    light_radius: float = 1.0
    root_min: np.ndarray | None = None
    root_max: np.ndarray | None = None
    centroid: np.ndarray | None = None
    radius: float = 0.0
    norm_scale: float = 0.0
    vertex_ids: np.ndarray | None = None   # [T,3]
    obj_verts: np.ndarray | None = None    # [NV,3]
    octree_stats: np.ndarray | None = None
    width: int = 0
    height: int = 0
    area: int = 0
    point: int = 1

    @property
    def n_faces(self):
        return int(self.verts.shape[0])


def load_scene_dump(path: str) -> BakedScene:
    """Read a --dump-scene file written by ref_driver.cpp (format version 2)."""
    b = open(path, "rb").read()
    assert b[:4] == b"RTSC", "not a scene dump"
    ver, T, M, L, NV, W, H, area, point = struct.unpack_from("<9i", b, 4)
    assert ver == 2
    o = 40

    def take(n, dt="<f4"):
        nonlocal o
        a = np.frombuffer(b, dt, n, o).copy()
        o += a.nbytes
        return a

    eye = take(3)
    view_inv = take(12).reshape(3, 4)
    view = take(12).reshape(3, 4)
    viewport = take(4)
    fovy, aspect, pscale = take(3)
    light_color = take(3)
    light_radius = float(take(1)[0])
    root_min = take(3)
    root_max = take(3)
    centroid = take(3)
    radius = float(take(1)[0])
    norm_scale = float(take(1)[0])
    model = take(12).reshape(3, 4)
    lights = take(3 * L).reshape(L, 3)
    mats = np.zeros((M, 9), np.float32)
    for m in range(M):
        mats[m, :8] = take(8)
        mats[m, 8] = float(take(1, "<i4")[0])
    verts = take(T * 9).reshape(T, 3, 3)
    fn = take(T * 3).reshape(T, 3)
    vn = take(T * 9).reshape(T, 3, 3)
    mat_id = take(T, "<i4")
    vids = take(T * 3, "<i4").reshape(T, 3)
    obj_verts = take(NV * 3).reshape(NV, 3)
    stats = take(4, "<i8")
    return BakedScene(verts, fn, vn, mat_id, mats, model, None, None, eye, view_inv, view, viewport,
                      float(fovy), float(aspect), float(pscale), lights, light_color, light_radius,
                      root_min, root_max, centroid, radius, norm_scale, vids, obj_verts, stats, W, H,
                      area, point)


@dataclass
class RefRender:
    width: int
    height: int
    stride: int
    threads: int
    pxy: np.ndarray    # [N,2] i32 (x, y)
    rgb: np.ndarray    # [N,3] f32
    face: np.ndarray   # [N] i32
    t: np.ndarray      # [N] f32
    render_s: float
    build_s: float
    init_s: float


def load_render_dump(path: str) -> RefRender:
    b = open(path, "rb").read()
    assert b[:4] == b"RTOR"
    ver, W, H, stride, thr = struct.unpack_from("<5i", b, 4)
    (N,) = struct.unpack_from("<q", b, 24)
    rs, bs, is_ = struct.unpack_from("<3d", b, 32)
    o = 56
    pxy = np.frombuffer(b, "<i4", N * 2, o).reshape(N, 2).copy(); o += N * 8
    rgb = np.frombuffer(b, "<f4", N * 3, o).reshape(N, 3).copy(); o += N * 12
    fid = np.frombuffer(b, "<i4", N, o).copy(); o += N * 4
    t = np.frombuffer(b, "<f4", N, o).copy()
    return RefRender(W, H, stride, thr, pxy, rgb, fid, t, rs, bs, is_)


def run_ref(scene_obj: str, out: str, w: int, h: int, area: int = 0, point: int = 1, stride: int = 1,
            offx: int = 0, offy: int = 0, threads: int = 0, dump_scene: str | None = None,
            lights=None, cam_rot=None, cam_trans=None, primary_only=False, max_depth=None, grid=None,
            timeout=3600, capture=True) -> str:
    """Run the reference (oracle/_ref/ref_oracle[_patched]) headless; returns its stdout (JSON line).
    capture=False (timing runs) skips the driver's untimed second pass that records primary face id / t."""
    patched = max_depth is not None or grid is not None
    cmd = [REF_BIN_PATCHED if patched else REF_BIN, "--scene", scene_obj, "--w", str(w), "--h", str(h),
           "--area", str(area), "--point", str(point), "--stride", str(stride), "--offx", str(offx),
           "--offy", str(offy), "--threads", str(threads)]
    if out:
        cmd += ["--out", out]
    if dump_scene:
        cmd += ["--dump-scene", dump_scene]
    if lights is not None:
        cmd += ["--lights", ";".join(",".join(repr(float(c)) for c in l) for l in lights)]
    if cam_rot is not None:
        cmd += ["--cam-rot", repr(float(cam_rot[0])), repr(float(cam_rot[1]))]
    if cam_trans is not None:
        cmd += ["--cam-trans"] + [repr(float(c)) for c in cam_trans]
    if primary_only:
        cmd += ["--primary-only"]
    if not capture:
        cmd += ["--no-capture"]
    if max_depth is not None:
        cmd += ["--max-depth", str(max_depth)]
    if grid is not None:
        cmd += ["--grid", str(grid[0]), str(grid[1])]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"ref_oracle failed rc={r.returncode}: {r.stderr[-2000:]}")
    return r.stdout


def run_ref_raytrace_scene(scene_obj: str, size: int, area: int = 0, point: int = 1, timeout=3600):
    """The reference's OWN frame driver, Flyscene::raytraceScene() (src/flyscene.cpp:519-648): its pixel
    pre-pass, its ThreadPool (hardware_concurrency() - 1 workers), traceRay per pixel and the ASCII result.ppm
    write, timed by its own clock (the "ELAPSED TIME:" line it prints, :646).  Square images only (the
    reference indexes pixel_data[x][y] on a [H][W] array).  Returns (elapsed_s, threads, path of result.ppm)."""
    import re
    cmd = [REF_BIN, "--scene", scene_obj, "--w", str(size), "--h", str(size), "--area", str(area), "--point", str(point),
           "--mode", "rts"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    m = re.search(r"ELAPSED TIME:([0-9.eE+-]+)", r.stdout)
    if m is None:
        raise RuntimeError(f"raytraceScene printed no ELAPSED TIME (rc={r.returncode}): {r.stderr[-1000:]}")
    thr = re.search(r"Threads utilized:\s+(\d+)", r.stdout)
    cwd = re.search(r"scratch_cwd (\S+)", r.stderr)
    return float(m.group(1)), int(thr.group(1)) if thr else 0, os.path.join(cwd.group(1), "result.ppm") if cwd else None


# ----------------------------------------------------------------------------------------------
# High-level oracle wrapper
# ----------------------------------------------------------------------------------------------
class Oracle:
    """CPU restatement of the reference render path over a BakedScene."""

    def __init__(self, scene: BakedScene, area=0, point=1, max_depth=-1, grid=(5, 5), candidates=0,
                 capacity=1000, light_color=(1.0, 1.0, 0.0), sphere_seed=1):
        L = lib()
        self.scene = scene
        self.params = OrParams()
        L.or_default_params(C.byref(self.params))
        self.params.area_light = int(area)
        self.params.point_light = int(point)
        self.params.max_depth = int(max_depth)
        self.params.usteps, self.params.vsteps = int(grid[0]), int(grid[1])
        self.params.candidates = int(candidates)
        self.params.capacity = int(capacity)
        self.params.sphere_seed = int(sphere_seed)
        for k in range(3):
            self.params.light_color[k] = float(light_color[k])
        self._keep = []
        d = OrSceneDesc()
        T = scene.n_faces
        d.n_faces = T

        def keep(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            self._keep.append(a)
            return a

        d.verts = _ptr(keep(scene.verts, np.float32))
        d.fnormals = _ptr(keep(scene.fnormals, np.float32))
        d.vnormals = _ptr(keep(scene.vnormals, np.float32))
        d.mat_id = _ptr(keep(scene.mat_id, np.int32))
        M = scene.mats.shape[0]
        mats = (OrMaterial * M)()
        for m in range(M):
            for k in range(3):
                mats[m].kd[k] = float(scene.mats[m, k])
                mats[m].ks[k] = float(scene.mats[m, 3 + k])
            mats[m].ns = float(scene.mats[m, 6])
            mats[m].ni = float(scene.mats[m, 7])
            mats[m].illum = int(scene.mats[m, 8])
        self._keep.append(mats)
        d.n_mats = M
        d.mats = C.cast(mats, C.c_void_p)
        mm = np.ascontiguousarray(scene.model_matrix, np.float32).reshape(-1)
        for k in range(12):
            d.model_matrix[k] = float(mm[k])
        if scene.spheres is not None and len(scene.spheres):
            d.n_spheres = int(scene.spheres.shape[0])
            d.spheres = _ptr(keep(scene.spheres, np.float32))
            d.sphere_mat = _ptr(keep(scene.sphere_mat, np.int32))
        self.handle = L.or_scene_create(C.byref(d), C.byref(self.params))

    def close(self):
        if self.handle:
            lib().or_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def camera(eye, view_inv, viewport, fovy, aspect) -> OrCamera:
        cam = OrCamera()
        for k in range(3):
            cam.eye[k] = float(eye[k])
        vi = np.asarray(view_inv, np.float32).reshape(-1)
        for k in range(12):
            cam.view_inv[k] = float(vi[k])
        for k in range(4):
            cam.viewport[k] = float(viewport[k])
        cam.fovy = float(fovy)
        cam.aspect = float(aspect)
        return cam

    def scene_camera(self) -> OrCamera:
        s = self.scene
        return self.camera(s.eye, s.view_inv, s.viewport, s.fovy, s.aspect)

    def root_box(self):
        mn = np.zeros(3, np.float32)
        mx = np.zeros(3, np.float32)
        lib().or_scene_root_box(self.handle, _ptr(mn), _ptr(mx))
        return mn, mx

    def octree_stats(self):
        out = np.zeros(4, np.int64)
        lib().or_scene_octree_stats(self.handle, _ptr(out))
        return out

    def render_pixels(self, cam: OrCamera, lights, pxy, threads=8):
        pxy = np.ascontiguousarray(pxy, np.int32)
        n = pxy.shape[0]
        px = np.ascontiguousarray(pxy[:, 0])
        py = np.ascontiguousarray(pxy[:, 1])
        lights = np.ascontiguousarray(lights, np.float32).reshape(-1, 3)
        rgb = np.zeros((n, 3), np.float32)
        face = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        rgb8 = np.zeros((n, 3), np.uint8)
        lib().or_render_pixels(self.handle, C.byref(cam), _ptr(lights), lights.shape[0], _ptr(px), _ptr(py),
                               n, _ptr(rgb), _ptr(face), _ptr(t), _ptr(rgb8), int(threads))
        return rgb, face, t, rgb8

    def render(self, cam: OrCamera, lights, width, height, stride=1, offx=0, offy=0, threads=8):
        xs = np.arange(offx, width, stride, dtype=np.int32)
        ys = np.arange(offy, height, stride, dtype=np.int32)
        pxy = np.stack(np.meshgrid(xs, ys, indexing="ij"), -1).reshape(-1, 2)
        return (pxy,) + self.render_pixels(cam, lights, pxy, threads)

    def census(self, reset=False):
        out = np.zeros(3, np.int64)
        lib().or_census_get(_ptr(out))
        if reset:
            lib().or_census_reset()
        return out


def quantize(rgb: np.ndarray) -> np.ndarray:
    """ppmIO.hpp:145 quantiser, vectorised: min(255, (int)(255*c)) with float32 multiply."""
    v = (np.float32(255) * rgb.astype(np.float32)).astype(np.float32)
    q = np.trunc(v).astype(np.int64)
    return np.minimum(255, q)
