"""ctypes binding of the C ABI (include/rt_api.h) exported by librt_b200.so.

This is plumbing for tests and bench.py: every compute call goes through the C ABI into the CUDA
kernels.  There is no CPU path here -- if the shared library is missing, or no CUDA device is
present, the calls raise RtError.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# RT_B200_LIB: developer override for A/B runs of two builds of the same C ABI
LIB_PATH = os.environ.get("RT_B200_LIB") or os.path.join(HERE, "librt_b200.so")

RT_MAX_LIGHTS = 25


class RtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rt error {code}: {msg}")
        self.code = code


class RtMaterial(C.Structure):
    _fields_ = [("kd", C.c_float * 3), ("ks", C.c_float * 3), ("ns", C.c_float), ("ni", C.c_float),
                ("illum", C.c_int32)]


class RtSceneDesc(C.Structure):
    _fields_ = [("n_faces", C.c_int32), ("verts", C.c_void_p), ("face_normals", C.c_void_p),
                ("vertex_normals", C.c_void_p), ("material_id", C.c_void_p), ("n_materials", C.c_int32),
                ("materials", C.c_void_p), ("model_matrix", C.c_float * 12), ("n_spheres", C.c_int32),
                ("spheres", C.c_void_p), ("sphere_material", C.c_void_p)]


class RtCamera(C.Structure):
    _fields_ = [("eye", C.c_float * 3), ("view_inv", C.c_float * 12), ("viewport", C.c_float * 4),
                ("fovy", C.c_float), ("aspect", C.c_float)]


class RtLights(C.Structure):
    _fields_ = [("n", C.c_int32), ("pos", C.c_void_p), ("color", C.c_float * 3)]


class RtParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("area_light", C.c_int32),
                ("point_light", C.c_int32), ("max_depth", C.c_int32), ("usteps", C.c_int32),
                ("vsteps", C.c_int32), ("area_len_x", C.c_float), ("area_len_y", C.c_float),
                ("band_rows", C.c_int32), ("band_rank", C.c_int32), ("band_world", C.c_int32),
                ("out_full_frame", C.c_int32), ("sphere_seed", C.c_uint32), ("sphere_radius", C.c_float)]


class RtStats(C.Structure):
    _fields_ = [("rays_primary", C.c_int64), ("rays_shadow", C.c_int64), ("rays_secondary", C.c_int64),
                ("pixels", C.c_int64), ("levels", C.c_int32), ("ms_total", C.c_float), ("ms_trace", C.c_float),
                ("ms_shadow", C.c_float), ("ms_shade", C.c_float), ("kernel_launches", C.c_int32),
                ("box_tests", C.c_int64), ("tri_tests", C.c_int64), ("shade_samples", C.c_int64),
                ("box_tests_shadow", C.c_int64), ("tri_tests_shadow", C.c_int64),
                ("filter_checks", C.c_int64), ("filter_slow", C.c_int64), ("filter_rejects", C.c_int64),
                ("shadow_rays_traced", C.c_int64), ("fused", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/rt_api.h declares (tests check that the library exports all of them)
API_SYMBOLS = [
    "rt_api_version", "rt_init", "rt_shutdown", "rt_last_error", "rt_device_name", "rt_set_option",
    "rt_default_params", "rt_mesh_load_obj", "rt_mesh_desc", "rt_mesh_info", "rt_mesh_destroy",
    "rt_scene_create", "rt_scene_destroy", "rt_scene_root_box", "rt_scene_info", "rt_ref_octree_stats",
    "rt_scene_debug_bvh", "rt_bvh_check", "rt_scene_build_info",
    "rt_render", "rt_render_device", "rt_render_submit", "rt_render_wait", "rt_local_rows", "rt_local_row_map", "rt_shared_frame_create",
    "rt_shared_frame_open", "rt_shared_frame_close", "rt_device_copy_to_host", "rt_trace_rays",
    "rt_light_strikes", "rt_box_intersect", "rt_box_intersect_box", "rt_ray_triangle", "rt_octree_candidates",
    "rt_phong_shade", "rt_screen_to_world", "rt_light_samples", "rt_write_ppm", "rt_selftest_div3",
    "rt_init_devices", "rt_shared_frame_signal", "rt_shared_frame_wait", "rt_host_frame_open", "rt_host_frame_ptr",
    "rt_host_frame_barrier", "rt_host_frame_close", "rt_render_into_frame", "rt_multi_create", "rt_multi_destroy",
    "rt_multi_device_count", "rt_multi_scene", "rt_multi_render", "rt_multi_render_device",
]

_lib = None


def lib():
    """Load librt_b200.so (built in-tree by __graft_entry__.build()).  Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RtError(-100, f"{LIB_PATH} not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32p = C.c_void_p, C.c_int32, C.c_int64, C.c_void_p
    L.rt_api_version.restype = C.c_int
    L.rt_init.argtypes = [C.c_int]
    L.rt_last_error.restype = C.c_char_p
    L.rt_device_name.argtypes = [C.c_char_p, C.c_size_t]
    L.rt_set_option.argtypes = [C.c_char_p, C.c_int]
    L.rt_default_params.argtypes = [C.POINTER(RtParams)]
    L.rt_default_params.restype = None
    L.rt_mesh_load_obj.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.rt_mesh_desc.argtypes = [vp, C.POINTER(RtSceneDesc)]
    L.rt_mesh_info.argtypes = [vp, vp, vp, vp, vp]
    L.rt_mesh_destroy.argtypes = [vp]
    L.rt_mesh_destroy.restype = None
    L.rt_scene_create.argtypes = [C.POINTER(RtSceneDesc), C.POINTER(vp)]
    L.rt_scene_destroy.argtypes = [vp]
    L.rt_scene_destroy.restype = None
    L.rt_scene_root_box.argtypes = [vp, vp, vp]
    L.rt_scene_info.argtypes = [vp, vp, vp, vp, vp, vp]
    L.rt_scene_build_info.argtypes = [vp, vp]
    L.rt_scene_debug_bvh.argtypes = [vp, vp, i64, vp, i64]
    L.rt_ref_octree_stats.argtypes = [C.POINTER(RtSceneDesc), i32, vp]
    L.rt_bvh_check.argtypes = [C.POINTER(RtSceneDesc), i32, vp]
    L.rt_render.argtypes = [vp, C.POINTER(RtCamera), C.POINTER(RtLights), C.POINTER(RtParams), vp, vp, vp, vp, vp]
    L.rt_render_device.argtypes = [vp, C.POINTER(RtCamera), C.POINTER(RtLights), C.POINTER(RtParams), vp, vp, vp,
                                   vp, vp, vp]
    L.rt_render_submit.argtypes = [vp, C.POINTER(RtCamera), C.POINTER(RtLights), C.POINTER(RtParams), vp, C.POINTER(C.c_int)]
    L.rt_render_wait.argtypes = [vp, C.c_int]
    L.rt_shared_frame_create.argtypes = [C.c_size_t, C.POINTER(vp), vp]
    L.rt_shared_frame_open.argtypes = [vp, C.POINTER(vp)]
    L.rt_shared_frame_close.argtypes = [vp, C.c_int]
    L.rt_device_copy_to_host.argtypes = [vp, vp, C.c_size_t]
    L.rt_local_rows.argtypes = [C.POINTER(RtParams)]
    L.rt_local_row_map.argtypes = [C.POINTER(RtParams), vp]
    L.rt_trace_rays.argtypes = [vp, i64, f32p, f32p, C.POINTER(RtLights), C.POINTER(RtParams), vp, vp, vp]
    L.rt_light_strikes.argtypes = [vp, i64, f32p, C.POINTER(RtLights), vp]
    L.rt_box_intersect.argtypes = [vp, i64, f32p, f32p, vp]
    L.rt_box_intersect_box.argtypes = [vp, vp, i64, f32p, f32p, vp]
    L.rt_ray_triangle.argtypes = [vp, i64, f32p, f32p, vp, vp]
    L.rt_octree_candidates.argtypes = [vp, vp, vp, vp, i32]
    L.rt_phong_shade.argtypes = [vp, i64, f32p, f32p, vp, C.POINTER(RtLights), C.POINTER(RtParams), vp]
    L.rt_screen_to_world.argtypes = [C.POINTER(RtCamera), i64, f32p, vp]
    L.rt_light_samples.argtypes = [C.POINTER(RtParams), vp, vp]
    L.rt_write_ppm.argtypes = [C.c_char_p, vp, i32, i32, i32]
    L.rt_selftest_div3.argtypes = [i64, C.c_uint32, i32, C.POINTER(C.c_int64)]
    L.rt_init_devices.argtypes = [C.c_int, vp]
    L.rt_shutdown.restype = None
    L.rt_shared_frame_signal.argtypes = [vp, C.c_size_t, C.c_int, C.c_uint, vp]
    L.rt_shared_frame_wait.argtypes = [vp, C.c_size_t, C.c_int, C.c_uint, vp]
    L.rt_host_frame_open.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(vp)]
    L.rt_host_frame_ptr.argtypes = [vp]
    L.rt_host_frame_ptr.restype = vp
    L.rt_host_frame_barrier.argtypes = [vp, C.c_int]
    L.rt_host_frame_close.argtypes = [vp]
    L.rt_host_frame_close.restype = None
    L.rt_render_into_frame.argtypes = [vp, C.POINTER(RtCamera), C.POINTER(RtLights), C.POINTER(RtParams), vp]
    L.rt_multi_create.argtypes = [C.POINTER(RtSceneDesc), C.c_int, vp, C.POINTER(vp)]
    L.rt_multi_destroy.argtypes = [vp]
    L.rt_multi_destroy.restype = None
    L.rt_multi_device_count.argtypes = [vp]
    L.rt_multi_scene.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.rt_multi_render.argtypes = [vp, C.POINTER(RtCamera), C.POINTER(RtLights), C.POINTER(RtParams), vp, vp]
    L.rt_multi_render_device.argtypes = [vp, C.POINTER(RtCamera), C.POINTER(RtLights), C.POINTER(RtParams), C.POINTER(vp),
                                         C.POINTER(C.c_float)]
    _lib = L
    return L


def _check(rc):
    if rc < 0:
        raise RtError(rc, lib().rt_last_error().decode(errors="replace"))
    return rc


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def init(device: int = 0):
    _check(lib().rt_init(device))


def device_name() -> str:
    buf = C.create_string_buffer(256)
    _check(lib().rt_device_name(buf, 256))
    return buf.value.decode()


def set_option(key: str, value: int):
    _check(lib().rt_set_option(key.encode(), int(value)))


def make_params(width, height, area=0, point=1, max_depth=-1, grid=(5, 5), band_rows=8, band_rank=0,
                band_world=1, sphere_seed=1) -> RtParams:
    p = RtParams()
    lib().rt_default_params(C.byref(p))
    p.width, p.height = int(width), int(height)
    p.area_light, p.point_light = int(area), int(point)
    p.max_depth = int(max_depth)
    p.usteps, p.vsteps = int(grid[0]), int(grid[1])
    p.band_rows, p.band_rank, p.band_world = int(band_rows), int(band_rank), int(band_world)
    p.sphere_seed = int(sphere_seed)
    return p


def make_camera(eye, view_inv, viewport, fovy, aspect) -> RtCamera:
    cam = RtCamera()
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


def default_camera(width, height) -> RtCamera:
    """Flycamera reset state (tucano/utils/flycamera.hpp:76-86): eye (0,0,2) looking down -z,
    fovy 60, aspect w/h (src/flyscene.cpp:46-47)."""
    vi = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 2]], np.float32)
    return make_camera((0, 0, 2), vi, (0, 0, width, height), 60.0, np.float32(width) / np.float32(height))


class Lights:
    def __init__(self, pos, color=(1.0, 1.0, 0.0)):
        self.pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        self.c = RtLights()
        self.c.n = self.pos.shape[0]
        self.c.pos = _ptr(self.pos)
        for k in range(3):
            self.c.color[k] = float(color[k])


@dataclass
class Frame:
    rgba: np.ndarray          # [rows, W, 4] uint8
    face: np.ndarray | None   # [rows, W] int32
    t: np.ndarray | None      # [rows, W] float32
    rgb: np.ndarray | None    # [rows, W, 3] float32
    stats: dict | None


class Mesh:
    """Host-side OBJ/MTL load + bake (rt_mesh_load_obj)."""

    def __init__(self, obj_path: str):
        self.h = C.c_void_p()
        _check(lib().rt_mesh_load_obj(obj_path.encode(), C.byref(self.h)))
        self.desc = RtSceneDesc()
        _check(lib().rt_mesh_desc(self.h, C.byref(self.desc)))

    def arrays(self):
        T = self.desc.n_faces
        M = self.desc.n_materials

        def arr(ptr, n, dt):
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(dt)), shape=(n,)).copy() if n else np.zeros(0, dt)

        verts = arr(self.desc.verts, T * 9, C.c_float).reshape(T, 3, 3)
        fn = arr(self.desc.face_normals, T * 3, C.c_float).reshape(T, 3)
        vn = arr(self.desc.vertex_normals, T * 9, C.c_float).reshape(T, 3, 3)
        mid = arr(self.desc.material_id, T, C.c_int32)
        mats = np.zeros((M, 9), np.float32)
        mp = C.cast(self.desc.materials, C.POINTER(RtMaterial))
        for m in range(M):
            mats[m] = list(mp[m].kd) + list(mp[m].ks) + [mp[m].ns, mp[m].ni, mp[m].illum]
        return verts, fn, vn, mid, mats

    def info(self):
        c = np.zeros(3, np.float32)
        r = C.c_float()
        s = C.c_float()
        nv = C.c_int32()
        _check(lib().rt_mesh_info(self.h, _ptr(c), C.byref(r), C.byref(s), C.byref(nv)))
        return c, r.value, s.value, nv.value

    def close(self):
        if self.h:
            lib().rt_mesh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Scene:
    """Device-resident scene (rt_scene_create): flattened BVH + triangle soup + shading tables."""

    def __init__(self, verts, fnormals, vnormals, mat_id, mats, model_matrix=None, spheres=None, sphere_mat=None):
        d = self._make_desc(verts, fnormals, vnormals, mat_id, mats, model_matrix, spheres, sphere_mat)
        self.n_faces = d.n_faces
        self.h = C.c_void_p()
        _check(lib().rt_scene_create(C.byref(d), C.byref(self.h)))

    def _make_desc(self, verts, fnormals, vnormals, mat_id, mats, model_matrix=None, spheres=None, sphere_mat=None):
        self._keep = []

        def keep(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            self._keep.append(a)
            return a

        d = RtSceneDesc()
        verts = keep(verts, np.float32)
        d.n_faces = int(verts.reshape(-1, 9).shape[0]) if verts.size else 0
        d.verts = _ptr(verts)
        d.face_normals = _ptr(keep(fnormals, np.float32))
        d.vertex_normals = _ptr(keep(vnormals, np.float32))
        d.material_id = _ptr(keep(mat_id, np.int32))
        mats = np.asarray(mats, np.float32).reshape(-1, 9)
        M = mats.shape[0]
        cm = (RtMaterial * M)()
        for m in range(M):
            for k in range(3):
                cm[m].kd[k] = float(mats[m, k])
                cm[m].ks[k] = float(mats[m, 3 + k])
            cm[m].ns, cm[m].ni, cm[m].illum = float(mats[m, 6]), float(mats[m, 7]), int(mats[m, 8])
        self._keep.append(cm)
        d.n_materials = M
        d.materials = C.cast(cm, C.c_void_p)
        mm = np.eye(4, dtype=np.float32)[:3] if model_matrix is None else np.asarray(model_matrix, np.float32)
        for k, v in enumerate(mm.reshape(-1)[:12]):
            d.model_matrix[k] = float(v)
        if spheres is not None and len(spheres):
            sp = keep(spheres, np.float32)
            d.n_spheres = int(sp.shape[0])
            d.spheres = _ptr(sp)
            d.sphere_material = _ptr(keep(sphere_mat, np.int32))
        return d

    @classmethod
    def from_mesh(cls, mesh: Mesh, spheres=None, sphere_mat=None):
        verts, fn, vn, mid, mats = mesh.arrays()
        return cls(verts, fn, vn, mid, mats, None, spheres, sphere_mat)

    def close(self):
        if getattr(self, "h", None):
            lib().rt_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def root_box(self):
        mn, mx = np.zeros(3, np.float32), np.zeros(3, np.float32)
        _check(lib().rt_scene_root_box(self.h, _ptr(mn), _ptr(mx)))
        return mn, mx

    def info(self):
        v = [C.c_int64() for _ in range(4)]
        ms = C.c_float()
        _check(lib().rt_scene_info(self.h, *[C.byref(x) for x in v], C.byref(ms)))
        return dict(nodes=v[0].value, leaves=v[1].value, prims=v[2].value, device_bytes=v[3].value, build_ms=ms.value)

    def build_info(self):
        """rt_scene_build_info: which builder made the acceleration structures and what came out."""
        o = np.zeros(16, np.int64)
        _check(lib().rt_scene_build_info(self.h, _ptr(o)))
        return dict(gpu=bool(o[0]), nodes=int(o[1]), leaves=int(o[2]), depth=int(o[3]), sah=o[4] / 1000.0,
                    octree=dict(leaves=int(o[5]), inner=int(o[6]), refs=int(o[7]), max_leaf=int(o[8])),
                    build_ms=o[9] / 1000.0, upload_ms=o[10] / 1000.0, octree_ms=o[11] / 1000.0, sort_ms=o[12] / 1000.0,
                    bvh_ms=o[13] / 1000.0, emit_ms=o[14] / 1000.0, octree_levels=int(o[15] // 1000), bvh_levels=int(o[15] % 1000))

    def debug_bvh(self):
        inf = self.info()
        nodes = np.zeros((inf["nodes"], 16), np.float32)
        tri = np.zeros(inf["prims"], np.int32)
        _check(lib().rt_scene_debug_bvh(self.h, _ptr(nodes), nodes.shape[0], _ptr(tri), tri.shape[0]))
        return nodes, tri

    def render(self, cam: RtCamera, lights: Lights, params: RtParams, want_face=True, want_t=True, want_rgb=True,
               want_stats=True, out_rgba=None) -> Frame:
        rows = lib().rt_local_rows(C.byref(params))
        W = params.width
        rgba = out_rgba if out_rgba is not None else np.zeros((rows, W, 4), np.uint8)
        face = np.zeros((rows, W), np.int32) if want_face else None
        t = np.zeros((rows, W), np.float32) if want_t else None
        rgb = np.zeros((rows, W, 3), np.float32) if want_rgb else None
        st = RtStats() if want_stats else None
        _check(lib().rt_render(self.h, C.byref(cam), C.byref(lights.c), C.byref(params), _ptr(rgba), _ptr(face),
                               _ptr(t), _ptr(rgb), C.byref(st) if st is not None else None))
        return Frame(rgba, face, t, rgb, st.as_dict() if st is not None else None)

    def submit(self, cam: RtCamera, lights: Lights, params: RtParams, out_rgba: np.ndarray) -> int:
        """Queue one frame and the copy of its RGBA into out_rgba (ideally pinned); returns a ticket.
        At most two frames in flight; call wait(ticket) before reading out_rgba."""
        t = C.c_int(-1)
        _check(lib().rt_render_submit(self.h, C.byref(cam), C.byref(lights.c), C.byref(params), _ptr(out_rgba), C.byref(t)))
        return t.value

    def wait(self, ticket: int) -> None:
        _check(lib().rt_render_wait(self.h, ticket))

    def render_device(self, cam, lights, params, d_rgba: int, d_face: int = 0, d_t: int = 0, d_rgb: int = 0,
                      stream: int = 0, stats: RtStats | None = None):
        """Outputs are raw device pointers (e.g. torch.Tensor.data_ptr())."""
        _check(lib().rt_render_device(self.h, C.byref(cam), C.byref(lights.c), C.byref(params), d_rgba or None,
                                      d_face or None, d_t or None, d_rgb or None, stream or None,
                                      C.byref(stats) if stats is not None else None))

    def render_into_frame(self, cam, lights, params, frame_ptr):
        """This rank's bands, copied to their global rows of the full host frame at frame_ptr (blocking)."""
        _check(lib().rt_render_into_frame(self.h, C.byref(cam), C.byref(lights.c), C.byref(params), C.c_void_p(frame_ptr)))

    def trace_rays(self, origins, dirs, lights: Lights, params: RtParams):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = o.shape[0]
        rgb = np.zeros((n, 3), np.float32)
        face = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        _check(lib().rt_trace_rays(self.h, n, _ptr(o), _ptr(d), C.byref(lights.c), C.byref(params), _ptr(rgb),
                                   _ptr(face), _ptr(t)))
        return rgb, face, t

    def light_strikes(self, hits, lights: Lights):
        h = np.ascontiguousarray(hits, np.float32).reshape(-1, 3)
        out = np.zeros((h.shape[0], lights.c.n), np.uint8)
        _check(lib().rt_light_strikes(self.h, h.shape[0], _ptr(h), C.byref(lights.c), _ptr(out)))
        return out

    def ray_triangle(self, origins, dirs, faces):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        f = np.ascontiguousarray(faces, np.int32)
        out = np.zeros(o.shape[0], np.float32)
        _check(lib().rt_ray_triangle(self.h, o.shape[0], _ptr(o), _ptr(d), _ptr(f), _ptr(out)))
        return out

    def octree_candidates(self, origin, dest):
        o = np.ascontiguousarray(origin, np.float32)
        d = np.ascontiguousarray(dest, np.float32)
        ids = np.zeros(max(1, self.n_faces), np.int32)
        n = _check(lib().rt_octree_candidates(self.h, _ptr(o), _ptr(d), _ptr(ids), ids.shape[0]))
        return ids[:n].copy()

    def phong_shade(self, origins, hits, faces, lights: Lights, params: RtParams):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        h = np.ascontiguousarray(hits, np.float32).reshape(-1, 3)
        f = np.ascontiguousarray(faces, np.int32)
        out = np.zeros((o.shape[0], 3), np.float32)
        _check(lib().rt_phong_shade(self.h, o.shape[0], _ptr(o), _ptr(h), _ptr(f), C.byref(lights.c), C.byref(params),
                                    _ptr(out)))
        return out

    def box_intersect(self, origins, dests):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dests, np.float32).reshape(-1, 3)
        out = np.zeros(o.shape[0], np.uint8)
        _check(lib().rt_box_intersect(self.h, o.shape[0], _ptr(o), _ptr(d), _ptr(out)))
        return out


def box_intersect_box(mn, mx, origins, dests):
    mn = np.ascontiguousarray(mn, np.float32)
    mx = np.ascontiguousarray(mx, np.float32)
    o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
    d = np.ascontiguousarray(dests, np.float32).reshape(-1, 3)
    out = np.zeros(o.shape[0], np.uint8)
    _check(lib().rt_box_intersect_box(_ptr(mn), _ptr(mx), o.shape[0], _ptr(o), _ptr(d), _ptr(out)))
    return out


def screen_to_world(cam: RtCamera, pixels_xy):
    p = np.ascontiguousarray(pixels_xy, np.float32).reshape(-1, 2)
    out = np.zeros((p.shape[0], 3), np.float32)
    _check(lib().rt_screen_to_world(C.byref(cam), p.shape[0], _ptr(p), _ptr(out)))
    return out


def bvh_check(verts, leaf_size=2):
    """Host-only BVH build + invariant check; returns dict(nodes, leaves, depth, max_leaf, refs, sah)."""
    v = np.ascontiguousarray(verts, np.float32)
    d = RtSceneDesc()
    d.n_faces = int(v.reshape(-1, 9).shape[0]) if v.size else 0
    d.verts = _ptr(v)
    out = np.zeros(6, np.int64)
    _check(lib().rt_bvh_check(C.byref(d), int(leaf_size), _ptr(out)))
    return dict(nodes=int(out[0]), leaves=int(out[1]), depth=int(out[2]), max_leaf=int(out[3]), refs=int(out[4]),
                sah=out[5] / 1000.0)


def selftest_div3(n_trials: int, seed: int, exp_range: int) -> int:
    bad = C.c_int64(-1)
    _check(lib().rt_selftest_div3(int(n_trials), int(seed), int(exp_range), C.byref(bad)))
    return bad.value


def light_samples(params: RtParams, light):
    l = np.ascontiguousarray(light, np.float32)
    out = np.zeros((25, 3), np.float32)
    n = _check(lib().rt_light_samples(C.byref(params), _ptr(l), _ptr(out)))
    return out[:n].copy()


class SharedFrame:
    """A full-frame RGBA buffer on rank 0's GPU that the other ranks map through CUDA IPC (NVLink)."""

    def __init__(self, width, height, handle: bytes | None = None):
        self.w, self.h = width, height
        self.ptr = C.c_void_p()
        self.owner = handle is None
        if self.owner:
            buf = (C.c_ubyte * 64)()
            _check(lib().rt_shared_frame_create(width * height * 4, C.byref(self.ptr), buf))
            self.handle = bytes(buf)
        else:
            buf = (C.c_ubyte * 64).from_buffer_copy(handle)
            _check(lib().rt_shared_frame_open(buf, C.byref(self.ptr)))
            self.handle = handle

    def signal(self, rank, seq, stream):
        _check(lib().rt_shared_frame_signal(self.ptr, self.w * self.h * 4, int(rank), int(seq), C.c_void_p(stream)))

    def wait(self, world, seq, stream):
        _check(lib().rt_shared_frame_wait(self.ptr, self.w * self.h * 4, int(world), int(seq), C.c_void_p(stream)))

    def to_host(self):
        out = np.zeros((self.h, self.w, 4), np.uint8)
        _check(lib().rt_device_copy_to_host(_ptr(out), self.ptr, out.nbytes))
        return out

    def close(self):
        if self.ptr:
            lib().rt_shared_frame_close(self.ptr, int(self.owner))
            self.ptr = C.c_void_p()


class HostFrame:
    """A full-frame RGBA buffer in POSIX shared memory, page-locked in every process that opens it
    (rt_host_frame_*): the per-GPU processes copy their own bands into it."""

    def __init__(self, name: str, width, height, create: bool):
        self.w, self.h = width, height
        self.hd = C.c_void_p()
        _check(lib().rt_host_frame_open(name.encode(), width * height * 4, int(create), C.byref(self.hd)))
        ptr = lib().rt_host_frame_ptr(self.hd)
        self.ptr = ptr
        self.array = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(height, width, 4))

    def barrier(self, world):
        _check(lib().rt_host_frame_barrier(self.hd, int(world)))

    def close(self):
        if self.hd:
            self.array = None
            lib().rt_host_frame_close(self.hd)
            self.hd = C.c_void_p()


class Multi:
    """One host process, N GPUs (rt_multi_*): large scenes are built by every device itself, all at once; small ones are baked once on the host and uploaded."""

    def __init__(self, devices, verts, fnormals, vnormals, mat_id, mats, model_matrix=None, spheres=None, sphere_mat=None):
        d = Scene._make_desc(self, verts, fnormals, vnormals, mat_id, mats, model_matrix, spheres, sphere_mat)
        self.devices = np.ascontiguousarray(devices, np.int32)
        self.h = C.c_void_p()
        _check(lib().rt_multi_create(C.byref(d), len(self.devices), _ptr(self.devices), C.byref(self.h)))

    def render(self, cam, lights: Lights, params: RtParams, out=None, want_stats=False):
        H, W = params.height, params.width
        if out is None:
            out = np.zeros((H, W, 4), np.uint8)
        st = RtStats() if want_stats else None
        _check(lib().rt_multi_render(self.h, C.byref(cam), C.byref(lights.c), C.byref(params), _ptr(out),
                                     C.byref(st) if st is not None else None))
        return (out, st.as_dict()) if want_stats else out

    def render_device(self, cam, lights: Lights, params: RtParams):
        """Returns (device-0 pointer of the assembled frame, device ms of the slowest device)."""
        ptr = C.c_void_p()
        ms = C.c_float()
        _check(lib().rt_multi_render_device(self.h, C.byref(cam), C.byref(lights.c), C.byref(params), C.byref(ptr), C.byref(ms)))
        return ptr, ms.value

    def close(self):
        if getattr(self, "h", None):
            lib().rt_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def local_row_map(params: RtParams):
    rows = lib().rt_local_rows(C.byref(params))
    out = np.zeros(rows, np.int32)
    _check(lib().rt_local_row_map(C.byref(params), _ptr(out)))
    return out


def write_ppm(path: str, rgba: np.ndarray, binary=False):
    rgba = np.ascontiguousarray(rgba, np.uint8)
    _check(lib().rt_write_ppm(path.encode(), _ptr(rgba), rgba.shape[1], rgba.shape[0], int(binary)))
