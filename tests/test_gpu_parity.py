"""GPU parity tests proper: the CUDA path (through the C ABI) against the golden vectors produced
by the reference itself, and against the oracle on the same seeded inputs.

Bars (BASELINE.json north_star): primary-hit face ids bit-exact (except documented float-tie /
sliver pixels), final 8-bit RGB within +-1 LSB on >= 99.9 % of pixels, max error stated."""
from __future__ import annotations

import numpy as np
import pytest

from conftest import GENERATED, SCENE_OF, case_params, load_golden, scene_arrays

pytestmark = pytest.mark.gpu

ALL_CASES = list(SCENE_OF) + list(GENERATED)
# No exceptions: the BVH path filters its hits through the reference's octree candidate sets
# (host/ref_octree.hpp), so sliver phantom hits and split-plane NaN holes match the reference too.
MAX_FACE_MISMATCH = {}


def quant(rgb):
    v = (np.float32(255) * rgb.astype(np.float32)).astype(np.float32)
    with np.errstate(invalid="ignore"):
        q = np.trunc(np.nan_to_num(v, nan=0.0, posinf=3e9, neginf=-3e9)).astype(np.int64)
    return np.clip(np.minimum(255, q), 0, 255)


@pytest.fixture(params=[0, 1], ids=["wavefront", "fused"])
def frame_path(request, capi):
    """Every golden case is rendered twice: by the per-level wavefront kernels (csrc/rt_kernels.cuh) and as one
    persistent kernel (csrc/rt_frame.cuh).  The library's default picks between them by frame size."""
    capi.set_option("fused_frame", request.param)
    yield request.param
    capi.set_option("fused_frame", 2)


@pytest.mark.parametrize("case", ALL_CASES)
def test_render_matches_reference_golden(case, frame_path, pkg, capi, scene_dir):
    g = load_golden(case)
    verts, fn, vn, mid, mats = scene_arrays(case, pkg, scene_dir)
    capi.init(0)
    scene = capi.Scene(verts, fn, vn, mid, mats, g["model_matrix"])
    cp = case_params(g)
    cam = capi.make_camera(g["eye"], g["view_inv"], g["viewport"], float(g["cam"][0]), float(g["cam"][1]))
    lights = capi.Lights(g["lights"], g["light_color"])
    params = capi.make_params(cp["w"], cp["h"], cp["area"], cp["point"], cp["max_depth"], cp["grid"])
    fr = scene.render(cam, lights, params)
    J = len(g["lights"]) * (1 + (0 if cp["point"] else cp["grid"][0] * cp["grid"][1]))
    # the path asked for is the one that ran (0 = wavefront graph, 1 = fused kernel, 2 = wavefront, level by level)
    stepped = 0 if 0 <= cp["max_depth"] <= 8 else 2
    assert fr.stats["fused"] == (1 if frame_path and J <= 64 else stepped)
    px, py = g["pxy"][:, 0], g["pxy"][:, 1]

    face = fr.face[py, px]
    bad_face = int((face != g["face"]).sum())
    assert bad_face <= MAX_FACE_MISMATCH.get(case, 0), f"{bad_face} primary-hit ids differ from the reference"
    same = face == g["face"]
    t = fr.t[py, px]
    assert (t[same].view(np.uint32) == g["t"][same].view(np.uint32)).all(), "hit parameter t differs bitwise"

    ref8 = quant(g["rgb"])
    got8 = fr.rgba[py, px, :3].astype(np.int64)
    err = np.abs(got8 - ref8).max(-1)
    frac_ok = float((err <= 1).mean())
    exact = float((err == 0).mean())
    print(f"{case}: pixels {len(px)} face mismatches {bad_face} rgb8 exact {exact:.6f} within1 {frac_ok:.6f} "
          f"max err {int(err.max())} float-rgb bit-identical {(fr.rgb[py, px].view(np.uint32) == g['rgb'].view(np.uint32)).all(-1).mean():.6f}")
    assert frac_ok >= 0.999, f"only {frac_ok:.5f} of pixels within +-1 LSB (max err {int(err.max())})"
    scene.close()


@pytest.mark.parametrize("case", ["cube_area_640x360", "gallery_area_200x150", "hf32_point_256x144"])
def test_full_frame_matches_oracle(case, frame_path, pkg, capi, oracle_mod, scene_dir):
    """Every pixel of the frame (not only the golden subset) against the CPU oracle."""
    O = oracle_mod
    g = load_golden(case)
    verts, fn, vn, mid, mats = scene_arrays(case, pkg, scene_dir)
    cp = case_params(g)
    capi.init(0)
    scene = capi.Scene(verts, fn, vn, mid, mats, g["model_matrix"])
    cam = capi.make_camera(g["eye"], g["view_inv"], g["viewport"], float(g["cam"][0]), float(g["cam"][1]))
    lights = capi.Lights(g["lights"], g["light_color"])
    params = capi.make_params(cp["w"], cp["h"], cp["area"], cp["point"], cp["max_depth"], cp["grid"])
    fr = scene.render(cam, lights, params)

    orc = O.Oracle(O.BakedScene(verts, fn, vn, mid, mats, g["model_matrix"]), area=cp["area"], point=cp["point"],
                   max_depth=cp["max_depth"], grid=cp["grid"], light_color=g["light_color"])
    ocam = O.Oracle.camera(g["eye"], g["view_inv"], g["viewport"], float(g["cam"][0]), float(g["cam"][1]))
    orc.census(reset=True)
    pxy, rgb, face, t, rgb8 = orc.render(ocam, g["lights"], cp["w"], cp["h"], stride=1, threads=8)
    n_prim, n_shadow, n_sec = [int(x) for x in orc.census(reset=True)]
    px, py = pxy[:, 0], pxy[:, 1]
    assert (fr.face[py, px] == face).all()
    assert (fr.t[py, px].view(np.uint32) == t.view(np.uint32)).all()
    err = np.abs(fr.rgba[py, px, :3].astype(np.int64) - quant(rgb)).max(-1)
    assert (err <= 1).mean() >= 0.999, f"max err {err.max()}"
    # ray census (the "ray" of Mrays/s): the kernels' counts equal the reference-semantics census of the
    # oracle; illum 6/7 child rays the reference traces and discards are the documented exception
    print(f"{case}: full frame {len(px)} px, rgb8 exact {(err == 0).mean():.6f}, max err {int(err.max())}; census gpu "
          f"{fr.stats['rays_primary']}/{fr.stats['rays_shadow']}/{fr.stats['rays_secondary']} oracle {n_prim}/{n_shadow}/{n_sec}")
    assert fr.stats["rays_primary"] == n_prim == len(px)
    if case != "gallery_area_200x150":
        assert fr.stats["rays_shadow"] == n_shadow and fr.stats["rays_secondary"] == n_sec
    scene.close()


def test_known_answers_of_the_default_scene_on_gpu(pkg, capi):
    """SURVEY.md section 4 known answers of the reference's own 1000x1000 cube render (point and area
    light), reproduced by the CUDA path: background count, bounding box, pixel values, colour count, mean."""
    g = load_golden("cube_point_1000")
    capi.init(0)
    scene = capi.Scene(g["verts"], g["fnormals"], g["vnormals"], g["mat_id"], g["mats"])
    cam = capi.default_camera(1000, 1000)
    lights = capi.Lights(np.array([[-1, 1, 1]], np.float32))
    fr = scene.render(cam, lights, capi.make_params(1000, 1000, 0, 1, -1))
    img = fr.rgba[..., :3].astype(np.int64)
    bg = (img == 255).all(-1)
    assert int(bg.sum()) == 505791
    nb = np.argwhere(~bg)
    assert tuple(nb.min(0)) == (149, 149) and tuple(nb.max(0)) == (851, 851)
    assert tuple(img[500, 500]) == (223, 223, 216) and tuple(img[250, 400]) == (255, 255, 216) and tuple(img[160, 160]) == (236, 236, 216)
    assert len(np.unique(img.reshape(-1, 3), axis=0)) == 41
    assert np.allclose(img.reshape(-1, 3).mean(0), (239.143, 239.143, 235.726), atol=2e-3)
    assert set(np.unique(fr.face)) == {-1, 10, 11}
    # ray census of App. A.8: 494 209 hit pixels x (1 primary + 1 merged gate/sample ray + 1 child) + misses
    assert fr.stats["rays_primary"] == 1000000 and fr.stats["rays_shadow"] == 494209 and fr.stats["rays_secondary"] == 494209
    fa = scene.render(cam, lights, capi.make_params(1000, 1000, 1, 0, -1))
    ia = fa.rgba[..., :3].astype(np.int64)
    assert tuple(ia[500, 500]) == (232, 232, 216) and tuple(ia[250, 400]) == (239, 239, 216) and tuple(ia[160, 160]) == (220, 220, 216)
    assert np.allclose(ia.reshape(-1, 3).mean(0), (239.273, 239.273, 235.726), atol=2e-3)
    # (1 gate + 25 samples) per hit; the mirror children leave the convex cube and hit nothing (App. A.8: x28)
    assert fa.stats["rays_shadow"] == 494209 * 26 and fa.stats["rays_secondary"] == 494209


def test_headline_config_full_frame_vs_oracle(frame_path, pkg, capi, oracle_mod):
    """BASELINE configs[1] exactly as bench.py runs it (cube, 1920x1080, 4x4 area light, depth 3):
    every one of the 2 073 600 pixels against the oracle."""
    O = oracle_mod
    g = load_golden("cube_point_1000")
    arrs = (g["verts"], g["fnormals"], g["vnormals"], g["mat_id"], g["mats"])
    capi.init(0)
    scene = capi.Scene(*arrs)
    W, H = 1920, 1080
    lights_np = np.array([[-1, 1, 1]], np.float32)
    fr = scene.render(capi.default_camera(W, H), capi.Lights(lights_np), capi.make_params(W, H, 1, 0, 3, (4, 4)))
    orc = O.Oracle(O.BakedScene(*arrs), area=1, point=0, max_depth=3, grid=(4, 4))
    cam = O.Oracle.camera((0, 0, 2), np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 2]], np.float32), (0, 0, W, H), 60.0,
                          np.float32(W) / np.float32(H))
    pxy, rgb, face, t, rgb8 = orc.render(cam, lights_np, W, H, stride=1, threads=16)
    px, py = pxy[:, 0], pxy[:, 1]
    assert (fr.face[py, px] == face).all()
    assert (fr.t[py, px].view(np.uint32) == t.view(np.uint32)).all()
    err = np.abs(fr.rgba[py, px, :3].astype(np.int64) - quant(rgb)).max(-1)
    same_f32 = (fr.rgb[py, px].view(np.uint32) == rgb.view(np.uint32)).all(-1).mean()
    print(f"C2 full frame: {len(px)} px, rgb8 exact {(err == 0).mean():.7f}, max err {int(err.max())}, float RGB bit-identical {same_f32:.7f}")
    assert err.max() <= 1 and (err == 0).mean() >= 0.99999


@pytest.mark.parametrize("case,seed", [("gallery_area_200x150", 1), ("hf32_point_256x144", 77)])
def test_spherical_light_mode_matches_oracle(case, seed, pkg, capi, oracle_mod, scene_dir):
    """areaLight = pointLight = 0: the reference's 25-point spherical light with its random_device draws
    replaced by the documented counter hash (RtParams.sphere_seed).  The reference cannot pin this mode
    (no two of its runs agree); parity is against the oracle's restatement of :974-993."""
    O = oracle_mod
    g = load_golden(case)
    verts, fn, vn, mid, mats = scene_arrays(case, pkg, scene_dir)
    cp = case_params(g)
    W, H = min(cp["w"], 160), min(cp["h"], 120)
    capi.init(0)
    scene = capi.Scene(verts, fn, vn, mid, mats, g["model_matrix"])
    vp = (0, 0, W, H)
    aspect = float(np.float32(W) / np.float32(H))
    cam = capi.make_camera(g["eye"], g["view_inv"], vp, float(g["cam"][0]), aspect)
    lights = capi.Lights(g["lights"], g["light_color"])
    fr = scene.render(cam, lights, capi.make_params(W, H, 0, 0, 3, sphere_seed=seed))
    orc = O.Oracle(O.BakedScene(verts, fn, vn, mid, mats, g["model_matrix"]), area=0, point=0, max_depth=3,
                   light_color=g["light_color"], sphere_seed=seed)
    ocam = O.Oracle.camera(g["eye"], g["view_inv"], vp, float(g["cam"][0]), aspect)
    pxy, rgb, face, t, rgb8 = orc.render(ocam, g["lights"], W, H, stride=1, threads=8)
    px, py = pxy[:, 0], pxy[:, 1]
    assert (fr.face[py, px] == face).all()
    err = np.abs(fr.rgba[py, px, :3].astype(np.int64) - quant(rgb)).max(-1)
    print(f"{case} spherical: {len(px)} px, rgb8 exact {(err == 0).mean():.6f}, max err {int(err.max())}, "
          f"shadow rays {fr.stats['rays_shadow']}")
    assert (err <= 1).mean() >= 0.999, f"max err {err.max()}"
    other = scene.render(cam, lights, capi.make_params(W, H, 0, 0, 3, sphere_seed=seed + 1))
    assert (other.rgba != fr.rgba).any()  # the seed matters
    scene.close()


def test_c4_benchmark_config_vs_port(pkg, capi, oracle_mod):
    """BASELINE configs[3] EXACTLY as bench.py --workload c4 builds and renders it (100 352 triangles + 1000
    analytic spheres, 3840x2160, 4x4 area light, depth cap 5): every 16th pixel in x and y (32 400 pixels)
    against the CPU port.  The spheres have no reference implementation (parity unpinned for them);
    the triangle part of this configuration is pinned against the reference by the two hf224*_area_d5_g4
    goldens of test_render_matches_reference_golden."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    O = oracle_mod
    wl = bench.WORKLOADS["c4"]
    arrs, spheres, sphere_mat = bench.workload_arrays(wl)
    assert arrs[0].shape[0] == 100352 and len(spheres) == 1000
    W, H = wl["w"], wl["h"]
    capi.init(0)
    scene = capi.Scene(*arrs, None, spheres, sphere_mat)
    lights_np = np.array([[-1.0, 1.0, 1.0]], np.float32)
    params = capi.make_params(W, H, wl["area"], wl["point"], wl["max_depth"], wl["grid"])
    fr = scene.render(capi.default_camera(W, H), capi.Lights(lights_np), params)
    orc = O.Oracle(O.BakedScene(*arrs, spheres=spheres, sphere_mat=sphere_mat), area=wl["area"], point=wl["point"],
                   max_depth=wl["max_depth"], grid=wl["grid"])
    cam = O.Oracle.camera((0, 0, 2), np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 2]], np.float32), (0, 0, W, H), 60.0,
                          np.float32(W) / np.float32(H))
    pxy, rgb, face, t, rgb8 = orc.render(cam, lights_np, W, H, stride=16, offx=5, offy=3, threads=max(1, (os.cpu_count() or 2) - 1))
    px, py = pxy[:, 0], pxy[:, 1]
    T = arrs[0].shape[0]
    assert (face >= T).sum() > 100 and ((face >= 0) & (face < T)).sum() > 3000, "spheres and triangles should both be visible"
    assert (fr.face[py, px] == face).all()
    assert (fr.t[py, px].view(np.uint32) == t.view(np.uint32)).all()
    err = np.abs(fr.rgba[py, px, :3].astype(np.int64) - quant(rgb)).max(-1)
    same_f32 = (fr.rgb[py, px].view(np.uint32) == rgb.view(np.uint32)).all(-1).mean()
    print(f"C4 as benchmarked: {len(px)} px, sphere hits {(face >= T).sum()}, rgb8 exact {(err == 0).mean():.6f}, max err "
          f"{int(err.max())}, float RGB bit-identical {same_f32:.6f}, levels {fr.stats['levels']}, secondary rays {fr.stats['rays_secondary']}")
    assert err.max() <= 1 and (err == 0).mean() >= 0.9999
    scene.close()


@pytest.mark.parametrize("case", ["hf224_area_d5_g4_3840x2160_s24", "hf224m_area_d5_g4_3840x2160_s48", "hf707_point_1920x1080_s20",
                                  "gallery_area_200x150", "dodge_area_rot_400x300"])
@pytest.mark.parametrize("donate_min", [101, 108, 124])
def test_work_donation_gives_the_same_frame(case, donate_min, pkg, capi, scene_dir):
    """Trav::run_split (idle lanes of a warp take over stack entries of busy lanes) must find every nearest hit and
    decide every visibility exactly like the plain warp traversal: the goldens of the traversal-heavy scenes again,
    for several donation thresholds, against the frame rendered without donation and against the reference."""
    g = load_golden(case)
    verts, fn, vn, mid, mats = scene_arrays(case, pkg, scene_dir)
    capi.init(0)
    scene = capi.Scene(verts, fn, vn, mid, mats, g["model_matrix"])
    cp = case_params(g)
    cam = capi.make_camera(g["eye"], g["view_inv"], g["viewport"], float(g["cam"][0]), float(g["cam"][1]))
    lights = capi.Lights(g["lights"], g["light_color"])
    params = capi.make_params(cp["w"], cp["h"], cp["area"], cp["point"], cp["max_depth"], cp["grid"])
    try:
        capi.set_option("fused_frame", 0)
        capi.set_option("donate_min_lanes", 0)
        ref = scene.render(cam, lights, params)
        capi.set_option("donate_min_lanes", donate_min)
        fr = scene.render(cam, lights, params)
    finally:
        capi.set_option("donate_min_lanes", 12)
        capi.set_option("fused_frame", 2)
    assert (fr.rgba == ref.rgba).all() and (fr.face == ref.face).all()
    assert (fr.t.view(np.uint32) == ref.t.view(np.uint32)).all()
    assert (fr.rgb.view(np.uint32) == ref.rgb.view(np.uint32)).all()
    for k in ("rays_primary", "rays_shadow", "rays_secondary"):
        assert fr.stats[k] == ref.stats[k]
    px, py = g["pxy"][:, 0], g["pxy"][:, 1]
    assert (fr.face[py, px] == g["face"]).all()
    err = np.abs(fr.rgba[py, px, :3].astype(np.int64) - quant(g["rgb"])).max(-1)
    assert (err <= 1).mean() >= 0.999
    scene.close()
