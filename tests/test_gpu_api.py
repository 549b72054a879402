"""GPU tests of the batched per-function entry points of the C ABI against the oracle, of band
sharding (multi-GPU partitioning on one device) and of size-independent properties at full size."""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

from conftest import case_params, load_golden, scene_arrays

pytestmark = pytest.mark.gpu


def _setup(case, pkg, O, scene_dir, **okw):
    g = load_golden(case)
    arrs = scene_arrays(case, pkg, scene_dir)
    cp = case_params(g)
    capi = pkg.capi
    capi.init(0)
    scene = capi.Scene(*arrs, g["model_matrix"])
    orc = O.Oracle(O.BakedScene(*arrs, g["model_matrix"]), area=cp["area"], point=cp["point"], max_depth=cp["max_depth"],
                   grid=cp["grid"], light_color=g["light_color"], **okw)
    return g, cp, scene, orc


def test_box_intersect_matches_reference_semantics(pkg, oracle_mod, scene_dir):
    O = oracle_mod
    g, cp, scene, orc = _setup("cube_point_1000", pkg, O, scene_dir)
    rng = np.random.default_rng(7)
    n = 20000
    o = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
    d = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
    # edge cases: zero direction components (inf / NaN paths), origin on the box planes
    d[:500, 0] = 0; d[500:1000, 1] = 0; d[1000:1500] = 0
    o[1500:2000, 0] = g["root_min"][0]; d[1500:2000, 0] = 0
    dest = (o + d).astype(np.float32)
    got = scene.box_intersect(o, dest)
    mn, mx = scene.root_box()
    assert (mn == g["root_min"]).all() and (mx == g["root_max"]).all()
    exp = np.array([O.lib().or_box_intersect(mn.ctypes.data, mx.ctypes.data, o[i].ctypes.data, dest[i].ctypes.data)
                    for i in range(n)], np.uint8)
    assert (got == exp).all()


def test_screen_to_world_bit_exact(pkg, oracle_mod):
    O = oracle_mod
    capi = pkg.capi
    capi.init(0)
    g = load_golden("cube_rot_500x400")  # rotated + translated camera
    cam = capi.make_camera(g["eye"], g["view_inv"], g["viewport"], float(g["cam"][0]), float(g["cam"][1]))
    ocam = O.Oracle.camera(g["eye"], g["view_inv"], g["viewport"], float(g["cam"][0]), float(g["cam"][1]))
    xs, ys = np.meshgrid(np.arange(0, 500, 7), np.arange(0, 400, 5))
    pix = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.float32)
    got = capi.screen_to_world(cam, pix)
    exp = np.zeros_like(got)
    for k in range(len(pix)):
        O.lib().or_screen_to_world(ocam, float(pix[k, 0]), float(pix[k, 1]), exp[k].ctypes.data)
    assert (got.view(np.uint32) == exp.view(np.uint32)).all()


@pytest.mark.parametrize("case", ["gallery_area_200x150", "dodge_point_1000"])
def test_trace_rays_arbitrary_rays(case, pkg, oracle_mod, scene_dir):
    """Flyscene::traceRay for rays that do not come from the camera (incoherent origins/directions)."""
    O = oracle_mod
    g, cp, scene, orc = _setup(case, pkg, O, scene_dir)
    rng = np.random.default_rng(11)
    n = 4000
    o = rng.uniform(-1.5, 1.5, (n, 3)).astype(np.float32)
    tgt = rng.uniform(-0.6, 0.6, (n, 3)).astype(np.float32)
    d = ((tgt - o) * rng.uniform(0.2, 3.0, (n, 1))).astype(np.float32)
    lights = pkg.capi.Lights(g["lights"], g["light_color"])
    params = pkg.capi.make_params(8, 8, cp["area"], cp["point"], cp["max_depth"], cp["grid"])
    rgb, face, t = scene.trace_rays(o, d, lights, params)
    L = np.ascontiguousarray(g["lights"], np.float32)
    e_rgb = np.zeros((n, 3), np.float32); e_face = np.zeros(n, np.int32); e_t = np.zeros(n, np.float32)
    for k in range(n):
        O.lib().or_trace_ray(orc.handle, o[k].ctypes.data, d[k].ctypes.data, 0, L.ctypes.data, L.shape[0],
                             e_rgb[k].ctypes.data, e_face[k:].ctypes.data, e_t[k:].ctypes.data)
    assert (face == e_face).all()
    assert (t.view(np.uint32) == e_t.view(np.uint32)).all()
    same = (rgb.view(np.uint32) == e_rgb.view(np.uint32)).all(1)
    close = np.isclose(rgb, e_rgb, rtol=1e-5, atol=1e-6, equal_nan=True).all(1)
    assert close.all(), f"{(~close).sum()} rays differ"
    print(f"{case}: {n} arbitrary rays, float RGB bit-identical for {same.mean():.5f}")


def test_light_strikes(pkg, oracle_mod, scene_dir):
    O = oracle_mod
    g, cp, scene, orc = _setup("gallery_area_200x150", pkg, O, scene_dir)
    rng = np.random.default_rng(3)
    n = 5000
    hits = rng.uniform(-0.7, 0.7, (n, 3)).astype(np.float32)
    lights_np = np.array([[-1, 1.2, 1.5], [1.5, 1.0, 1.0], [0.0, 0.1, 0.2]], np.float32)
    got = scene.light_strikes(hits, pkg.capi.Lights(lights_np))
    exp = np.zeros((n, 3), np.uint8)
    for k in range(n):
        O.lib().or_light_strikes(orc.handle, hits[k].ctypes.data, lights_np.ctypes.data, 3, exp[k].ctypes.data)
    assert (got == exp).all()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_band_sharded_render_equals_full_frame(world, pkg, scene_dir):
    """Multi-GPU partitioning, emulated on one device: the interleaved row bands rendered by each
    'rank' stitch to exactly the single-GPU frame (pixels are independent)."""
    capi = pkg.capi
    capi.init(0)
    g = load_golden("gallery_area_200x150")
    arrs = scene_arrays("gallery_area_200x150", pkg, scene_dir)
    scene = capi.Scene(*arrs)
    W, H = 160, 101  # odd height: the last band is partial
    cam = capi.default_camera(W, H)
    lights = capi.Lights(g["lights"])
    full = scene.render(cam, lights, capi.make_params(W, H, 1, 0, 2, (3, 3)))
    stitched = np.zeros_like(full.rgba)
    faces = np.full((H, W), -2, np.int32)
    for r in range(world):
        p = capi.make_params(W, H, 1, 0, 2, (3, 3), band_rows=8, band_rank=r, band_world=world)
        part = scene.render(cam, lights, p)
        rows = capi.local_row_map(p)
        assert part.rgba.shape[0] == len(rows)
        stitched[rows] = part.rgba
        faces[rows] = part.face
    assert (stitched == full.rgba).all()
    assert (faces == full.face).all()


def test_full_size_properties_1m_triangles(pkg, scene_dir):
    """BASELINE configs[2] at full size (999 698 triangles, 1920x1080): properties that do not need
    the oracle -- determinism, band invariance, background exactly where nothing is hit, t > 1e-5,
    face ids in range, shadowed pixels black."""
    capi = pkg.capi
    capi.init(0)
    arrs = scene_arrays("hf707_point_1920x1080_s20", pkg, scene_dir)
    scene = capi.Scene(*arrs)
    W, H = 1920, 1080
    cam = capi.default_camera(W, H)
    lights = capi.Lights(np.array([[-1, 1, 1]], np.float32))
    p = capi.make_params(W, H, 0, 1, 0)
    a = scene.render(cam, lights, p)
    b = scene.render(cam, lights, p)
    assert (a.rgba == b.rgba).all() and (a.face == b.face).all()  # deterministic despite atomics
    miss = a.face < 0
    assert (a.rgba[miss][:, :3] == 255).all()
    assert (a.t[~miss] > 1e-5).all() and (a.face[~miss] < arrs[0].shape[0]).all()
    assert a.stats["rays_primary"] == W * H and a.stats["rays_secondary"] == 0
    assert a.stats["rays_shadow"] == int((~miss).sum())
    # a 4-way band split reproduces the frame
    st = np.zeros_like(a.rgba)
    for r in range(4):
        pr = capi.make_params(W, H, 0, 1, 0, band_rows=8, band_rank=r, band_world=4)
        st[capi.local_row_map(pr)] = scene.render(cam, lights, pr, want_face=False, want_t=False, want_rgb=False).rgba
    assert (st == a.rgba).all()


def test_limits_and_errors(pkg, scene_dir):
    capi = pkg.capi
    capi.init(0)
    g = load_golden("cube_point_1000")
    scene = capi.Scene(g["verts"], g["fnormals"], g["vnormals"], g["mat_id"], g["mats"])
    cam = capi.default_camera(32, 32)
    with pytest.raises(capi.RtError):  # > 25 lights (visibleLights[25])
        scene.render(cam, capi.Lights(np.zeros((26, 3), np.float32)), capi.make_params(32, 32))
    with pytest.raises(capi.RtError):  # 6 x 5 grid > 25 samples
        scene.render(cam, capi.Lights(np.zeros((1, 3), np.float32)), capi.make_params(32, 32, 1, 0, -1, (6, 5)))
    # empty scene and zero lights render (all background / all shadow)
    empty = capi.Scene(np.zeros((0, 3, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros((0, 3, 3), np.float32),
                       np.zeros(0, np.int32), g["mats"])
    fr = empty.render(cam, capi.Lights(np.array([[0, 0, 1]], np.float32)), capi.make_params(32, 32))
    assert (fr.rgba[..., :3] == 255).all() and (fr.face == -1).all()
    fr = scene.render(cam, capi.Lights(np.zeros((0, 3), np.float32)), capi.make_params(32, 32))
    assert (fr.rgba[fr.face >= 0][:, :3] == 0).all()  # no light visible -> SHADOW (src/flyscene.cpp:699-710)


def test_ray_triangle_pairs(pkg, oracle_mod, scene_dir):
    """Flyscene::rayTriangleIntersection for explicit (ray, face) pairs, incl. the -72 miss sentinel,
    degenerate (sliver) faces of dodgeColorTest and den == 0."""
    O = oracle_mod
    g, cp, scene, orc = _setup("dodge_point_1000", pkg, O, scene_dir)
    rng = np.random.default_rng(5)
    n = 20000
    T = g["verts"].shape[0]
    faces = rng.integers(0, T, n).astype(np.int32)
    faces[:200] = 11034  # collinear sliver named in SURVEY.md A.10
    cent = g["verts"][faces].mean(1)
    o = (cent + rng.normal(0, 0.5, (n, 3))).astype(np.float32)
    d = ((cent + rng.normal(0, 0.01, (n, 3)) - o) * rng.uniform(0.5, 2, (n, 1))).astype(np.float32)
    d[200:400] = 0  # den == 0
    got = scene.ray_triangle(o, d, faces)
    exp = np.array([O.lib().or_ray_triangle(orc.handle, o[k].ctypes.data, d[k].ctypes.data, int(faces[k])) for k in range(n)],
                   np.float32)
    assert (got.view(np.uint32) == exp.view(np.uint32)).all()
    assert (got == -72).any() and (got != -72).any()


def test_octree_candidates_match_reference_boxtree(pkg, oracle_mod, scene_dir):
    """BoxTree::intersect: candidate face sets for a handful of queries, incl. a ray lying in an octree
    split plane (zero direction component -> NaN slab tests in the reference)."""
    O = oracle_mod
    g, cp, scene, orc = _setup("dodge_point_1000", pkg, O, scene_dir)
    T = g["verts"].shape[0]
    rng = np.random.default_rng(9)
    queries = [((0, 0, 2), (0.1, 0.05, 1)), ((0, 0, 2), (0, 0, 1)), ((0, 0, 2), (0.3, 0, 1)), ((-1, 1, 1), (0.2, -0.1, 0.0))]
    for _ in range(6):
        queries.append((tuple(rng.uniform(-1.5, 1.5, 3)), tuple(rng.uniform(-0.5, 0.5, 3))))
    buf = np.zeros(T, np.int32)
    for o, dest in queries:
        o = np.array(o, np.float32); dest = np.array(dest, np.float32)
        got = scene.octree_candidates(o, dest)
        n = O.lib().or_octree_candidates(orc.handle, o.ctypes.data, dest.ctypes.data, buf.ctypes.data, T)
        exp = buf[:n]  # the oracle's walk starts with the root box test, like BoxTree::intersect
        assert len(got) == n and (got == exp).all(), (o, dest, len(got), n)


def test_phong_shade_batch(pkg, oracle_mod, scene_dir):
    """Flyscene::phongShade for explicit (origin, hit, face) triples, area-light mode with two lights."""
    O = oracle_mod
    g, cp, scene, orc = _setup("gallery_area_200x150", pkg, O, scene_dir)
    lights_np = np.ascontiguousarray(g["lights"], np.float32)
    # take real hit points from a render so that the triples are meaningful
    capi = pkg.capi
    cam = capi.make_camera(g["eye"], g["view_inv"], g["viewport"], float(g["cam"][0]), float(g["cam"][1]))
    fr = scene.render(cam, capi.Lights(lights_np, g["light_color"]), capi.make_params(cp["w"], cp["h"], cp["area"], cp["point"], 0, cp["grid"]))
    ys, xs = np.nonzero(fr.face >= 0)
    sel = np.random.default_rng(2).choice(len(xs), 3000, replace=False)
    ys, xs = ys[sel], xs[sel]
    pix = np.stack([xs, ys], 1).astype(np.float32)
    screen = capi.screen_to_world(cam, pix)
    o = np.tile(np.asarray(g["eye"], np.float32), (len(xs), 1))
    d = (screen - o).astype(np.float32)
    hit = (o + fr.t[ys, xs][:, None].astype(np.float32) * d).astype(np.float32)
    faces = fr.face[ys, xs].astype(np.int32)
    got = scene.phong_shade(o, hit, faces, capi.Lights(lights_np, g["light_color"]),
                            capi.make_params(8, 8, cp["area"], cp["point"], 0, cp["grid"]))
    exp = np.zeros_like(got)
    for k in range(len(xs)):
        O.lib().or_phong_shade(orc.handle, o[k].ctypes.data, hit[k].ctypes.data, int(faces[k]), lights_np.ctypes.data,
                               lights_np.shape[0], exp[k].ctypes.data)
    assert np.allclose(got, exp, rtol=2e-6, atol=1e-7)
    print(f"phong batch: bit-identical {(got.view(np.uint32) == exp.view(np.uint32)).all(1).mean():.5f}")


def test_box_intersect_arbitrary_box(pkg, oracle_mod):
    O = oracle_mod
    pkg.capi.init(0)
    rng = np.random.default_rng(13)
    mn = np.array([-0.3, -0.2, -0.5], np.float32); mx = np.array([0.4, 0.1, 0.0], np.float32)
    n = 5000
    o = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    dest = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    dest[:300, 1] = o[:300, 1]  # zero y direction
    got = pkg.capi.box_intersect_box(mn, mx, o, dest)
    exp = np.array([O.lib().or_box_intersect(mn.ctypes.data, mx.ctypes.data, o[i].ctypes.data, dest[i].ctypes.data)
                    for i in range(n)], np.uint8)
    assert (got == exp).all()


@pytest.mark.parametrize("area,point,depth", [(0, 1, 2), (1, 0, 3)])
def test_analytic_spheres_match_the_port(area, point, depth, pkg, oracle_mod, scene_dir):
    """BASELINE configs[3] adds analytic spheres, which the reference does not have (parity unpinned):
    the CUDA path is checked against the definition in oracle/rt_oracle.c (ray_sphere) -- mixed
    triangle + sphere scene, diffuse and mirror spheres, shadows cast by and onto spheres."""
    O = oracle_mod
    capi = pkg.capi
    capi.init(0)
    arrs = list(scene_arrays("hf32_point_256x144", pkg, scene_dir))
    base = arrs[4].shape[0]
    arrs[4] = np.concatenate([arrs[4], np.array([[0.8, 0.3, 0.2, 1, 1, 1, 30, 0, 2], [0.9, 0.9, 0.9, 1, 1, 1, 60, 0, 4],
                                                 [0.2, 0.9, 0.9, 1, 1, 1, 20, 0, 9]], np.float32)], 0)
    spheres = pkg.scenes.sphere_cloud(60, seed=4321)
    spheres[:, 3] *= 3.0            # visible at this resolution
    spheres[:, 2] = np.abs(spheres[:, 2]) * 0.6 + 0.15   # in front of the height field
    sphere_mat = (base + (np.arange(len(spheres)) % 3)).astype(np.int32)
    scene = capi.Scene(*arrs, None, spheres, sphere_mat)
    W, H = 320, 200
    lights_np = np.array([[-1, 1, 1.5], [0.8, 0.6, 2.0]], np.float32)
    fr = scene.render(capi.default_camera(W, H), capi.Lights(lights_np), capi.make_params(W, H, area, point, depth, (3, 3)))
    orc = O.Oracle(O.BakedScene(*arrs, spheres=spheres, sphere_mat=sphere_mat), area=area, point=point, max_depth=depth,
                   grid=(3, 3))
    cam = O.Oracle.camera((0, 0, 2), np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 2]], np.float32), (0, 0, W, H), 60.0,
                          np.float32(W) / np.float32(H))
    pxy, rgb, face, t, rgb8 = orc.render(cam, lights_np, W, H, stride=1, threads=8)
    px, py = pxy[:, 0], pxy[:, 1]
    T = arrs[0].shape[0]
    assert (face >= T).sum() > 500, "spheres should be visible"
    assert (fr.face[py, px] == face).all()
    assert (fr.t[py, px].view(np.uint32) == t.view(np.uint32)).all()
    exp = np.clip(O.quantize(rgb), 0, 255)
    err = np.abs(fr.rgba[py, px, :3].astype(np.int64) - exp).max(-1)
    assert err.max() <= 1 and (err == 0).mean() > 0.999, f"max err {err.max()}, exact {(err == 0).mean()}"


def test_plain_bvh_mode_is_the_exact_nearest_hit(pkg, oracle_mod, scene_dir):
    """rt_set_option("reference_candidates", 0): without the octree filter the BVH returns the nearest
    hit over ALL faces, i.e. the oracle in brute-force candidate mode -- including the centre row of the
    image where the reference itself has holes (rays in an octree split plane)."""
    O = oracle_mod
    capi = pkg.capi
    capi.init(0)
    case = "hf224_point_3840x2160_s24"
    g = load_golden(case)
    arrs = scene_arrays(case, pkg, scene_dir)
    capi.set_option("reference_candidates", 0)
    try:
        scene = capi.Scene(*arrs)
    finally:
        capi.set_option("reference_candidates", 1)
    W, H = 480, 270  # even height: row 135 has dir.y == 0
    cam = capi.default_camera(W, H)
    lights = capi.Lights(g["lights"])
    fr = scene.render(cam, lights, capi.make_params(W, H, 0, 1, 0))
    orc = O.Oracle(O.BakedScene(*arrs), area=0, point=1, max_depth=0, candidates=1)
    ocam = O.Oracle.camera((0, 0, 2), np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 2]], np.float32), (0, 0, W, H), 60.0,
                           np.float32(W) / np.float32(H))
    ys = np.array([100, 134, 135, 136, 200], np.int32)
    xs = np.arange(0, W, 1, dtype=np.int32)
    pxy = np.stack(np.meshgrid(xs, ys, indexing="ij"), -1).reshape(-1, 2)
    rgb, face, t, rgb8 = orc.render_pixels(ocam, g["lights"], pxy, threads=8)
    assert (fr.face[pxy[:, 1], pxy[:, 0]] == face).all()
    assert (fr.t[pxy[:, 1], pxy[:, 0]].view(np.uint32) == t.view(np.uint32)).all()
    # and with the filter on, the same rows match the faithful (octree) oracle instead
    scene2 = capi.Scene(*arrs)
    fr2 = scene2.render(cam, lights, capi.make_params(W, H, 0, 1, 0))
    orc2 = O.Oracle(O.BakedScene(*arrs), area=0, point=1, max_depth=0, candidates=0)
    rgb2, face2, t2, _ = orc2.render_pixels(ocam, g["lights"], pxy, threads=8)
    assert (fr2.face[pxy[:, 1], pxy[:, 0]] == face2).all()
    print(f"plain BVH vs filtered: {(face != face2).sum()} of {len(face)} sampled pixels differ (reference octree holes)")


def test_unbounded_depth_matches_bounded_when_cap_is_large(pkg, scene_dir):
    """max_depth < 0 (the reference's unbounded recursion, host-synchronised level loop) and a cap
    larger than the scene's natural depth (CUDA-graph path) must give the same frame."""
    capi = pkg.capi
    capi.init(0)
    g = load_golden("gallery_small_point_320x240")
    arrs = scene_arrays("gallery_small_point_320x240", pkg, scene_dir)
    scene = capi.Scene(*arrs)
    cam = capi.default_camera(160, 120)
    lights = capi.Lights(g["lights"])
    a = scene.render(cam, lights, capi.make_params(160, 120, 0, 1, -1))
    levels = a.stats["levels"]
    assert levels <= 8, levels
    b = scene.render(cam, lights, capi.make_params(160, 120, 0, 1, 8))
    assert (a.rgba == b.rgba).all()
    # repeated graph replays are stable
    for _ in range(3):
        c = scene.render(cam, lights, capi.make_params(160, 120, 0, 1, 8), want_stats=False)
        assert (c.rgba == b.rgba).all()


def test_submit_wait_pipeline_matches_blocking_render(pkg, scene_dir):
    """rt_render_submit / rt_render_wait (two frames in flight, copy of frame k overlapping the kernels
    of frame k+1) returns the same frames as the blocking call, in order, for a moving camera."""
    import torch
    capi = pkg.capi
    capi.init(0)
    g = load_golden("gallery_area_200x150")
    arrs = scene_arrays("gallery_area_200x150", pkg, scene_dir)
    scene = capi.Scene(*arrs)
    lights = capi.Lights(g["lights"])
    W, H = 200, 150
    p = capi.make_params(W, H, 1, 0, 3, (4, 4))
    cams = []
    for k in range(5):
        vi = np.array([[1, 0, 0, 0.05 * k], [0, 1, 0, -0.03 * k], [0, 0, 1, 2.0 + 0.1 * k]], np.float32)
        cams.append(capi.make_camera((0.05 * k, -0.03 * k, 2.0 + 0.1 * k), vi, (0, 0, W, H), 60.0, np.float32(W) / np.float32(H)))
    want = [scene.render(c, lights, p, want_face=False, want_t=False, want_rgb=False, want_stats=False).rgba.copy() for c in cams]
    assert any((want[0] != w).any() for w in want[1:])
    bufs = [torch.zeros((H, W, 4), dtype=torch.uint8).pin_memory().numpy() for _ in range(2)]
    got, tickets = [], []
    for k, c in enumerate(cams):
        if k >= 2:
            scene.wait(tickets[k - 2])
            got.append(bufs[k & 1].copy())
        tickets.append(scene.submit(c, lights, p, bufs[k & 1]))
    with pytest.raises(RuntimeError):
        scene.submit(cams[0], lights, p, np.zeros((H, W, 4), np.uint8))  # a third frame in flight
    scene.wait(tickets[-1])  # completes the older one too
    got.append(bufs[(len(cams) - 2) & 1].copy())
    got.append(bufs[(len(cams) - 1) & 1].copy())
    scene.wait(tickets[0])   # long finished: no-op
    for k in range(len(cams)):
        assert (got[k] == want[k]).all(), k
    scene.close()


def test_shared_reciprocal_division_is_ieee_exact(pkg):
    """normalized() in the shading kernels divides a 3-vector by its norm with ONE refined reciprocal
    (rt_device.cuh div3).  On the device: 2^31 random operand sets per exponent window, plus unrestricted
    bit patterns (zero, denormal, inf, NaN), every quotient compared bitwise with IEEE division."""
    capi = pkg.capi
    capi.init(0)
    for seed, exp_range in [(1, 2), (2, 20), (3, 59), (4, 70), (5, 0), (6, 126)]:
        assert capi.selftest_div3(1 << 31, seed, exp_range) == 0, (seed, exp_range)


@pytest.mark.parametrize("n_tri,leaf", [(1, 2), (2, 2), (2, 1), (3, 16), (12, 16)])
def test_scenes_that_fit_one_leaf(n_tri, leaf, pkg, oracle_mod):
    """A scene with no more primitives than the leaf size used to get a root pair node with an empty slot,
    whose inverted box passes the symmetric slab test: every ray then stopped at the bottom-of-stack
    sentinel before reaching the only leaf and the frame came out all BACKGROUND.  Rendered against the
    oracle: one triangle, two triangles, and the bundled cube with leaf_size 16 (12 triangles, one leaf)."""
    O = oracle_mod
    capi = pkg.capi
    capi.init(0)
    if n_tri == 12:
        g = load_golden("cube_point_1000")
        arrs = [g["verts"], g["fnormals"], g["vnormals"], g["mat_id"], g["mats"]]
    else:
        v = np.array([[[-0.8, -0.6, 0.0], [0.7, -0.5, 0.1], [0.0, 0.8, -0.1]],
                      [[-0.5, 0.2, 0.4], [0.6, 0.3, 0.5], [0.1, -0.7, 0.45]],
                      [[-0.9, 0.7, -0.3], [-0.2, 0.9, -0.2], [-0.6, 0.1, -0.35]]], np.float32)[:n_tri]
        e0, e1 = v[:, 1] - v[:, 0], v[:, 2] - v[:, 0]
        fn = np.cross(e0, e1)
        fn = (fn / np.linalg.norm(fn, axis=1, keepdims=True)).astype(np.float32)
        vn = np.repeat(fn[:, None, :], 3, axis=1).copy()
        arrs = [v, fn, vn, np.zeros(n_tri, np.int32), np.array([[0.7, 0.6, 0.5, 1, 1, 1, 20, 0, 2]], np.float32)]
    capi.set_option("leaf_size", leaf)
    try:
        scene = capi.Scene(*arrs)
    finally:
        capi.set_option("leaf_size", 2)
    W, H = 200, 150
    lights_np = np.array([[-1, 1, 1.5]], np.float32)
    fr = scene.render(capi.default_camera(W, H), capi.Lights(lights_np), capi.make_params(W, H, 1, 0, 2, (3, 3)))
    orc = O.Oracle(O.BakedScene(*arrs), area=1, point=0, max_depth=2, grid=(3, 3))
    cam = O.Oracle.camera((0, 0, 2), np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 2]], np.float32), (0, 0, W, H), 60.0,
                          np.float32(W) / np.float32(H))
    pxy, rgb, face, t, rgb8 = orc.render(cam, lights_np, W, H, stride=1, threads=4)
    px, py = pxy[:, 0], pxy[:, 1]
    assert (face >= 0).sum() > 1000, "the primitives should be visible"
    assert (fr.face[py, px] == face).all()
    assert (fr.t[py, px].view(np.uint32) == t.view(np.uint32)).all()
    err = np.abs(fr.rgba[py, px, :3].astype(np.int64) - np.clip(O.quantize(rgb), 0, 255)).max(-1)
    assert err.max() <= 1 and (err == 0).mean() > 0.999
    scene.close()


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0], [0, 1], [0, 1, 2, 3]])
def test_multi_gpu_frame_equals_single_gpu_frame(devices, pkg, scene_dir):
    """rt_multi_*: one host process, one worker thread per device, interleaved row bands.  The host frame (every
    device copies its own bands into it) and the device frame (kernels store into device 0's frame over peer
    access) must both equal the single-GPU frame byte for byte.  Listing device 0 several times runs the same
    band split on one GPU (what a single-GPU box can check); real device lists need that many GPUs."""
    if max(devices) >= _n_gpus():
        pytest.skip(f"needs {max(devices) + 1} GPUs")
    capi = pkg.capi
    capi.init(0)
    g = load_golden("gallery_area_200x150")
    arrs = (g["verts"], g["fnormals"], g["vnormals"], g["mat_id"], g["mats"])
    W, H = 333, 187   # odd sizes: partial tiles and a partial last band
    cam = capi.default_camera(W, H)
    lights = capi.Lights(g["lights"][:1], g["light_color"])
    single = capi.Scene(*arrs, g["model_matrix"])
    for (area, point, depth, grid, band_rows) in [(1, 0, 3, (4, 4), 0), (0, 1, -1, (5, 5), 8), (1, 0, 2, (3, 3), 24)]:
        params = capi.make_params(W, H, area, point, depth, grid)
        ref = single.render(cam, lights, params, want_face=False, want_t=False, want_rgb=False).rgba
        multi = capi.Multi(devices, *arrs, g["model_matrix"])
        params.band_rows = band_rows
        out, st = multi.render(cam, lights, params, want_stats=True)
        assert (out == ref).all(), f"host frame of {devices} differs from the single-GPU frame"
        assert st["rays_primary"] == W * H
        out2 = multi.render(cam, lights, params)       # the timed path (no stats)
        assert (out2 == ref).all()
        ptr, ms = multi.render_device(cam, lights, params)
        dev = np.zeros((H, W, 4), np.uint8)
        assert capi.lib().rt_device_copy_to_host(dev.ctypes.data, ptr, dev.nbytes) == 0
        assert (dev == ref).all(), f"device-resident frame of {devices} differs"
        assert ms > 0
        multi.close()
    single.close()


def test_render_into_frame_assembles_bands_on_the_host(pkg):
    """rt_render_into_frame (what each per-GPU process of a torchrun job calls): 3 emulated ranks copy their own
    bands into one host frame; the result equals rt_render's frame."""
    capi = pkg.capi
    capi.init(0)
    g = load_golden("cube_point_1000")
    scene = capi.Scene(g["verts"], g["fnormals"], g["vnormals"], g["mat_id"], g["mats"])
    W, H = 320, 203
    cam = capi.default_camera(W, H)
    lights = capi.Lights(np.array([[-1, 1, 1]], np.float32))
    ref = scene.render(cam, lights, capi.make_params(W, H, 1, 0, 3, (4, 4)), want_face=False, want_t=False, want_rgb=False).rgba
    frame = np.zeros((H, W, 4), np.uint8)
    for world, rows in [(3, 8), (2, 16), (1, 8)]:
        frame[:] = 0
        for r in range(world):
            scene.render_into_frame(cam, lights, capi.make_params(W, H, 1, 0, 3, (4, 4), rows, r, world), frame.ctypes.data)
        assert (frame == ref).all(), (world, rows)
    scene.close()


def test_shutdown_releases_workspaces_and_scenes_stay_usable(pkg):
    capi = pkg.capi
    capi.init(0)
    g = load_golden("cube_point_1000")
    scene = capi.Scene(g["verts"], g["fnormals"], g["vnormals"], g["mat_id"], g["mats"])
    cam = capi.default_camera(160, 90)
    lights = capi.Lights(np.array([[-1, 1, 1]], np.float32))
    params = capi.make_params(160, 90, 1, 0, 3, (4, 4))
    a = scene.render(cam, lights, params, want_face=False, want_t=False, want_rgb=False).rgba
    capi.lib().rt_shutdown()
    capi.init(0)
    b = scene.render(cam, lights, params, want_face=False, want_t=False, want_rgb=False).rgba
    assert (a == b).all()
    scene.close()


@pytest.mark.parametrize("chunks", [2, 3, 4])
def test_chunked_render_equals_single_launch(chunks, pkg):
    """rt_render of a plain scene (the bundled cube) renders the frame in row chunks so that the device->host copy of
    one chunk overlaps the rendering of the next ("render_chunks"): same frame, byte for byte, for chunk heights
    that are not multiples of the image height, for tiny images (no chunking) and for every depth mode."""
    capi = pkg.capi
    capi.init(0)
    g = load_golden("cube_point_1000")
    scene = capi.Scene(g["verts"], g["fnormals"], g["vnormals"], g["mat_id"], g["mats"])
    lights = capi.Lights(np.array([[-1, 1, 1]], np.float32))
    try:
        for (W, H, area, point, depth) in [(640, 363, 1, 0, 3), (501, 1001, 0, 1, -1), (320, 130, 1, 0, 3), (1920, 1080, 1, 0, 3)]:
            cam = capi.default_camera(W, H)
            params = capi.make_params(W, H, area, point, depth, (4, 4))
            capi.set_option("render_chunks", 1)
            ref = scene.render(cam, lights, params, want_face=False, want_t=False, want_rgb=False, want_stats=False).rgba
            capi.set_option("render_chunks", chunks)
            got = scene.render(cam, lights, params, want_face=False, want_t=False, want_rgb=False, want_stats=False).rgba
            assert (got == ref).all(), (W, H, chunks)
            full = scene.render(cam, lights, params)   # with face / t / float outputs: the unchunked path
            assert (full.rgba == ref).all()
    finally:
        capi.set_option("render_chunks", 2)
    scene.close()


@pytest.mark.parametrize("case", ["cube_point_1000", "hf32_point_256x144"])
def test_page_locked_host_frame_is_written_directly(case, pkg):
    """rt_render with a page-locked host frame: the kernels store the pixels straight into it ("host_direct", no
    device->host copy).  Same frame, byte for byte, as through the staging copy, on the fused path (cube) and on the
    wavefront path (height field), alone and together with the other outputs."""
    import torch
    capi = pkg.capi
    capi.init(0)
    g = load_golden(case)
    scene = capi.Scene(g["verts"], g["fnormals"], g["vnormals"], g["mat_id"], g["mats"])
    lights = capi.Lights(np.array([[-1, 1, 1]], np.float32))
    try:
        for (W, H, area, point, depth) in [(640, 363, 1, 0, 3), (501, 257, 0, 1, -1)]:
            cam = capi.default_camera(W, H)
            params = capi.make_params(W, H, area, point, depth, (4, 4))
            ref = scene.render(cam, lights, params)   # pageable numpy outputs: staged copies
            pinned = torch.full((H, W, 4), 7, dtype=torch.uint8).pin_memory().numpy()
            capi.set_option("host_direct", 1)
            got = scene.render(cam, lights, params, want_face=False, want_t=False, want_rgb=False, want_stats=False, out_rgba=pinned)
            assert got.rgba is pinned and (pinned == ref.rgba).all(), (case, W, H)
            pinned[:] = 9
            full = scene.render(cam, lights, params, out_rgba=pinned)
            assert (pinned == ref.rgba).all() and (full.face == ref.face).all() and (full.t.view(np.uint32) == ref.t.view(np.uint32)).all()
            capi.set_option("host_direct", 0)
            pinned[:] = 3
            scene.render(cam, lights, params, want_face=False, want_t=False, want_rgb=False, want_stats=False, out_rgba=pinned)
            assert (pinned == ref.rgba).all()
    finally:
        capi.set_option("host_direct", 1)
    scene.close()


@pytest.mark.parametrize("case", ["cube_rot_500x400", "dodge_area_rot_400x300", "hf32_point_256x144"])
def test_tile_order_does_not_change_the_frame(case, pkg, scene_dir):
    """Pixel tiles inside the screen rectangle of the scene's bounds are handed out first ("tile_order"): only the
    order of the work changes.  Same frame with and without, for rotated cameras, both frame paths, every tile shape
    and a band-sharded share."""
    from conftest import case_params, scene_arrays
    capi = pkg.capi
    capi.init(0)
    g = load_golden(case)
    verts, fn, vn, mid, mats = scene_arrays(case, pkg, scene_dir)
    scene = capi.Scene(verts, fn, vn, mid, mats, g["model_matrix"])
    cp = case_params(g)
    cam = capi.make_camera(g["eye"], g["view_inv"], g["viewport"], float(g["cam"][0]), float(g["cam"][1]))
    lights = capi.Lights(g["lights"], g["light_color"])
    try:
        for fused in (0, 1):
            capi.set_option("fused_frame", fused)
            for twl in (3, 4, 5):
                capi.set_option("tile_w_log2", twl)
                for world in (1, 3):
                    params = capi.make_params(cp["w"], cp["h"], cp["area"], cp["point"], cp["max_depth"], cp["grid"],
                                              8 if world > 1 else 0, 1 if world > 1 else 0, world)
                    capi.set_option("tile_order", 0); capi.set_option("tile_cull", 0)
                    a = scene.render(cam, lights, params)
                    capi.set_option("tile_cull", 1)  # tiles outside the scene's screen rectangle: background untraced
                    c = scene.render(cam, lights, params)
                    assert (a.rgba == c.rgba).all() and (a.face == c.face).all() and (a.t.view(np.uint32) == c.t.view(np.uint32)).all()
                    assert (a.rgb.view(np.uint32) == c.rgb.view(np.uint32)).all()
                    for order in (1, 2):  # the scene's rectangle first (2: also for direct host frames)
                        capi.set_option("tile_order", order)
                        b = scene.render(cam, lights, params)
                        assert (a.rgba == b.rgba).all() and (a.face == b.face).all(), (fused, twl, world, order)
                        assert (a.rgb.view(np.uint32) == b.rgb.view(np.uint32)).all()
    finally:
        capi.set_option("tile_order", 1); capi.set_option("tile_cull", 1); capi.set_option("tile_w_log2", 3); capi.set_option("fused_frame", 2)
    scene.close()


def test_tile_cull_with_unusual_cameras(pkg, scene_dir):
    """The tile rectangle must stay conservative (or switch itself off) for cameras the goldens do not cover: far away
    (the scene is a few tiles), close up (the scene is larger than the screen), eye inside the scene's bounds, scene
    behind the eye, rotated and sheared view matrices, an eye that is not the view matrix' translation, an offset
    viewport, a very wide field of view.  Frames with and without culling must be identical in every output."""
    from conftest import scene_arrays
    capi = pkg.capi
    capi.init(0)
    verts, fn, vn, mid, mats = scene_arrays("hf32_point_256x144", pkg, scene_dir)
    scene = capi.Scene(verts, fn, vn, mid, mats)
    lights = capi.Lights(np.array([[-1, 1, 1]], np.float32))
    W, H = 322, 187

    def rot(ax, ay):
        cx, sx, cy, sy = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay)
        return np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])

    cams = []
    for eye, R, fovy, vp, eye_off in [
            ((0, 0, 12), np.eye(3), 60.0, (0, 0, W, H), 0),          # far away
            ((0, 0.1, 0.6), np.eye(3), 60.0, (0, 0, W, H), 0),       # close up
            ((0, 0.0, 0.0), np.eye(3), 60.0, (0, 0, W, H), 0),       # inside the bounds
            ((0, 0, -3), np.eye(3), 60.0, (0, 0, W, H), 0),          # scene behind the eye
            ((1.5, 1.0, 2.0), rot(-0.4, 0.6), 45.0, (0, 0, W, H), 0),  # rotated
            ((0.3, 0.2, 2.5), rot(0.2, -0.3) @ np.array([[1, 0.2, 0], [0, 1, 0], [0, 0, 1.0]]), 70.0, (0, 0, W, H), 0),  # sheared
            ((0, 0, 2.0), np.eye(3), 60.0, (0, 0, W, H), 1),         # eye differs from the matrix' translation
            ((0, 0, 2.5), np.eye(3), 60.0, (13, -7, W + 40, H + 25), 0),  # offset / larger viewport
            ((0, 0, 1.2), rot(0.1, 0.0), 150.0, (0, 0, W, H), 0)]:   # very wide
        t = np.asarray(eye, np.float64) + (np.array([0.3, -0.2, 0.1]) if eye_off else 0)
        vi = np.concatenate([R, t.reshape(3, 1)], 1).astype(np.float32)
        cams.append(capi.make_camera(eye, vi, vp, fovy, np.float32(W) / np.float32(H)))
    try:
        for k, cam in enumerate(cams):
            for fused in (0, 1):
                capi.set_option("fused_frame", fused)
                params = capi.make_params(W, H, 0, 1, 1, (5, 5))
                capi.set_option("tile_cull", 0); capi.set_option("tile_order", 0)
                a = scene.render(cam, lights, params)
                capi.set_option("tile_cull", 1); capi.set_option("tile_order", 1)
                b = scene.render(cam, lights, params)
                assert (a.rgba == b.rgba).all() and (a.face == b.face).all(), (k, fused)
                assert (a.t.view(np.uint32) == b.t.view(np.uint32)).all() and (a.rgb.view(np.uint32) == b.rgb.view(np.uint32)).all()
    finally:
        capi.set_option("tile_cull", 1); capi.set_option("tile_order", 1); capi.set_option("fused_frame", 2)
    scene.close()
