"""Host-side logic of the product on CPU: OBJ/MTL bake against the scenes the reference actually
traced, the C-ABI surface, band sharding arithmetic, PPM writer, and the no-CPU-fallback rule."""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden


def _bits(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.uint32)
    b = np.ascontiguousarray(b, np.float32).view(np.uint32)
    return int((a != b).sum())


GEN_SCENES = {
    "gallery_small_point_320x240": ("write_gallery", (1,)),
    "gallery_area_200x150": ("write_gallery", (3,)),
    "hf32_point_256x144": ("write_heightfield", (32,)),
}


@pytest.mark.parametrize("case", list(GEN_SCENES))
def test_loader_bakes_the_scene_the_reference_traced(case, pkg, tmp_path):
    """rt_mesh_load_obj output == the reference's own world vertices / normals / materials, bitwise."""
    gen, args = GEN_SCENES[case]
    path = str(tmp_path / "scene.obj")
    getattr(pkg.scenes, gen)(path, *args)
    g = load_golden(case)
    mesh = pkg.capi.Mesh(path)
    verts, fn, vn, mid, mats = mesh.arrays()
    assert _bits(verts, g["verts"]) == 0
    assert _bits(fn, g["fnormals"]) == 0
    assert _bits(vn, g["vnormals"]) == 0
    assert (mid == g["mat_id"]).all()
    assert _bits(mats, g["mats"]) == 0
    c, r, s, nv = mesh.info()
    assert _bits(c, g["centroid"]) == 0 and np.float32(r) == g["radius"] and np.float32(s) == g["norm_scale"]
    assert nv == g["obj_verts"].shape[0]


@pytest.mark.parametrize("case,obj", [("cube_point_1000", "cube.obj"), ("dodge_point_1000", "dodgeColorTest.obj")])
def test_loader_on_bundled_scenes(case, obj, pkg):
    path = os.path.join(ROOT, "oracle", "_ref", "scenes", obj)
    if not os.path.exists(path):
        pytest.skip("bundled reference scenes not present (oracle/_ref is built from /root/reference)")
    g = load_golden(case)
    verts, fn, vn, mid, mats = pkg.capi.Mesh(path).arrays()
    assert _bits(verts, g["verts"]) == 0 and _bits(fn, g["fnormals"]) == 0 and _bits(vn, g["vnormals"]) == 0
    assert (mid == g["mat_id"]).all() and _bits(mats, g["mats"]) == 0


def test_loader_errors(pkg, tmp_path):
    with pytest.raises(pkg.capi.RtError) as e:
        pkg.capi.Mesh(str(tmp_path / "missing.obj"))
    assert e.value.code == -4
    # OBJ without mtllib: default material (reference: UB materials[-1]); empty file: zero faces
    p = tmp_path / "nomtl.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    m = pkg.capi.Mesh(str(p))
    verts, fn, vn, mid, mats = m.arrays()
    assert verts.shape == (1, 3, 3) and mid[0] == 0 and mats.shape[0] == 1
    assert np.allclose(fn[0], (0, 0, 1))
    e2 = tmp_path / "empty.obj"
    e2.write_text("# nothing\n")
    assert pkg.capi.Mesh(str(e2)).arrays()[0].shape[0] == 0


def test_loader_rejects_face_indices_outside_the_vertex_list(pkg, tmp_path):
    """OBJ ids that do not name an existing vertex used to index past the vertex / normal arrays (heap
    corruption); they must come back as RT_ERR_IO.  Negative ids are relative to the vertices read so far."""
    capi = pkg.capi
    tri = "v 0 0 0\nv 1 0 0\nv 0 1 0\n"
    for body in ("f 1 2 4\n", "f 0 1 2\n", "f 1 2 -4\n", "f 1 2 99999999\n"):
        p = tmp_path / "bad.obj"
        p.write_text(tri + body)
        with pytest.raises(capi.RtError) as e:
            capi.Mesh(str(p))
        assert e.value.code == -4 and "face index" in str(e.value)
    # a relative index resolves against the vertex count AT THAT LINE, not the final one
    p = tmp_path / "rel.obj"
    p.write_text(tri + "f -3 -2 -1\nv 5 5 5\nf 1 2 -1\n")
    a = capi.Mesh(str(p), ).arrays()
    q = tmp_path / "abs.obj"
    q.write_text(tri + "f 1 2 3\nv 5 5 5\nf 1 2 4\n")
    b = capi.Mesh(str(q)).arrays()
    assert a[0].shape == (2, 3, 3) and _bits(a[0], b[0]) == 0 and _bits(a[1], b[1]) == 0 and _bits(a[2], b[2]) == 0


def test_library_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "rt_api.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = C.CDLL(pkg.capi.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, f"librt_b200.so does not export {missing}"
    assert declared == set(pkg.capi.API_SYMBOLS)
    assert pkg.capi.lib().rt_api_version() == 3


def test_no_cpu_fallback(pkg):
    """Without a CUDA device every compute entry point must fail loudly (RT_ERR_NO_DEVICE)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    capi = pkg.capi
    with pytest.raises(capi.RtError) as e:
        capi.init(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    g = load_golden("cube_point_1000")
    with pytest.raises(capi.RtError) as e:
        capi.Scene(g["verts"], g["fnormals"], g["vnormals"], g["mat_id"], g["mats"])
    assert e.value.code == -2


def test_product_does_not_reference_the_oracle():
    """The product tree must never import, link or execute anything under oracle/."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "raytracer-in-cpp_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(base, f), errors="replace").read()
                if re.search(r"liboracle|rt_oracle\.h|from oracle|import oracle|oracle/_ref", txt):
                    bad.append(f)
    assert not bad, bad


def test_band_sharding_partitions_the_image(pkg):
    capi = pkg.capi
    for H, B, world in [(1080, 8, 8), (1080, 8, 3), (1000, 16, 4), (37, 8, 2), (5, 8, 4), (4320, 8, 8)]:
        seen = []
        for r in range(world):
            p = capi.make_params(64, H, band_rows=B, band_rank=r, band_world=world)
            rows = capi.local_row_map(p)
            assert len(rows) == capi.lib().rt_local_rows(C.byref(p))
            assert (np.diff(rows) > 0).all() if len(rows) > 1 else True
            assert all((row // B) % world == r for row in rows)
            seen.append(rows)
        allrows = np.sort(np.concatenate(seen))
        assert (allrows == np.arange(H)).all()
        if H >= B * world:
            sizes = [len(s) for s in seen]
            assert max(sizes) - min(sizes) <= B  # interleaving balances the load


def test_light_samples_host(pkg):
    capi = pkg.capi
    p = capi.make_params(10, 10, area=1, point=0)
    s = capi.light_samples(p, (-1, 1, 1))
    assert s.shape == (25, 3)
    assert np.allclose(s[0], (-0.0700000003, 0.114999995, 1), atol=1e-8)
    assert np.allclose(s[24], (-0.629999995, 1.03499997, 1), atol=1e-8)
    assert capi.light_samples(capi.make_params(10, 10, area=0, point=1), (3, 4, 5)).tolist() == [[3, 4, 5]]
    # spherical mode (src/flyscene.cpp:974-993) with deterministic draws: 25 points on the sphere of
    # radius R/5 around the light, a pure function of (seed, light)
    L = np.array((-1, 1, 1), np.float32)
    sp = capi.light_samples(capi.make_params(10, 10, area=0, point=0), L)
    assert sp.shape == (25, 3)
    assert np.allclose(np.linalg.norm(sp.astype(np.float64) - L, axis=1), 1.0000001 / 5, atol=2e-7)
    assert (sp == capi.light_samples(capi.make_params(10, 10, area=0, point=0), L)).all()
    assert (sp != capi.light_samples(capi.make_params(10, 10, area=0, point=0, sphere_seed=2), L)).any()
    assert len(np.unique(sp, axis=0)) == 25
    with pytest.raises(capi.RtError):
        capi.light_samples(capi.make_params(10, 10, area=1, point=0, grid=(6, 5)), (0, 0, 0))  # > 25 samples


def test_ppm_writer_matches_reference_text_layout(pkg, tmp_path):
    """writePPMImage (ppmIO.hpp:130-151): 'P3', 'W H', '255', then 'r g b ' per pixel, one line per row."""
    rgba = np.zeros((2, 3, 4), np.uint8)
    rgba[0, 0] = (255, 0, 7, 255)
    rgba[1, 2] = (12, 216, 100, 255)
    p = str(tmp_path / "o.ppm")
    pkg.capi.write_ppm(p, rgba)
    assert open(p).read() == "P3\n3 2\n255\n255 0 7 0 0 0 0 0 0 \n0 0 0 0 0 0 12 216 100 \n"
    pkg.capi.write_ppm(p, rgba, binary=True)
    raw = open(p, "rb").read()
    assert raw.startswith(b"P6\n3 2\n255\n") and raw[-3:] == bytes([12, 216, 100])


@pytest.mark.parametrize("case", ["cube_point_1000", "dodge_point_1000", "gallery_area_200x150", "hf32_point_256x144"])
def test_reference_octree_shape(case, pkg):
    """The product's re-derivation of the reference BoxTree (candidate filter) has the shape the
    reference built: reachable leaves, inner nodes, face references, largest leaf."""
    g = load_golden(case)
    capi = pkg.capi
    d = capi.RtSceneDesc()
    verts = np.ascontiguousarray(g["verts"], np.float32)
    d.n_faces = verts.shape[0]
    d.verts = verts.ctypes.data
    out = np.zeros(4, np.int64)
    assert capi.lib().rt_ref_octree_stats(C.byref(d), 1000, out.ctypes.data) == 0
    assert (out == g["octree_stats"]).all(), (out, g["octree_stats"])


def test_spherical_light_samples_library_equals_oracle(pkg, oracle_mod):
    """The deterministic replacement of the reference's random spherical light: the library's host code
    and the oracle restate the same expression (src/flyscene.cpp:981-990) -> identical floats."""
    import ctypes as C
    O = oracle_mod
    capi = pkg.capi
    for seed in (1, 2, 12345, 0xFFFFFFFF):
        for light in ((-1, 1, 1), (0.25, -3.5, 2.0)):
            l = np.array(light, np.float32)
            got = capi.light_samples(capi.make_params(8, 8, area=0, point=0, sphere_seed=seed), l)
            p = O.OrParams()
            O.lib().or_default_params(C.byref(p))
            p.area_light, p.point_light, p.sphere_seed = 0, 0, seed
            want = np.zeros((25, 3), np.float32)
            assert O.lib().or_light_samples(C.byref(p), l.ctypes.data, want.ctypes.data) == 25
            assert (got.view(np.uint32) == want.view(np.uint32)).all()


def test_scene_validation_precedes_device_work(pkg):
    """Bad scene descriptions are rejected with a message before any CUDA call (so also on a CPU-only box):
    non-finite geometry, material ids outside the table, negative sphere radii."""
    from conftest import load_golden
    capi = pkg.capi
    g = load_golden("cube_point_1000")
    good = (g["verts"], g["fnormals"], g["vnormals"], g["mat_id"], g["mats"])
    v = g["verts"].copy()
    v[3, 1, 2] = np.nan
    with pytest.raises(capi.RtError, match="non-finite"):
        capi.Scene(v, *good[1:])
    v[3, 1, 2] = np.inf
    with pytest.raises(capi.RtError, match="non-finite"):
        capi.Scene(v, *good[1:])
    mid = g["mat_id"].copy()
    mid[5] = g["mats"].shape[0]
    with pytest.raises(capi.RtError, match="material id"):
        capi.Scene(good[0], good[1], good[2], mid, good[4])
    with pytest.raises(capi.RtError, match="negative radius"):
        capi.Scene(*good, None, np.array([[0, 0, 0, -1.0]], np.float32), np.zeros(1, np.int32))


def test_bvh_builder_invariants_host_only(pkg):
    """host/bvh_builder.cpp through rt_bvh_check (no GPU): every face in exactly one leaf, boxes nested,
    depth within the device stack -- on a regular height field, on heavily clustered and duplicated
    triangles (SAH degenerates: the depth guard must switch to median splits) and on tiny inputs."""
    capi = pkg.capi
    sc = pkg.scenes
    hv, hf = sc.heightfield_mesh(64)
    rng = np.random.default_rng(5)
    cases = {"heightfield": np.asarray(hv, np.float32)[hf].reshape(-1, 9)}
    tri = rng.uniform(-1, 1, (1, 9)).astype(np.float32)
    cases["4000 copies of one triangle"] = np.repeat(tri, 4000, axis=0)
    c = rng.uniform(-1, 1, (3000, 1, 3)).astype(np.float32) ** 7  # clustered towards the origin
    cases["clustered"] = (c + rng.normal(0, 1e-4, (3000, 3, 3)).astype(np.float32)).reshape(-1, 9)
    big = rng.uniform(-1, 1, (500, 9)).astype(np.float32)
    big[::50] *= 1e4  # a few huge triangles among small ones
    cases["mixed sizes"] = big
    cases["one"] = tri
    cases["none"] = np.zeros((0, 9), np.float32)
    for name, v in cases.items():
        for leaf in (1, 2, 4, 16):
            r = capi.bvh_check(v, leaf)
            assert r["refs"] == v.shape[0], name
            assert r["max_leaf"] <= leaf or v.shape[0] == 0, (name, r)
            assert r["depth"] <= 60, (name, r)
    # scenes that fit one leaf: the root pair must not keep an empty slot (its inverted box passes the
    # device's slab test); rt_bvh_check refuses such a tree
    for n in (1, 2, 3, 12):
        v = rng.uniform(-1, 1, (n, 9)).astype(np.float32)
        for leaf in (1, 2, 16):
            r = capi.bvh_check(v, leaf)
            assert r["refs"] == n and r["nodes"] >= 1
            assert r["leaves"] >= min(n, 2), (n, leaf, r)
    with pytest.raises(capi.RtError, match="non-finite"):
        bad = cases["mixed sizes"].copy()
        bad[7, 4] = np.nan
        capi.bvh_check(bad)


def test_adaptive_band_height_partitions_every_frame(pkg):
    """bench.band_rows_for: every rank gets ~8 bands; whatever it returns, the ranks' row maps must tile the
    image exactly once (rt_local_row_map / rt_local_rows are host-only)."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    capi = pkg.capi
    for H in (1, 7, 36, 1000, 1080, 2160, 4320):
        for world in (1, 2, 3, 4, 8):
            B = bench.band_rows_for(H, world)
            assert B >= 8 and B % 8 == 0
            rows = []
            for r in range(world):
                p = capi.make_params(64, H, band_rows=B, band_rank=r, band_world=world)
                rows.append(capi.local_row_map(p))
            allrows = np.sort(np.concatenate(rows))
            assert (allrows == np.arange(H)).all(), (H, world, B)
            if H >= 1080 and world > 1:
                sizes = [len(x) for x in rows]
                assert max(sizes) - min(sizes) <= B
                assert 4 <= -(-H // B) // world <= 9  # about 8 bands per rank


def test_loader_slices_parse_like_one_pass(pkg, tmp_path, monkeypatch):
    """The OBJ text is parsed in slices by several host threads and merged in file order.  Whatever the number of
    slices -- cuts fall between any two lines -- the bake is the one-pass bake, bit for bit: interleaved v / vn / f
    lines, several usemtl runs, face ids counted back from the vertices read so far (resolved against the vertex
    offset of their slice), vt/vn suffixes, blank and comment lines, CRLF."""
    capi = pkg.capi
    rng = np.random.default_rng(3)
    (tmp_path / "m.mtl").write_text("newmtl a\nKd 1 0 0\nnewmtl b\nKd 0 1 0\nillum 4\nnewmtl c\nKd 0 0 1\n")
    lines = ["mtllib m.mtl", "# generated"]
    nv = 0
    for block in range(60):
        lines.append("usemtl " + "abc"[block % 3])
        for _ in range(int(rng.integers(3, 40))):
            x, y, z = rng.normal(size=3)
            lines.append(f"v {x:.6f} {y:.6f} {z:.6f}" + ("\r" if block % 7 == 0 else ""))
            nv += 1
            if rng.random() < 0.2:
                lines.append(f"vn {rng.normal():.4f} {rng.normal():.4f} {rng.normal():.4f}")
        for _ in range(int(rng.integers(1, 30))):
            ids = rng.choice(nv, 3, replace=False)
            style = rng.integers(0, 3)
            if style == 0:
                lines.append("f " + " ".join(str(i + 1) for i in ids))
            elif style == 1:
                lines.append("f " + " ".join(str(int(i) - nv) for i in ids))         # relative ids
            else:
                lines.append("f " + " ".join(f"{i + 1}/1/{i + 1}" for i in ids))
        if block % 5 == 0:
            lines.append("")
    p = tmp_path / "sliced.obj"
    p.write_text("\n".join(lines) + "\n")
    monkeypatch.setenv("RT_LOAD_THREADS", "1")
    ref = capi.Mesh(str(p)).arrays()
    assert ref[0].shape[0] > 500
    for n in (2, 3, 7, 16, 64):
        monkeypatch.setenv("RT_LOAD_THREADS", str(n))
        got = capi.Mesh(str(p)).arrays()
        assert all(a.shape == b.shape for a, b in zip(got, ref)), n
        assert _bits(got[0], ref[0]) == 0 and _bits(got[1], ref[1]) == 0 and _bits(got[2], ref[2]) == 0, n
        assert (got[3] == ref[3]).all() and (got[4] == ref[4]).all(), n
    # a relative id that reaches before the first vertex of the file is an error in every slicing
    q = tmp_path / "bad_rel.obj"
    q.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\n" + "f 1 2 3\n" * 40 + "f 1 2 -4\n" + "v 2 2 2\n" * 40)
    for n in (1, 5):
        monkeypatch.setenv("RT_LOAD_THREADS", str(n))
        with pytest.raises(capi.RtError) as e:
            capi.Mesh(str(q))
        assert e.value.code == -4
