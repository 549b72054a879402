"""The oracle (oracle/rt_oracle.c, a C restatement of the reference hot path) against the golden
vectors that tests/golden/make_golden.py produced by running the reference itself, and -- when the
headless reference binary is present (oracle/_ref, build container / shipped to the GPU box) -- live
against that binary.  Everything here is bit-exact: face ids, hit parameter t, float RGB."""
from __future__ import annotations

import os

import numpy as np
import pytest

from conftest import GENERATED, SCENE_OF, case_params, load_golden, scene_arrays

FAST_CASES = list(SCENE_OF) + ["hf224_point_3840x2160_s24", "hf224_area_d5_g4_3840x2160_s24", "hf224m_area_d5_g4_3840x2160_s48"]


def _oracle_for(case, pkg, O, scene_dir, **kw):
    g = load_golden(case)
    verts, fn, vn, mid, mats = scene_arrays(case, pkg, scene_dir)
    cp = case_params(g)
    baked = O.BakedScene(verts, fn, vn, mid, mats, g["model_matrix"])
    orc = O.Oracle(baked, area=cp["area"], point=cp["point"], max_depth=cp["max_depth"], grid=cp["grid"],
                   light_color=g["light_color"], **kw)
    cam = O.Oracle.camera(g["eye"], g["view_inv"], g["viewport"], float(g["cam"][0]), float(g["cam"][1]))
    return g, orc, cam


@pytest.mark.parametrize("case", FAST_CASES)
def test_oracle_bit_exact_vs_reference_golden(case, pkg, oracle_mod, scene_dir):
    g, orc, cam = _oracle_for(case, pkg, oracle_mod, scene_dir)
    rgb, face, t, rgb8 = orc.render_pixels(cam, g["lights"], g["pxy"], threads=8)
    assert (face == g["face"]).all()
    assert (t.view(np.uint32) == g["t"].view(np.uint32)).all()
    assert (rgb.view(np.uint32) == g["rgb"].view(np.uint32)).all()
    # octree shape (leaves, inner nodes, face references, largest leaf) as the reference built it
    assert (orc.octree_stats() == g["octree_stats"]).all()
    mn, mx = orc.root_box()
    assert (mn == g["root_min"]).all() and (mx == g["root_max"]).all()


@pytest.mark.slow
@pytest.mark.parametrize("case", ["hf707_point_1920x1080_s20"])
def test_oracle_bit_exact_1m_triangles(case, pkg, oracle_mod, scene_dir):
    g, orc, cam = _oracle_for(case, pkg, oracle_mod, scene_dir)
    rgb, face, t, rgb8 = orc.render_pixels(cam, g["lights"], g["pxy"], threads=8)
    assert (face == g["face"]).all()
    assert (rgb.view(np.uint32) == g["rgb"].view(np.uint32)).all()
    assert (orc.octree_stats() == g["octree_stats"]).all()


def test_known_answers_of_the_default_scene(pkg, oracle_mod, scene_dir):
    """SURVEY.md section 4 / App. D known answers (cube.obj, 1000x1000, point light)."""
    O = oracle_mod
    g, orc, cam = _oracle_for("cube_point_1000", pkg, O, scene_dir)
    pxy, rgb, face, t, rgb8 = orc.render(cam, g["lights"], 1000, 1000, stride=1, threads=8)
    img = np.zeros((1000, 1000, 3), np.int64)
    img[pxy[:, 1], pxy[:, 0]] = O.quantize(rgb)
    bg = (img == 255).all(-1)
    assert int(bg.sum()) == 505791
    assert tuple(img[500, 500]) == (223, 223, 216)
    assert tuple(img[250, 400]) == (255, 255, 216)
    assert tuple(img[160, 160]) == (236, 236, 216)
    nb = np.argwhere(~bg)
    assert tuple(nb.min(0)) == (149, 149) and tuple(nb.max(0)) == (851, 851)
    assert len(np.unique(img.reshape(-1, 3), axis=0)) == 41
    assert set(np.unique(face)) == {-1, 10, 11}
    # camera known answers (App. D)
    out = np.zeros(3, np.float32)
    O.lib().or_screen_to_world(cam, 500.0, 500.0, out.ctypes.data)
    assert tuple(out) == (0.0, 0.0, 1.0)
    O.lib().or_screen_to_world(cam, 0.0, 0.0, out.ctypes.data)
    assert np.allclose(out, (-0.577350259, 0.577350259, 1.0), atol=1e-7)


def test_area_light_grid_known_answers(oracle_mod):
    """App. D: createAreaLight((-1,1,1), 0.3, 0.15, 5, 5).getPointLights()."""
    O = oracle_mod
    import ctypes as C
    p = O.OrParams()
    O.lib().or_default_params(C.byref(p))
    p.area_light, p.point_light = 1, 0
    out = np.zeros((25, 3), np.float32)
    light = np.array([-1, 1, 1], np.float32)
    n = O.lib().or_light_samples(C.byref(p), light.ctypes.data, out.ctypes.data)
    assert n == 25
    assert np.allclose(out[0], (-0.0700000003, 0.114999995, 1), atol=1e-8)
    assert np.allclose(out[1], (-0.07, 0.344999969, 1), atol=1e-8)
    assert np.allclose(out[5], (-0.210000008, 0.114999995, 1), atol=1e-8)
    assert np.allclose(out[24], (-0.629999995, 1.03499997, 1), atol=1e-8)


def test_quantiser(oracle_mod):
    q = oracle_mod.lib().or_quantize
    assert q(1.0) == 255 and q(0.85) == 216 and q(0.0) == 0 and q(2.0) == 255 and q(0.999) == 254
    assert q(-0.5) == -127  # no lower clamp in the reference (ppmIO.hpp:145)


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "ref_oracle")),
                    reason="headless reference binary not built (needs /root/reference)")
def test_oracle_live_vs_reference_binary(pkg, oracle_mod, tmp_path):
    """Fresh configuration not in the golden set: two lights, rotated camera, area light, gallery scene."""
    O = oracle_mod
    obj = str(tmp_path / "gallery.obj")
    pkg.scenes.write_gallery(obj, 2)
    out, dump = str(tmp_path / "o.bin"), str(tmp_path / "s.bin")
    O.run_ref(obj, out, 160, 120, area=1, point=0, stride=1, dump_scene=dump, lights=[(-1.2, 1.0, 1.4), (0.8, 1.5, 0.5)],
              cam_rot=(0.15, -0.35), cam_trans=(0.05, 0.1, -0.2))
    ref = O.load_render_dump(out)
    sc = O.load_scene_dump(dump)
    orc = O.Oracle(sc, area=1, point=0)
    rgb, face, t, _ = orc.render_pixels(orc.scene_camera(), sc.lights, ref.pxy, threads=8)
    assert (face == ref.face).all()
    assert (t.view(np.uint32) == ref.t.view(np.uint32)).all()
    assert (rgb.view(np.uint32) == ref.rgb.view(np.uint32)).all()


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "ref_oracle_patched")),
                    reason="headless reference binary not built (needs /root/reference)")
def test_patched_reference_is_identical_when_knobs_are_off(oracle_mod, tmp_path):
    """The sed-patched build (depth cap / grid knobs) must equal the unmodified reference at defaults."""
    O = oracle_mod
    obj = os.path.join(O.REF_SCENES, "cube.obj")
    a, b = str(tmp_path / "a.bin"), str(tmp_path / "b.bin")
    O.run_ref(obj, a, 200, 160, area=1, point=0, stride=1)
    O.run_ref(obj, b, 200, 160, area=1, point=0, stride=1, max_depth=1 << 30, grid=(5, 5))
    ra, rb = O.load_render_dump(a), O.load_render_dump(b)
    assert (ra.rgb.view(np.uint32) == rb.rgb.view(np.uint32)).all() and (ra.face == rb.face).all()
