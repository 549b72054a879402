"""The headless CLI / C++ facade (Flyscene::initialize + raytraceScene -> result.ppm) on the GPU,
against the reference's OWN result.ppm (written by the unmodified Flyscene::raytraceScene through
oracle/_ref/ref_oracle --mode rts) when that binary is present, and against the oracle otherwise."""
from __future__ import annotations

import json
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu

REF = os.path.join(ROOT, "oracle", "_ref", "ref_oracle")
CUBE = os.path.join(ROOT, "oracle", "_ref", "scenes", "cube.obj")


def read_p3(path):
    tok = open(path).read().split()
    assert tok[0] == "P3"
    w, h = int(tok[1]), int(tok[2])
    return np.array(tok[4:], np.int64).reshape(h, w, 3)


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(CUBE)), reason="reference binary / bundled scene not shipped")
@pytest.mark.parametrize("area,point", [(0, 1), (1, 0)])
def test_result_ppm_is_byte_identical_to_the_reference(area, point, pkg, tmp_path):
    """BASELINE configs[0]: bundled cube.obj, reference resolution 1000x1000 -> result.ppm."""
    cli = pkg.build.build_cli()
    ours = tmp_path / "ours"
    ours.mkdir()
    r = subprocess.run([cli, "--scene", CUBE, "--width", "1000", "--height", "1000", "--area", str(area), "--point", str(point)],
                       cwd=ours, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    stats = json.loads(r.stdout.strip().splitlines()[-1])
    assert stats["faces"] == 12
    # the reference's own frame driver (square images only; it crashes in ~ThreadPool after the file is written)
    rr = subprocess.run([REF, "--scene", CUBE, "--w", "1000", "--h", "1000", "--area", str(area), "--point", str(point),
                         "--mode", "rts"], capture_output=True, text=True, timeout=1200)
    m = re.search(r"scratch_cwd (\S+)", rr.stderr)
    assert m, rr.stderr[-500:]
    ref_ppm = os.path.join(m.group(1), "result.ppm")
    a, b = open(ours / "result.ppm", "rb").read(), open(ref_ppm, "rb").read()
    if a != b:
        ia, ib = read_p3(ours / "result.ppm"), read_p3(ref_ppm)
        diff = np.abs(ia - ib).max(-1)
        pytest.fail(f"result.ppm differs: {(diff > 0).sum()} pixels, max err {diff.max()}")
    print(f"result.ppm byte-identical to the reference ({len(a)} bytes), frame {stats['frame_ms']:.3f} ms")


def test_cli_on_generated_scene_matches_oracle(pkg, oracle_mod, tmp_path):
    """Full host path (OBJ loader -> BVH -> render -> P3 writer) on a generated scene, non-square image,
    two lights and a moved camera, compared with the oracle through the facade's own camera maths."""
    O = oracle_mod
    cli = pkg.build.build_cli()
    obj = str(tmp_path / "gallery.obj")
    pkg.scenes.write_gallery(obj, 2)
    r = subprocess.run([cli, "--scene", obj, "--width", "320", "--height", "200", "--area", "1", "--point", "0",
                        "--max-depth", "2", "--grid", "3", "3", "--light", "1.5", "1.0", "1.0"],
                       cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    img = read_p3(tmp_path / "result.ppm")
    assert img.shape == (200, 320, 3)
    mesh = pkg.capi.Mesh(obj)
    orc = O.Oracle(O.BakedScene(*mesh.arrays()), area=1, point=0, max_depth=2, grid=(3, 3))
    cam = O.Oracle.camera((0, 0, 2), np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 2]], np.float32), (0, 0, 320, 200),
                          60.0, np.float32(320) / np.float32(200))
    pxy, rgb, face, t, rgb8 = orc.render(cam, np.array([[-1, 1, 1], [1.5, 1.0, 1.0]], np.float32), 320, 200, stride=1)
    exp = np.clip(O.quantize(rgb), 0, 255)
    err = np.abs(img[pxy[:, 1], pxy[:, 0]] - exp).max(-1)
    assert (err <= 1).mean() >= 0.999 and err.max() <= 1, f"max err {err.max()}"


def test_cpp_facade_members_match_oracle(pkg, oracle_mod, tmp_path):
    """Every render-path member of the C++ facade (Flyscene / BoxTree / BoundingBox / arealight /
    Flycamera), called the way the reference's debug-ray tool calls them, against the oracle."""
    import ctypes as C
    O = oracle_mod
    probe = pkg.build.build_probe()
    obj = str(tmp_path / "gallery.obj")
    pkg.scenes.write_gallery(obj, 3)
    px, py = 300.0, 260.0
    r = subprocess.run([probe, obj, str(px), str(py)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout[r.stdout.index("{"):])
    mesh = pkg.capi.Mesh(obj)
    arrs = mesh.arrays()
    orc = O.Oracle(O.BakedScene(*arrs), area=1, point=0, max_depth=2, grid=(5, 5))
    W, H = 640, 480
    cam = O.Oracle.camera((0, 0, 2), np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 2]], np.float32), (0, 0, W, H), 60.0,
                          np.float32(W) / np.float32(H))
    f32 = lambda v: np.array(v, np.float32)
    screen = np.zeros(3, np.float32)
    O.lib().or_screen_to_world(cam, px, py, screen.ctypes.data)
    assert (f32(out["screen"]) == screen).all() and out["origin"] == [0, 0, 2]
    assert out["faces"] == arrs[0].shape[0]
    mn, mx = orc.root_box()
    assert (f32(out["root_min"]) == mn).all() and (f32(out["root_max"]) == mx).all()
    origin = f32([0, 0, 2])
    d = (screen - origin).astype(np.float32)
    dest = (d + origin).astype(np.float32)
    ids = np.zeros(arrs[0].shape[0], np.int32)
    n = O.lib().or_octree_candidates(orc.handle, origin.ctypes.data, dest.ctypes.data, ids.ctypes.data, len(ids))
    assert out["candidates"] == n and out["candidate_sum"] == int(ids[:n].astype(np.int64).sum())
    lights = np.array([[-1, 1, 1], [1.5, 1.0, 1.0]], np.float32)
    rgb = np.zeros(3, np.float32); face = np.zeros(1, np.int32); t = np.zeros(1, np.float32)
    O.lib().or_trace_ray(orc.handle, origin.ctypes.data, d.ctypes.data, 0, lights.ctypes.data, 2, rgb.ctypes.data,
                         face.ctypes.data, t.ctypes.data)
    assert out["best"] == int(face[0]) and np.float32(out["t"]) == t[0]
    assert np.allclose(f32(out["colour"]), rgb, rtol=2e-6, atol=1e-7)
    if face[0] >= 0:
        hit = (origin + t[0] * d).astype(np.float32)
        vis = np.zeros(2, np.uint8)
        anyv = O.lib().or_light_strikes(orc.handle, hit.ctypes.data, lights.ctypes.data, 2, vis.ctypes.data)
        assert out["light_any"] == anyv and out["light_vis"] == [int(vis[0]), int(vis[1])]
        ph = np.zeros(3, np.float32)
        O.lib().or_phong_shade(orc.handle, origin.ctypes.data, hit.ctypes.data, int(face[0]), lights.ctypes.data, 2, ph.ctypes.data)
        assert np.allclose(f32(out["phong"]), ph, rtol=2e-6, atol=1e-7)
    # headless debug ray: every level is a traceRay from the previous hit point along the mirror direction
    dbg = out["debug_ray"]
    assert dbg and dbg[0]["level"] == 0 and dbg[0]["face"] == int(face[0])
    for k, lv in enumerate(dbg):
        o_k, d_k = f32(lv["origin"]), f32(lv["direction"])
        e_rgb = np.zeros(3, np.float32); e_face = np.zeros(1, np.int32); e_t = np.zeros(1, np.float32)
        # (the facade lowers the depth cap by the level it starts at, like Flyscene::traceRay's level argument)
        orc_k = O.Oracle(O.BakedScene(*arrs), area=1, point=0, max_depth=max(0, 2 - k), grid=(5, 5))
        O.lib().or_trace_ray(orc_k.handle, o_k.ctypes.data, d_k.ctypes.data, 0, lights.ctypes.data, 2, e_rgb.ctypes.data,
                             e_face.ctypes.data, e_t.ctypes.data)
        assert lv["face"] == int(e_face[0]), (k, lv)
        if lv["face"] < 0:
            assert k == len(dbg) - 1
            break
        assert np.float32(lv["t"]) == e_t[0]
        assert np.allclose(f32(lv["colour"]), e_rgb, rtol=2e-6, atol=1e-7)
        hit_k = (o_k + e_t[0] * d_k).astype(np.float32)
        vis_k = np.zeros(2, np.uint8)
        O.lib().or_light_strikes(orc_k.handle, hit_k.ctypes.data, lights.ctypes.data, 2, vis_k.ctypes.data)
        assert lv["visible"] == [int(vis_k[0]), int(vis_k[1])]
        if k + 1 < len(dbg):  # next level starts at this hit point along the mirror direction
            nrm = arrs[1][lv["face"]].astype(np.float32)
            dn = np.float32(d_k[0] * nrm[0] + np.float32(d_k[1] * nrm[1] + d_k[2] * nrm[2]))
            refl = (d_k - (np.float32(2) * dn) * nrm).astype(np.float32)
            assert (f32(dbg[k + 1]["origin"]) == hit_k).all() and (f32(dbg[k + 1]["direction"]) == refl).all()
    assert out["n_samples"] == 25
    assert np.allclose(out["sample0"], (-0.0700000003, 0.114999995, 1), atol=1e-8)
    assert np.allclose(out["sample24"], (-0.629999995, 1.03499997, 1), atol=1e-8)
    assert out["bb_hit"] == 1 and out["bb_miss"] == 0
    assert out["octree"] == [int(x) for x in orc.octree_stats()]


def test_cli_multi_gpu_result_ppm_is_byte_identical(pkg, tmp_path):
    """`rt_cli --gpus N` (Flyscene::setDevices -> rt_multi_*: the reference's ThreadPool fan-out with GPUs as the
    workers) writes the same result.ppm, byte for byte, as the single-GPU run."""
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    cli = pkg.build.build_cli()
    obj = str(tmp_path / "gallery.obj")
    pkg.scenes.write_gallery(obj, 2)
    args = ["--scene", obj, "--width", "500", "--height", "301", "--area", "1", "--point", "0", "--max-depth", "3", "--grid", "4", "4"]
    outs = []
    for gpus in (1, n):
        d = tmp_path / f"g{gpus}"
        d.mkdir()
        r = subprocess.run([cli] + args + ["--gpus", str(gpus)], cwd=d, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        outs.append(open(d / "result.ppm", "rb").read())
    assert outs[0] == outs[1], "multi-GPU result.ppm differs from the single-GPU one"
